# final validation of a build: GPU tests, smoke, default bench, single-source timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
CHAIN_PROTOCOL=4 timeout 300 python tools/single_source_time.py > gpurun_out/single_source_p4.txt 2>&1
tail -4 gpurun_out/gputest.log; tail -2 gpurun_out/bench.err; grep -v Warn gpurun_out/single_source_p4.txt | head -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "sustained", d["sustained"]["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["clocks"])
print("ens", d["ensemble"]["ms_per_step"], d["ensemble"]["evals_per_s"], d["sources_fitted_per_s"], "fit", d["e2e_fit"]["sources_per_s"])
print("dict", d["dictionary_build"]["wall_s"], d["dictionary_build"]["entries_per_s"])
for k,v in d["configs"].items(): print(k, v["ms_per_step"], v["roofline_frac"], v["kernel"])
PY
