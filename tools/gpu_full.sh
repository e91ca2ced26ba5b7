mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -5 gpurun_out/gputest.log; tail -3 gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"])
print("ens", d["ensemble"]["ms_per_step"], d["ensemble"]["evals_per_s"], "fit", d["e2e_fit"]["sources_per_s"])
print("dict", d["dictionary_build"]["wall_s"], d["dictionary_build"]["entries_per_s"], d["dictionary_build"].get("modelset_wall_s"))
for k,v in d["configs"].items(): print(k, v["ms_per_step"], v["roofline_frac"], v["kernel"])
PY
