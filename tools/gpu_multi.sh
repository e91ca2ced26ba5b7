# multi-GPU bench record: bash tools/gpu_multi.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -c 600 gpurun_out/bench_n$N.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("value", d["value"], "sustained", d["sustained"]["value"], "ens", d["ensemble"]["evals_per_s"], d["ensemble"]["sources_fitted_per_s"], "fit", d["e2e_fit"]["sources_per_s"], "e2e", d["e2e"]["value"], "strong", d.get("strong",{}).get("value"))
PY
