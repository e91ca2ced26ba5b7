mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_k1_gpu.py tests/test_ensemble_gpu.py -x -q -m gpu > gpurun_out/gputest_k1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_k1.log
timeout 300 python tools/k1_target.py plain > gpurun_out/k1_plain.txt 2>&1
timeout 300 python tools/k1_target.py ens > gpurun_out/k1_ens.txt 2>&1
timeout 300 python tools/k1_target.py plain --opt k1_variant=2 > gpurun_out/k1_plain_coop.txt 2>&1
tail -15 gpurun_out/gputest_k1.log; tail -2 gpurun_out/k1_plain.txt gpurun_out/k1_ens.txt gpurun_out/k1_plain_coop.txt
