mkdir -p gpurun_out
timeout 300 python tools/k1_target.py plain > gpurun_out/k1_plain.txt 2>&1
timeout 600 python -m pytest tests/test_k1_gpu.py -x -q -m gpu > gpurun_out/gputest_k1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_k1.log
tail -15 gpurun_out/gputest_k1.log; grep -v Warn gpurun_out/k1_plain.txt | tail -3
