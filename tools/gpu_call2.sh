mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest.log
timeout 300 ./tools/ubench_gather 5991 > gpurun_out/ubench_gather.txt 2>&1
timeout 300 python tools/k1_target.py plain > gpurun_out/k1_plain.txt 2>&1
timeout 300 python tools/k1_target.py ens > gpurun_out/k1_ens.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_coop_kernel --launch-skip 4 -c 1 -f -o gpurun_out/k1_r02_plain python tools/k1_target.py plain --steps 3 > gpurun_out/ncu_plain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_coop_kernel --launch-skip 14 -c 1 -f -o gpurun_out/k1_r02_ens python tools/k1_target.py ens --steps 3 > gpurun_out/ncu_ens.log 2>&1
tail -3 gpurun_out/gputest.log; cat gpurun_out/ubench_gather.txt gpurun_out/k1_plain.txt gpurun_out/k1_ens.txt | tail -40
