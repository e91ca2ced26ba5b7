mkdir -p gpurun_out
timeout 300 python tools/k1_target.py plain > gpurun_out/k1_plain.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_g4_kernel --launch-skip 4 -c 1 -f -o gpurun_out/k1_r02_g4 python tools/k1_target.py plain --steps 3 > gpurun_out/ncu_g4.log 2>&1
grep -v Warn gpurun_out/k1_plain.txt | tail -3; tail -3 gpurun_out/ncu_g4.log
