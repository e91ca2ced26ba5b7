mkdir -p gpurun_out
export AGNFIT_B200_LIB=$PWD/build_alt/libagnfit_b200.so
for r in 7; do timeout 300 python tools/persist_prof.py $r 2>&1 | grep -v Warn | tail -1; done > gpurun_out/s3_persist.txt
cat gpurun_out/s3_persist.txt
