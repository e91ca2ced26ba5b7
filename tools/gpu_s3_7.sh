mkdir -p gpurun_out
export AGNFIT_B200_LIB=$PWD/build_alt/libagnfit_b200.so
timeout 600 python tools/line_check.py --ens > gpurun_out/s3_line.txt 2>&1
grep -v Warn gpurun_out/s3_line.txt | tail -8
