mkdir -p gpurun_out
echo "main lib (4 CTAs):"; timeout 300 python tools/k1_target.py plain --steps 2000 2>&1 | grep -v Warn | tail -1
export AGNFIT_B200_LIB=$PWD/build_alt/libagnfit_b200.so
echo "3 CTAs:"; timeout 300 python tools/k1_target.py plain --steps 2000 2>&1 | grep -v Warn | tail -1
timeout 300 python tools/k1_target.py plain --steps 2000 --opt coop_line=0 2>&1 | grep -v Warn | tail -1
