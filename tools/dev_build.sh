#!/bin/bash
# Quick development build of the cooperative K1 kernel (classic model set only, -DKC_DEV) into build_alt/;
# use with AGNFIT_B200_LIB=build_alt/libagnfit_b200.so.  Extra nvcc flags: $EXTRA.
set -e
cd "$(dirname "$0")/../agnfitter_b200/csrc"
OUT=${OUT:-../../build_alt}
mkdir -p $OUT
F="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $EXTRA"
nvcc $F -Xptxas -v -DKC_DEV -c k1_coop.cu -o $OUT/k1_coop.o > $OUT/k1_coop.log 2>&1 &
nvcc $F -Xptxas -v -DKC_DEV -c k1_coop_ens.cu -o $OUT/k1_coop_ens.o > $OUT/k1_coop_ens.log 2>&1 &
nvcc $F -c agnfit_capi.cu -o $OUT/agnfit_capi.o > $OUT/capi.log 2>&1 &
wait
grep -E "error" $OUT/*.log && exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libagnfit_b200.so $OUT/agnfit_capi.o $OUT/k1_coop.o $OUT/k1_coop_ens.o k1_lnprob.o k1_g4.o k2_dictbuild.o k3_ensemble.o k4_sed.o peaks.o -Xcompiler -fPIC
ls -la $OUT/libagnfit_b200.so
