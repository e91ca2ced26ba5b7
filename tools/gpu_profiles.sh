# ncu captures for profiles/ (round 2): each profiled command has run without ncu first (tools/gpu_full.sh).
# Summaries are produced on the box too (gpurun copies back at most 64 MiB).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_coop_kernel --launch-skip 4 -c 1 -f -o gpurun_out/k1_r02_coop_full python tools/k1_target.py plain --steps 3 > gpurun_out/ncu_plain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_coop_kernel --launch-skip 14 -c 1 -f -o gpurun_out/k1_r02_ens_full python tools/k1_target.py ens --steps 3 > gpurun_out/ncu_ens.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k2_contract_mma -c 1 -f -o gpurun_out/k2_r02_mma_full python tools/k2_wall.py > gpurun_out/ncu_k2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k3_split_kernel --launch-skip 3 -c 1 -f -o gpurun_out/k3_r02_split_full python tools/k1_target.py ens --steps 3 > gpurun_out/ncu_k3split.log 2>&1
python tools/ncu_summary.py gpurun_out/k3_r02_split_full.ncu-rep gpurun_out/k3_r02_split_ncu.md > /dev/null 2>&1
rm -f gpurun_out/k3_r02_split_full.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:k3_select_kernel -c 1 -f -o gpurun_out/k3_r02_select_full python -m pytest tests/test_ensemble_gpu.py -q -m gpu -k exact_percentiles_match > gpurun_out/ncu_k3.log 2>&1
python tools/ncu_summary.py gpurun_out/k1_r02_coop_full.ncu-rep gpurun_out/k1_r02_coop_ncu.md --traffic k1_coop_kernel 10000000 > /dev/null 2>&1
python tools/ncu_l1.py gpurun_out/k1_r02_coop_full.ncu-rep k1_coop_kernel 'ILb0ELi1ELb0ELi9ELb0ELi4ELb1' 1e7 > gpurun_out/k1_r02_coop_l1.txt 2>&1
python tools/ncu_lines.py gpurun_out/k1_r02_coop_full.ncu-rep k1_coop_kernel 'ILb0ELi1ELb0ELi9ELb0ELi4ELb1' > gpurun_out/k1_r02_coop_lines.txt 2>&1
python tools/ncu_summary.py gpurun_out/k1_r02_ens_full.ncu-rep gpurun_out/k1_r02_ens_ncu.md > /dev/null 2>&1
python tools/ncu_lines.py gpurun_out/k1_r02_ens_full.ncu-rep k1_coop_kernel 'ILb0ELi1ELb0ELi9ELb1ELi3ELb1' > gpurun_out/k1_r02_ens_lines.txt 2>&1
python tools/ncu_l1.py gpurun_out/k1_r02_ens_full.ncu-rep k1_coop_kernel 'ILb0ELi1ELb0ELi9ELb1ELi3ELb1' 5e6 > gpurun_out/k1_r02_ens_l1.txt 2>&1
python tools/ncu_summary.py gpurun_out/k2_r02_mma_full.ncu-rep gpurun_out/k2_r02_mma_ncu.md > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/k3_r02_select_full.ncu-rep gpurun_out/k3_r02_select_ncu.md > /dev/null 2>&1
rm -f gpurun_out/k1_r02_ens_full.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-config-legs --fit-sources 0 --ensemble-steps 20 --sustain-s 0 --mux-sources 0 > gpurun_out/ncu_launches.log 2>&1
CHAIN_PROTOCOL=4 timeout 300 python tools/single_source_time.py > gpurun_out/single_source_p4.txt 2>&1
tail -n 2 gpurun_out/ncu_plain.log gpurun_out/ncu_ens.log gpurun_out/ncu_k2.log gpurun_out/ncu_k3.log; tail -n 7 gpurun_out/single_source_p4.txt; du -sh gpurun_out; ls -la gpurun_out
