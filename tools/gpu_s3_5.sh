mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest.log
timeout 300 python tools/single_source_time.py > gpurun_out/single_source.txt 2>&1
timeout 300 python tools/k1_target.py ens --steps 200 > gpurun_out/s3_ens.txt 2>&1
timeout 300 python tools/k1_target.py plain --steps 50 > gpurun_out/s3_plain.txt 2>&1
tail -5 gpurun_out/gputest.log; grep -v Warn gpurun_out/single_source.txt | tail -8; grep -v Warn gpurun_out/s3_ens.txt | tail -1; grep -v Warn gpurun_out/s3_plain.txt | tail -1
