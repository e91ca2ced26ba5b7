mkdir -p gpurun_out
timeout 300 python tools/k1_target.py ens --steps 200 > gpurun_out/s3_ens.txt 2>&1
timeout 900 python -m pytest tests/test_ensemble_gpu.py tests/test_catalog_gpu.py tests/test_dropin_gpu.py -x -q -m gpu > gpurun_out/s3_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_test.log
tail -5 gpurun_out/s3_test.log; grep -v Warn gpurun_out/s3_ens.txt | tail -3
