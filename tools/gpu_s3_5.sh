mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest.log
timeout 300 python tools/line_check.py --ens > gpurun_out/s3_line.txt 2>&1
timeout 300 python tools/line_check.py --bands 49 --sources 20000 --ens > gpurun_out/s3_line49.txt 2>&1
timeout 300 python tools/line_check.py --c4 --bands 49 --sources 20000 > gpurun_out/s3_line_c4.txt 2>&1
timeout 300 python tools/line_check.py --ga-free --sources 20000 > gpurun_out/s3_line_gf.txt 2>&1
tail -5 gpurun_out/gputest.log; for f in s3_line s3_line49 s3_line_c4 s3_line_gf; do grep "pitch mult" gpurun_out/$f.txt; done
