mkdir -p gpurun_out
timeout 3000 python -m pytest tests/test_gpu_fullsize.py -x -q -s -k "crop or level10_matches" > gpurun_out/t14.log 2>&1; echo "pytest exit $?" >> gpurun_out/t14.log
grep -v "^$" gpurun_out/t14.log | grep -v Warning | tail -25
