mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -s -k "buffered or refine_ground or pip or demo_tile or pipeline_stages" > gpurun_out/t10.log 2>&1; echo "pytest exit $?" >> gpurun_out/t10.log
grep -v "^$" gpurun_out/t10.log | grep -v Warning | tail -30
