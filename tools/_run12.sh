mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_parity.py -x -q -s -k "radius_method or region_growing_matches" > gpurun_out/t12.log 2>&1; echo "pytest exit $?" >> gpurun_out/t12.log
grep -v "^$" gpurun_out/t12.log | grep -v Warning | tail -30
