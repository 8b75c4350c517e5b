mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -s -k "knn or region_growing or config1 or shared_npz" > gpurun_out/t7.log 2>&1; echo "pytest exit $?" >> gpurun_out/t7.log
grep -v "^$" gpurun_out/t7.log | grep -v Warning | tail -25
