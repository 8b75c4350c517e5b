mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -s -k "knn or region_growing or config1 or shared_npz" > gpurun_out/t6.log 2>&1; echo "pytest exit $?" >> gpurun_out/t6.log
grep -v "^$" gpurun_out/t6.log | tail -25
timeout 600 python tools/bench_knn.py --reps 2 --cells 0 > gpurun_out/bk6.log 2>&1; cat gpurun_out/bk6.log
