mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "knn or region_growing" > gpurun_out/t4.log 2>&1; echo "pytest exit $?" >> gpurun_out/t4.log
tail -8 gpurun_out/t4.log
timeout 900 python tools/bench_knn.py --reps 2 --cells 0 0.09 0.08 --tune knn_tiers=1 > gpurun_out/bk4.log 2>&1; echo "bench_knn exit $?" >> gpurun_out/bk4.log
cat gpurun_out/bk4.log
