set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "knn or region_growing" > gpurun_out/t1.log 2>&1; echo "pytest exit $?" >> gpurun_out/t1.log
tail -25 gpurun_out/t1.log
timeout 900 python tools/bench_knn.py --reps 2 > gpurun_out/bk1.log 2>&1; echo "bench_knn exit $?" >> gpurun_out/bk1.log
tail -20 gpurun_out/bk1.log
