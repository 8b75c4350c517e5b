mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "knn or region_growing" > gpurun_out/t5.log 2>&1; echo "pytest exit $?" >> gpurun_out/t5.log
tail -4 gpurun_out/t5.log
timeout 600 python tools/bench_knn.py --reps 2 --cells 0 0.09 --tune knn_tiers=1 knn_tier1_min_block=15,knn_tier2_min_block=15 knn_tier1_min_block=60,knn_tier2_min_block=60 knn_tier2_max_block=2000 > gpurun_out/bk5.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:knn_seg_kernel -c 2 -o gpurun_out/prof_seg_v3 python tools/bench_knn.py --reps 1 --cells 0 > gpurun_out/ncu5.log 2>&1
cat gpurun_out/bk5.log; tail -3 gpurun_out/ncu5.log
