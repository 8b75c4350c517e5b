set -x
mkdir -p gpurun_out
timeout 600 python tools/bench_knn.py --reps 2 --cells 0 0.09 > gpurun_out/bk3.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:knn_seg_kernel -c 2 -o gpurun_out/prof_seg20 python tools/bench_knn.py --reps 1 --cells 0 > gpurun_out/ncu3.log 2>&1
tail -5 gpurun_out/bk3.log; tail -3 gpurun_out/ncu3.log
