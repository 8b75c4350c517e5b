set -x
mkdir -p gpurun_out
./tools/micro/dp_rate > gpurun_out/dp_rate.log 2>&1; cat gpurun_out/dp_rate.log
timeout 600 python tools/bench_knn.py --reps 2 --points 4000000 --cells 0 > gpurun_out/bk2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_seg_kernel -c 2 -o gpurun_out/prof_seg python tools/bench_knn.py --reps 1 --points 4000000 --cells 0 > gpurun_out/ncu2.log 2>&1
tail -5 gpurun_out/bk2.log; tail -5 gpurun_out/ncu2.log; ls -la gpurun_out/
