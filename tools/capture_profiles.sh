#!/bin/bash
# ncu evidence of one pipeline pass of bench.py, sized to come back through gpurun (<= 64 MiB under gpurun_out/):
#   gpurun --timeout 900 -- 'bash tools/capture_profiles.sh r2_step1'
# writes gpurun_out/launches_<tag>.csv   every launch, gpu__time_duration only (cold, serialised)
#        gpurun_out/raw_<tag>.csv        --set full counters of the bracketed kernels, first pipeline pass
#        gpurun_out/src_<tag>.csv        per-instruction source page of the same kernels (-lineinfo)
# The .ncu-rep itself stays on the box (/tmp): with --import-source a whole pass comes to ~280 MB.
# Then here:  python tools/launch_summary.py gpurun_out/launches_<tag>.csv
#             python tools/traffic_from_raw.py gpurun_out/raw_<tag>.csv "<what was captured>" > profiles/<tag>_traffic.json
#             python tools/ncu_lines.py gpurun_out/src_<tag>.csv <sass> <mangled kernel> (see its docstring)
set -u
tag=${1:-capture}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --min-warmup 0 --no-cpu-baseline --no-e2e --workers 1"
KERNELS='regex:knn_block_kernel|knn_select_kernel|knn_epilogue_kernel|grow_kernel|grow_seed_pass_kernel|cc_link_kernel|raster_kernel|pip_query_kernel'
# a number printed by a run under ncu is never a bench value
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches_${tag}.csv $CMD > gpurun_out/ncu_${tag}_launches.log 2>&1
# first pipeline pass only: 5 raster + 3 pip + 3 cc_link + 3 knn + epilogue + seed pass + 2 grow launches
timeout 600 ncu --set full --clock-control none --import-source on -k "$KERNELS" -c 19 -f -o /tmp/prof_${tag} \
    $CMD > gpurun_out/ncu_${tag}_full.log 2>&1
ncu -i /tmp/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_${tag}.csv 2>/dev/null
ncu -i /tmp/prof_${tag}.ncu-rep --page source --csv > gpurun_out/src_${tag}.csv 2>/dev/null
ls -la /tmp/prof_${tag}.ncu-rep gpurun_out/*_${tag}*
du -sh gpurun_out
