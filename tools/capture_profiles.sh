#!/bin/bash
# ncu evidence of one pipeline pass of bench.py, sized to come back through gpurun (<= 64 MiB under gpurun_out/):
#   gpurun --timeout 1500 -- 'bash tools/capture_profiles.sh r2'
# writes gpurun_out/launches_<tag>.csv   every launch of a 1-step run, gpu__time_duration only (cold, serialised), with
#                                        the NVTX range stack of every launch (the ranges are named like bench.py's brackets)
#        gpurun_out/raw_<tag>.csv        --set full counters of every kernel of the first pipeline pass, same NVTX column
# The .ncu-rep itself stays on the box (/tmp).  Then here:
#   python tools/launch_summary.py gpurun_out/launches_<tag>.csv
#   python tools/traffic_from_raw.py gpurun_out/raw_<tag>.csv "<what was captured>" > profiles/<tag>_traffic.json
# Per-line source counters of single kernels: tools/ncu_lines2.py on a report captured with --import-source on.
set -u
tag=${1:-capture}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --min-warmup 0 --no-cpu-baseline --no-e2e --workers 1"
# a number printed by a run under ncu is never a bench value
timeout 400 ncu --nvtx --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/launches_${tag}.csv $CMD > gpurun_out/ncu_${tag}_launches.log 2>&1
echo "launch list exit $?"
# --set full of the bracketed kernels of the first pipeline pass.  Every replay pass restores the GBs of workspace a
# kernel writes (~5 s per kernel), so the capture is confined to the NVTX ranges of the brackets: the S4 stage
# (~25 launches), the growth, the raster and point-in-polygon launches.
timeout 1200 ncu --nvtx --nvtx-include "knn_stage (build+search+epilogue)/" --nvtx-include "grow_kernel/" \
    --nvtx-include "raster_kernel/" --nvtx-include "pip_query_kernel/" \
    --set full --clock-control none -c 60 -f -o /tmp/prof_${tag} $CMD > gpurun_out/ncu_${tag}_full.log 2>&1
echo "full capture exit $?"
if ! ncu -i /tmp/prof_${tag}.ncu-rep --page raw --csv 2>/dev/null | grep -q knn_epilogue_kernel; then
    echo "NVTX filter matched nothing: falling back to kernel names"
    timeout 1200 ncu --nvtx --set full --clock-control none -k 'regex:knn_|grow_|raster_kernel|pip_query_kernel' -c 40 \
        -f -o /tmp/prof_${tag} $CMD > gpurun_out/ncu_${tag}_full.log 2>&1
    echo "fallback capture exit $?"
fi
ncu -i /tmp/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_${tag}.csv 2>/dev/null
ls -la /tmp/prof_${tag}.ncu-rep gpurun_out/*_${tag}*
du -sh gpurun_out
