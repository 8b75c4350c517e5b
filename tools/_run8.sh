mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tile_upload or process_clouds" > gpurun_out/t8.log 2>&1; echo "pytest exit $?" >> gpurun_out/t8.log; tail -5 gpurun_out/t8.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/b8.json 2> gpurun_out/b8.err; echo "bench exit $?"; tail -5 gpurun_out/b8.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/b8.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','single_stream','concurrent','e2e','e2e_process_cloud','e2e_las','clocks','cpu_baseline','stage_ms'):
    print(k, l[k])
print('roofline', {k:v for k,v in l['roofline'].items()})
PY
