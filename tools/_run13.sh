mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_parity.py -x -q -s -k "demo_laz or las_ingest or process_folder or ahn_preprocessing" > gpurun_out/t13.log 2>&1; echo "pytest exit $?" >> gpurun_out/t13.log
grep -v "^$" gpurun_out/t13.log | grep -v Warning | tail -12
