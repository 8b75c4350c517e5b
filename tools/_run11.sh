mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -s -k "ahn_preprocessing or las_ingest or process_folder" > gpurun_out/t11.log 2>&1; echo "pytest exit $?" >> gpurun_out/t11.log
grep -v "^$" gpurun_out/t11.log | grep -v Warning | tail -30
