mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_plain.log 2>&1 &&
timeout 1800 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r2_sanitizer_memcheck.log python __graft_entry__.py smoke > gpurun_out/smoke_memcheck.log 2>&1
echo "memcheck exit $?"; tail -3 gpurun_out/smoke_plain.log; tail -15 gpurun_out/r2_sanitizer_memcheck.log; tail -3 gpurun_out/smoke_memcheck.log
