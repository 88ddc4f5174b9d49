set -x
timeout 60 ./build/shim_test 2>&1 | tail -4
timeout 120 python tools/ct_mode_bench.py cfg4 8 5 > gpurun_out/r2c35_ct_mode_cfg4.json 2> gpurun_out/r2c35_ct_mode.err; cat gpurun_out/r2c35_ct_mode_cfg4.json; tail -3 gpurun_out/r2c35_ct_mode.err
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ct_multiply_mode_db_multiply and 4096-16-16-2-11" > gpurun_out/r2c35_sanitizer_ctmul.log 2>&1; echo "sanitizer rc=$?"; tail -6 gpurun_out/r2c35_sanitizer_ctmul.log
