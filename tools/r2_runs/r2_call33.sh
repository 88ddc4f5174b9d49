set -x
timeout 120 ./build/shim_test 2>&1 | tail -4
timeout 120 ./build/shim_bench 4194304 256 2 4096 20 8 20 1 > gpurun_out/r2c33_shim.json 2> gpurun_out/r2c33_shim.err; cat gpurun_out/r2c33_shim.json; tail -3 gpurun_out/r2c33_shim.err
