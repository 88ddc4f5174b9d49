set -x
timeout 118 python bench.py > gpurun_out/r2c37_bench_default.json 2> gpurun_out/r2c37_bench_default.err; echo "rc=$?"
tail -c 2500 gpurun_out/r2c37_bench_default.json; tail -3 gpurun_out/r2c37_bench_default.err
