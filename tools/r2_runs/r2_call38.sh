set -x
PIRB_GRAPHS=0 timeout 75 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/r2c38_ct_launches.csv python tools/ct_mode_bench.py cfg4 8 1 > gpurun_out/r2c38_ct.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/r2c38_ct.log; wc -l gpurun_out/r2c38_ct_launches.csv
