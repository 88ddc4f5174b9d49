set -x
timeout 300 ./build/shim_test 2>&1 | tail -8
timeout 300 ./build/shim_bench 4194304 256 2 4096 20 8 10 1 2>&1 | tail -2
timeout 300 ./build/shim_bench 65536 288 2 4096 24 1 50 1 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c15_l.csv python tools/shard_probe.py cfg4 8 8 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c15_l.csv')) if len(r)>10]
hdr=rows[0]
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[ii], r[ki].split('(')[0][:40]), {})[r[mi]]=r[vi]
for (i,k),m in list(d.items())[-14:]:
    print(i,k,m)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ks_level_cluster -s 9 -c 1 -o gpurun_out/r2c15_ks_l9 -f python tools/run_answer.py cfg4 8 1 > gpurun_out/r2c15_ncu_ks.log 2>&1; tail -2 gpurun_out/r2c15_ncu_ks.log
