# round-end evidence on one GPU: tests, smoke, default bench + reference arm, ncu launch list and full captures
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -c 600 gpurun_out/final_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -c 400 gpurun_out/final_bench_ref.json
for w in cfg1 cfg3 cfg4 cfg5b; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/final_bench_other.jsonl; done
python bench.py --workload cfg2 --queries-per-gpu 16 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/final_bench_other.jsonl
python bench.py --workload cfg4 --queries-per-gpu 16 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/final_bench_other.jsonl
python bench.py --workload cfg4 --queries-per-gpu 8 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/final_bench_other.jsonl
python bench.py --workload cfg5a --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/final_bench_other.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/final_bench_other.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:28], d['config']['queries_per_step'], round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), d['stages_ms'])
PY
PIRB_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan -s 3 -c 1 -o gpurun_out/final_scan_cfg2 -f python tools/ncu_level6.py 1 > gpurun_out/final_ncu_scan.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ks_level_cluster -s 13 -c 1 -o gpurun_out/final_ks_l6 -f python tools/ncu_level6.py 1 > gpurun_out/final_ncu_ks.log 2>&1
ls -la gpurun_out/final_*
