# round-end evidence on one GPU: GPU tests, smoke(), default bench (+ reference arm is run by the driver)
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; tail -3 gpurun_out/r2_final_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'parity',d['parity_vs_oracle'])
print('stages',d['stages_ms']); print('cfg2',d['latency_cfg2']['p50_ms'], d['latency_cfg2']['e2e_p50_ms']); print('cpp',d['e2e_cpp']['value'], 'wire', d['e2e_wire']['value']); print('cpu', d['cpu_baseline']['value'])
PY
