set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c24_bench.json 2> gpurun_out/r2c24_bench.err; tail -3 gpurun_out/r2c24_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c24_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'parity',d['parity_vs_oracle'])
print('stages',d['stages_ms']); print('cfg2',d['latency_cfg2']); print('cpp',d['e2e_cpp']); print('wire',d['e2e_wire'])
PY
python tools/shard_probe.py cfg4 8 1 2>&1 | tail -2
python bench.py --workload cfg2 --steps 20 --no-cpu-baseline --no-shim 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 bench', d['value'], d['ms_per_step'], 'roof', d['roofline']['frac'], d['stages_ms'])"
