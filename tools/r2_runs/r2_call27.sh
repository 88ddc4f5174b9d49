set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-shim > gpurun_out/r2c27_bench.json 2> gpurun_out/r2c27_bench.err; tail -3 gpurun_out/r2c27_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c27_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'parity',d['parity_vs_oracle'])
print('stages',d['stages_ms']); print('cfg2', d['latency_cfg2']['p50_ms'], d['latency_cfg2']['stages_ms'])
PY
