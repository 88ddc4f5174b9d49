set -x
PIRB_KS_V2_MIN_NODES=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not shim and not peer_memory" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-shim --no-cfg2 > gpurun_out/r2c32_bench.json 2> gpurun_out/r2c32_bench.err; tail -3 gpurun_out/r2c32_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c32_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'parity',d['parity_vs_oracle'])
print('stages',d['stages_ms'])
PY
