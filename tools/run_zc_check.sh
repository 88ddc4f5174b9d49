timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "process_request or cpp_shim" 2>&1 | tail -4
for z in 0 1; do
PIRB_ZERO_COPY=$z python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('zero_copy=$z', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e']['p50_latency_ms'])"
PIRB_ZERO_COPY=$z python bench.py --steps 10 --warmup 3 --workload cfg1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg1 zero_copy=$z', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e']['p50_latency_ms'])"
PIRB_ZERO_COPY=$z python bench.py --steps 10 --warmup 3 --workload cfg3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 zero_copy=$z', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e']['p50_latency_ms'])"
done
