timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "substitute or expansion or process_request or custom_moduli or end_to_end or multi_dim" 2>&1 | tail -4
PIRB_STAMP_CLOCK=1 PIRB_DEBUG_STAMPS=0 python tools/stamp_probe.py 2>&1 | tail -9
python tools/stamp_gaps.py 2>&1 | tail -8
python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_tw.err | tail -1 > gpurun_out/bench_tw.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_tw.json').read()); print('cfg2 N=1', d['value'], d['ms_per_step'], d['p50_latency_ms'], d['e2e']['value'], d['stages_ms'], d['clocks'])"
python bench.py --steps 10 --warmup 3 --queries-per-gpu 16 --no-cpu-baseline 2>> gpurun_out/bench_tw.err | tail -1 > gpurun_out/bench_tw_q16.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_tw_q16.json').read()); print('cfg2 Q16', d['value'], d['ms_per_step'], d['stages_ms'])"
