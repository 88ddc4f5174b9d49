# full GPU test suite + default bench (N=1) + batch-16 bench
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_latest.err | tail -1 > gpurun_out/bench_latest.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_latest.json').read()); print('cfg2 N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['expansion']['frac'], d['stages_ms'], d['cpu_baseline']['value'] if d['cpu_baseline'] else None, d['parity_vs_oracle'])"
python bench.py --steps 10 --warmup 3 --queries-per-gpu 16 --no-cpu-baseline 2>> gpurun_out/bench_latest.err | tail -1 > gpurun_out/bench_q16.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_q16.json').read()); print('cfg2 Q16', d['value'], d['ms_per_step'], d['stages_ms'])"
python bench.py --steps 5 --warmup 3 --workload cfg4 --queries-per-gpu 16 --no-cpu-baseline 2>> gpurun_out/bench_latest.err | tail -1 > gpurun_out/bench_cfg4_q16.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg4_q16.json').read()); print('cfg4 Q16', d['value'], d['ms_per_step'], d['stages_ms'])"
