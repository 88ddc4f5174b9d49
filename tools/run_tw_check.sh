timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "process_request or multi_dim or end_to_end or row_sharded or graph or custom_moduli or unused" 2>&1 | tail -4
python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_tw.err | tail -1 > gpurun_out/bench_tw.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_tw.json').read()); print('cfg2 N=1', d['value'], d['ms_per_step'], d['p50_latency_ms'], d['e2e']['value'], d['stages_ms'], d['clocks'])"
python bench.py --steps 10 --warmup 3 --workload cfg4 --no-cpu-baseline 2>> gpurun_out/bench_tw.err | tail -1 > gpurun_out/bench_tw_cfg4.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_tw_cfg4.json').read()); print('cfg4', d['value'], d['ms_per_step'], d['stages_ms'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan -s 1 -c 1 -o gpurun_out/final_scan_cfg2 -f python tools/ncu_level6.py 1 > gpurun_out/final_ncu_scan.log 2>&1; tail -2 gpurun_out/final_ncu_scan.log
