# round 2, call 1: regression tests, the new default bench (cfg4 x 8 queries), reference arm, sanitizer, scan ncu
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
(time python bench.py --steps 20 --warmup 5 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err); tail -c 3000 gpurun_out/r2c1_bench.json; tail -5 gpurun_out/r2c1_bench.err
(time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c1_ref.json 2> gpurun_out/r2c1_ref.err); tail -c 1500 gpurun_out/r2c1_ref.json; tail -5 gpurun_out/r2c1_ref.err
python bench.py --workload cfg2 --steps 20 --no-cpu-baseline > gpurun_out/r2c1_bench_cfg2.json 2> gpurun_out/r2c1_bench_cfg2.err; tail -c 1500 gpurun_out/r2c1_bench_cfg2.json
# compute-sanitizer: one cfg2 answer (cluster expansion kernel) and a 4-query batch (batched scan)
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/run_answer.py cfg2 1 1 > gpurun_out/r2c1_memcheck_cfg2.log 2>&1; tail -8 gpurun_out/r2c1_memcheck_cfg2.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/run_answer.py cfg2 1 1 > gpurun_out/r2c1_racecheck_cfg2.log 2>&1; tail -8 gpurun_out/r2c1_racecheck_cfg2.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/run_answer.py cfg2 4 1 > gpurun_out/r2c1_memcheck_cfg2_q4.log 2>&1; tail -8 gpurun_out/r2c1_memcheck_cfg2_q4.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/run_answer.py cfg2 4 1 > gpurun_out/r2c1_racecheck_cfg2_q4.log 2>&1; tail -8 gpurun_out/r2c1_racecheck_cfg2_q4.log
# ncu: single-query scan on the cfg4 database (traffic + full set), launch list of one cfg4 x 8 step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan -s 2 -c 1 -o gpurun_out/r2c1_scan_cfg4 -f python tools/run_answer.py cfg4 1 3 1 > gpurun_out/r2c1_ncu_scan.log 2>&1; tail -3 gpurun_out/r2c1_ncu_scan.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c1_launches_cfg4_q8.csv python tools/run_answer.py cfg4 8 2 > gpurun_out/r2c1_ncu_launches.log 2>&1; tail -3 gpurun_out/r2c1_ncu_launches.log
ls -la gpurun_out/r2c1_*
