# scaling run: bench.py at N GPUs (argument)
N=$1
for fg in 0 1; do
PIRB_FULL_GATHER=$fg python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_scale_n${N}_fg$fg.json
python -c "import json; d=json.loads(open('gpurun_out/bench_scale_n${N}_fg$fg.json').read()); print('N=$N full_gather=$fg', d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 tests/helpers/dist_parity.py 2>&1 | tail -5
