# scaling run: bench.py at N GPUs (argument), both exchange variants are not needed: default p2p
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_scale_n$N.json
python -c "import json; d=json.loads(open('gpurun_out/bench_scale_n$N.json').read()); print('N=$N', d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 10 --warmup 3 --queries-per-gpu 8 2>/dev/null | tail -1 > gpurun_out/bench_scale_n${N}_q8.json
python -c "import json; d=json.loads(open('gpurun_out/bench_scale_n${N}_q8.json').read()); print('N=$N q8', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 tests/helpers/dist_parity.py 2>&1 | tail -5
