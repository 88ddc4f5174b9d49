timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -6
for p in 1 0; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 3 --p2p $p 2>/dev/null | tail -1 > gpurun_out/bench_n2_p2p$p.json
  python -c "import json; d=json.loads(open('gpurun_out/bench_n2_p2p$p.json').read()); print('p2p=$p', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
