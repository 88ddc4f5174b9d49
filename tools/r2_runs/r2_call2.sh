# round 2, call 2 (2 GPUs): NVLink exchange flow — single-GPU multi-rank test, torchrun parity, bench N=2 both transports
set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer_memory_exchange" 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/helpers/dist_parity.py 2>&1 | tail -15
for ex in nvlink nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --exchange $ex > gpurun_out/r2c2_bench_n2_$ex.json 2> gpurun_out/r2c2_bench_n2_$ex.err
  tail -c 2500 gpurun_out/r2c2_bench_n2_$ex.json; tail -5 gpurun_out/r2c2_bench_n2_$ex.err
done
