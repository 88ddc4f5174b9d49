# round 2, call 3 (8 GPUs): group test on one GPU, then the scaling points of the default workload
set -x
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer_memory_exchange" 2>&1 | tail -8
run() {  # name, n, port, extra args...
  name=$1; n=$2; port=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/r2c3_$name.json 2> gpurun_out/r2c3_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c3_$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "parity", d["parity_vs_oracle"], d.get("stages_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2c3_$name.err").read()[-1500:])
PY
}
run n8_nvlink 8 29601
run n8_nccl 8 29602 --exchange nccl --no-parity
run n4_nvlink 4 29603 --no-parity
run n8_nvlink_cfg2 8 29604 --workload cfg2 --steps 20 --no-parity
