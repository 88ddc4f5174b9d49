run() {  # name, n, port, extra args...
  name=$1; n=$2; port=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/r2c11_$name.json 2> gpurun_out/r2c11_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c11_$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "parity", d["parity_vs_oracle"], d.get("stages_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2c11_$name.err").read()[-1500:])
PY
}
run n8 8 29711 --no-parity
PIRB_PUSH_CTAS=32 run n8_push32 8 29712 --no-parity
PIRB_PUSH_CTAS=128 run n8_push128 8 29713 --no-parity
