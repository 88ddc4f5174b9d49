run() {  # name, n, port, extra args...
  name=$1; n=$2; port=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/r2c21_$name.json 2> gpurun_out/r2c21_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c21_$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "parity", d["parity_vs_oracle"], {k: round(v, 3) for k, v in d.get("stages_ms").items()})
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2c21_$name.err").read()[-1500:])
PY
}
PIRB_DIST_SUB=1 run n2_sub1 2 29751 --no-parity
run n2 2 29752 --no-parity
