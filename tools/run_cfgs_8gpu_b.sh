# second 8-GPU pass: config 5a (d=1, shards expand only their trees), config 4 with 64 concurrent queries (new batched
# scan + ownership-aware exchange), config 3 at 8 and 2 GPUs, distributed parity on 8 ranks
run() {  # name, visible devices, n, port, extra args...
  name=$1; vis=$2; n=$3; port=$4; shift 4
  CUDA_VISIBLE_DEVICES=$vis timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/cfgs8b_$name.err | tail -1 > gpurun_out/cfgs8b_$name.json
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/cfgs8b_$name.json").read())
    print("$name", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "scan frac", round(d["roofline"]["frac"], 3), d.get("stages_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/cfgs8b_$name.err").read()[-1500:])
PY
}
ALL=0,1,2,3,4,5,6,7
run cfg5a_n8 $ALL 8 29701 --workload cfg5a --steps 3
run cfg4_n8_q8 $ALL 8 29702 --workload cfg4 --queries-per-gpu 8
run cfg3_n8 $ALL 8 29703 --workload cfg3
run cfg3_n2 0,1 2 29704 --workload cfg3 &
run cfg5a_n2 2,3 2 29705 --workload cfg5a --steps 3 &
run cfg5b_n4 4,5,6,7 4 29706 --workload cfg5b &
wait
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29707 tests/helpers/dist_parity.py 2>&1 | tail -6
