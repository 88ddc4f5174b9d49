# final 8-GPU pass: config 2 at 2/4/8 GPUs (the driver's scaling bench), config 4 with 64 concurrent queries, config 5a
run() {  # name, visible devices, n, port, extra args...
  name=$1; vis=$2; n=$3; port=$4; shift 4
  CUDA_VISIBLE_DEVICES=$vis timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/run8_$name.err | tail -1 > gpurun_out/run8_$name.json
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/run8_$name.json").read())
    print("$name", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d.get("stages_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/run8_$name.err").read()[-1500:])
PY
}
ALL=0,1,2,3,4,5,6,7
run cfg2_n8 $ALL 8 29801 --steps 20
run cfg2_n8_q8 $ALL 8 29802 --steps 10 --queries-per-gpu 8
run cfg4_n8_q8 $ALL 8 29803 --steps 5 --workload cfg4 --queries-per-gpu 8
run cfg5a_n8 $ALL 8 29804 --steps 3 --workload cfg5a
run cfg2_n4 0,1,2,3 4 29805 --steps 20 &
run cfg2_n2 4,5 2 29806 --steps 20 &
wait
