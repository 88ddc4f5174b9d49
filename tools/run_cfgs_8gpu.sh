# BASELINE configs 2-4 on one 8-GPU box: config 3 (N=8192) row-sharded over 2/4/8 GPUs, config 4 (2^22 x 256 B) with
# 8 queries per GPU (64 concurrent queries sharing one scan at 8 GPUs) at 1/2/4/8 GPUs, config 5b at 8 GPUs.
# Usage: bash tools/run_cfgs_8gpu.sh   (writes gpurun_out/cfgs8_*.json)
run() {  # name, visible devices, n, port, extra args...
  name=$1; vis=$2; n=$3; port=$4; shift 4
  if [ "$n" = 1 ]; then
    CUDA_VISIBLE_DEVICES=$vis timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/cfgs8_$name.err | tail -1 > gpurun_out/cfgs8_$name.json
  else
    CUDA_VISIBLE_DEVICES=$vis timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/cfgs8_$name.err | tail -1 > gpurun_out/cfgs8_$name.json
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/cfgs8_$name.json").read())
    print("$name", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "scan frac", round(d["roofline"]["frac"], 3), d.get("stages_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/cfgs8_$name.err").read()[-1500:])
PY
}
ALL=0,1,2,3,4,5,6,7
run cfg4_n8_q8 $ALL 8 29601 --workload cfg4 --queries-per-gpu 8
run cfg3_n8 $ALL 8 29603 --workload cfg3
run cfg5b_n8 $ALL 8 29604 --workload cfg5b
# smaller GPU counts side by side on disjoint GPUs
run cfg3_n4 0,1,2,3 4 29605 --workload cfg3 &
run cfg3_n2 4,5 2 29606 --workload cfg3 &
run cfg4_n1_q8 6 1 0 --workload cfg4 --queries-per-gpu 8 &
wait
run cfg4_n4_q8 0,1,2,3 4 29607 --workload cfg4 --queries-per-gpu 8 &
run cfg4_n2_q8 4,5 2 29608 --workload cfg4 --queries-per-gpu 8 &
run cfg4_n1_q1 6 1 0 --workload cfg4 &
wait
