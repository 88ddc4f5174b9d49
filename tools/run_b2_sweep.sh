# sweep of the second-generation batched scan against the first (same outputs required)
V="PIRB_SCAN_BATCH_V=1"
for shape in "2,2,4,2" "2,2,2,2" "2,2,4,4" "2,4,4,1" "4,1,4,4"; do
  IFS=, read r q g u <<< "$shape"
  V="$V;PIRB_SCAN_BATCH_V=2,PIRB_B2_R=$r,PIRB_B2_QB=$q,PIRB_B2_RG=$g,PIRB_B2_U=$u"
done
python tools/bench_scan.py --workload cfg4 --queries 16 --iters 5 --variants "$V" > gpurun_out/b2_cfg4_q16.jsonl 2> gpurun_out/b2_cfg4_q16.err
python tools/bench_scan.py --workload cfg4 --queries 64 --shards 8 --iters 5 --variants "$V" > gpurun_out/b2_cfg4_q64_s8.jsonl 2> gpurun_out/b2_cfg4_q64_s8.err
python tools/bench_scan.py --workload cfg2 --queries 16 --iters 10 --variants "$V" > gpurun_out/b2_cfg2_q16.jsonl 2> gpurun_out/b2_cfg2_q16.err
for f in b2_cfg4_q16 b2_cfg4_q64_s8 b2_cfg2_q16; do echo == $f; python - <<PY
import json
for l in open("gpurun_out/$f.jsonl"):
    try: d=json.loads(l)
    except Exception: continue
    v=d.get("variant",{}); print(v.get("PIRB_SCAN_BATCH_V"), v.get("PIRB_B2_R"), v.get("PIRB_B2_QB"), v.get("PIRB_B2_RG"), v.get("PIRB_B2_U"), d.get("ms_med"), d.get("identical"), d.get("error"))
PY
tail -3 gpurun_out/$f.err; done
