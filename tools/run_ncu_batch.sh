# ncu --set full of the second-generation batched scan on config 4 with 16 queries
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_batch -s 1 -c 1 -o gpurun_out/ncu_batch_v2b -f \
  python tools/bench_scan.py --workload cfg4 --queries 16 --iters 1 --variants "PIRB_SCAN_BATCH_V=2,PIRB_B2_R=2,PIRB_B2_QB=2,PIRB_B2_RG=4,PIRB_B2_U=2" > gpurun_out/ncu_batch_v2b.log 2>&1
ls -la gpurun_out/*.ncu-rep
