timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ks_level_cluster -s 13 -c 1 -o gpurun_out/ncu_l6_q1 -f python tools/ncu_level6.py 1 > gpurun_out/ncu_l6_q1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ks_level_cluster -s 13 -c 1 -o gpurun_out/ncu_l6_q16 -f python tools/ncu_level6.py 16 > gpurun_out/ncu_l6_q16.log 2>&1
python tools/stamp_gaps.py 2>&1 | tail -12
ls -la gpurun_out/ncu_l6*.ncu-rep
