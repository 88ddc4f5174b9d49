set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer_memory_exchange or batched_scan" 2>&1 | tail -6
timeout 300 python tools/tc_check.py small 8 2>&1 | tail -3
