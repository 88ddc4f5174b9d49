set -x
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ct_multiply" 2>&1 | tail -25
