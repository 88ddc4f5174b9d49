set -x
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ct_multiply or golden or cpp_shim or graph" 2>&1 | tail -3
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
