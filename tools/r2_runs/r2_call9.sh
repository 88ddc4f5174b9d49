set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "engine_variant or sixty_bit or full_size or peer_memory or batched_scan" 2>&1 | tail -12
python tools/shard_probe.py cfg4 8 8 2>&1 | tail -2
python tools/shard_probe.py cfg4 1 8 2>&1 | tail -2
