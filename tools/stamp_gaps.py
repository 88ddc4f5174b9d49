#!/usr/bin/env python
"""Kernel-to-kernel gaps of the expansion levels (globaltimer at first CTA start / last CTA end per level)."""
import os, sys
os.environ["PIRB_DEBUG_STAMPS"] = "-2"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from pir_b200 import sharded, _lib
import pir_b200 as pb
items, size, d, n, bits, _ = bench.WORKLOADS["cfg2"]
params = pb.CreatePIRParameters(items, size, d, pb.GenerateEncryptionParams(n, bits))
_ep = params.encryption_parameters
srv = sharded.ShardServer(params, device=0); srv.db.fill_random(1)
q, elts, keys = bench.synth_inputs(_ep.poly_modulus_degree, _ep.coeff_modulus, list(params.dimensions), 1, 5)
srv.set_keys(pb.GaloisKeys(elts, keys.reshape(-1)))
dq = sharded.to_device(q, srv.device)
for _ in range(3): srv.answer(dq)
torch.cuda.synchronize()
n = 7 * 65536
buf = np.zeros(n, dtype=np.uint64)
_lib.lib().pirb_debug_stamps(srv.ctx.h, buf.ctypes.data_as(_lib.u64p), n)
prev_end = None
for j in range(7):
    ctas = (1 << j) * 6
    st = buf[j * 65536: j * 65536 + ctas * 8].reshape(-1, 8).astype(np.int64)
    start, end = st[:, 0].min(), st[:, 7].max()
    gap = (start - prev_end) if prev_end is not None else 0
    print("level %d: ctas %4d  in-kernel span %7.1f us  start spread %6.1f us  gap since previous level end %6.1f us" % (
        j, ctas, (end - start) / 1e3, (st[:, 0].max() - start) / 1e3, gap / 1e3))
    prev_end = end
