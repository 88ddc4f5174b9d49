#!/usr/bin/env python
"""Print per-phase clock64 deltas of one expansion level of the cluster kernel (PIRB_DEBUG_STAMPS=<level>)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from pir_b200 import sharded, _lib
import pir_b200 as pb
level = int(os.environ.setdefault("PIRB_DEBUG_STAMPS", "0"))
items, size, d, n, bits, _ = bench.WORKLOADS["cfg2"]
params = pb.CreatePIRParameters(items, size, d, pb.GenerateEncryptionParams(n, bits))
_ep = params.encryption_parameters
srv = sharded.ShardServer(params, device=0); srv.db.fill_random(1)
q, elts, keys = bench.synth_inputs(_ep.poly_modulus_degree, _ep.coeff_modulus, list(params.dimensions), 1, 5)
srv.set_keys(pb.GaloisKeys(elts, keys.reshape(-1)))
dq = sharded.to_device(q, srv.device)
for _ in range(3): srv.answer(dq)
torch.cuda.synchronize()
n = (1 << level) * 6 * 8
buf = np.zeros(n, dtype=np.uint64)
_lib.lib().pirb_debug_stamps(srv.ctx.h, buf.ctypes.data_as(_lib.u64p), n)
st = buf.reshape(-1, 8).astype(np.int64)
names = {1: "fwd NTT", 2: "canon", 3: "cluster.sync", 4: "MAC", 5: "inv NTT"}
if os.environ.get("PIRB_STAMP_CLOCK"):
    names = {0: "prologue+gather", **names, 6: "fin+sync+phase3"}
for i, nm in names.items():
    d = st[:, i + 1] - st[:, i]
    print("%-14s mean %8.0f  min %8d  max %8d cycles" % (nm, d.mean(), d.min(), d.max()))
span = st[:, 7] - st[:, 0]  # globaltimer ns
mid = (st[:, 6] - st[:, 1]).mean()
print("stamped phases total %.0f cycles; whole CTA %.1f us mean (globaltimer), kernel span %.1f us" % (
    mid, span.mean() / 1e3, (st[:, 7].max() - st[:, 0].min()) / 1e3))
