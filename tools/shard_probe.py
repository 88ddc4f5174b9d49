"""One rank's share of a sharded step, alone on one GPU: the multiply of `nq` queries against shard 0 of `world`
(stage times), for sizing the consumer side of the multi-GPU flow.  usage: shard_probe.py [workload] [world] [nq]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import pir_b200 as pb
from pir_b200 import sharded
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 8
items, size, d, n, bits, _ = bench.WORKLOADS[wl]
params = pb.CreatePIRParameters(items, size, d, pb.GenerateEncryptionParams(n, bits))
ep = params.encryption_parameters
N, mods, dims = ep.poly_modulus_degree, ep.coeff_modulus, list(params.dimensions)
k = len(mods) - 1
srv = sharded.ShardServer(params, device=0, shard_index=0, shard_count=world); srv.db.fill_random(1)
rng = np.random.default_rng(5)
sv = sharded.to_device(bench.random_limbs(rng, mods[:k], (nq, sum(dims), 2), N), srv.device)
for _ in range(3):
    srv.multiply_partial(sv)
torch.cuda.synchronize()
srv.set_profiling(True)
acc = {}
for _ in range(5):
    srv.multiply_partial(sv); torch.cuda.synchronize()
    for nm, v in srv.stage_ms().items():
        acc.setdefault(nm, []).append(v)
srv.set_profiling(False)
print(wl, "world", world, "nq", nq, "pt", srv.pt_count, {nm: round(sum(v) / len(v), 4) for nm, v in acc.items()})
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(10):
    srv.multiply_partial(sv)
ev1.record(); torch.cuda.synchronize()
print("graph-replayed multiply_partial ms:", round(ev0.elapsed_time(ev1) / 10, 4))
