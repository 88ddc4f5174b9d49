"""Eager answers of `nq` queries of a bench workload (for ncu / compute-sanitizer: no graphs, few launches).
usage: python tools/run_answer.py [workload=cfg2] [nq=1] [reps=2] [scan_only=0]"""
import os, sys
os.environ["PIRB_GRAPHS"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from pir_b200 import sharded
import pir_b200 as pb
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
scan_only = int(sys.argv[4]) if len(sys.argv) > 4 else 0
items, size, d, n, bits, _ = bench.WORKLOADS[wl]
params = pb.CreatePIRParameters(items, size, d, pb.GenerateEncryptionParams(n, bits))
ep = params.encryption_parameters
N, mods, dims = ep.poly_modulus_degree, ep.coeff_modulus, list(params.dimensions)
srv = sharded.ShardServer(params, device=0); srv.db.fill_random(1)
if scan_only:
    sv = sharded.to_device(bench.random_limbs(np.random.default_rng(5), mods[:-1], (nq, dims[-1], 2), N), srv.device)
    for _ in range(reps):
        srv.scan(sv, want_rows=False)
else:
    q, elts, keys = bench.synth_inputs(N, mods, dims, nq, 5)
    srv.set_keys(pb.GaloisKeys(elts, keys.reshape(-1)))
    dq = sharded.to_device(q, srv.device)
    for _ in range(reps):
        srv.answer(dq)
torch.cuda.synchronize()
print("run_answer ok", wl, nq, reps, scan_only)
