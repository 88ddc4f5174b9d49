"""Tensor-core batched scan vs the CUDA-core scans (and the oracle on sampled rows) on one GPU.
usage: python tools/tc_check.py [small|cfg2|cfg4|cfg3] [n_queries]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import pir_b200 as pb
from pir_b200 import sharded, _lib
from oracle import binding as ob

which = sys.argv[1] if len(sys.argv) > 1 else "small"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if which == "small":
    ep = pb.GenerateEncryptionParams(4096, 20); params = pb.CreatePIRParameters(300, 0, 2, ep)
elif which == "small8192":
    ep = pb.GenerateEncryptionParams(8192, 20); params = pb.CreatePIRParameters(120, 0, 2, ep)
else:
    items, size, d, n, bits, _ = bench.WORKLOADS[which]
    params = pb.CreatePIRParameters(items, size, d, pb.GenerateEncryptionParams(n, bits))
ep = params.encryption_parameters
N, mods, dims = ep.poly_modulus_degree, ep.coeff_modulus, list(params.dimensions)
k = len(mods) - 1
dimL = dims[-1]
rng = np.random.default_rng(3)
sv = bench.random_limbs(rng, mods[:k], (nq, dimL, 2), N)
d_sv = sharded.to_device(sv, "cuda:0")

def run(tc_min):
    os.environ["PIRB_TC_MIN"] = str(tc_min)
    srv = sharded.ShardServer(params, device=0); srv.db.fill_random(11)
    out = srv.scan(d_sv); torch.cuda.synchronize()
    rc = _lib.lib().pirb_sync(srv.ctx.h)
    if rc: print("pirb_sync:", _lib.last_error())
    ts = []
    srv.set_profiling(True)
    for _ in range(3):
        srv.scan(d_sv, want_rows=False); torch.cuda.synchronize(); ts.append(srv.last_scan_ms())
    srv.set_profiling(False)
    return srv, out.cpu().numpy().view(np.uint64), ts

srv_tc, got, t_tc = run(4)
print("tc scan ms", [round(x, 3) for x in t_tc], flush=True)
srv_ref, want, t_ref = run(0)
print("cuda-core scan ms", [round(x, 3) for x in t_ref], flush=True)
eq = np.array_equal(got, want)
print("TC == CUDA-core batched scan:", eq)
if not eq:
    bad = np.argwhere(got != want)
    print("mismatches", len(bad), "of", got.size, "first", bad[:5], got[tuple(bad[0])], want[tuple(bad[0])])
orc = ob.Oracle(N, list(mods), ep.plain_modulus)
n_rows = got.shape[1]
ok = True
for i, r in [(0, 0), (nq - 1, n_rows - 1), (nq // 2, n_rows // 2)]:
    first = r * dimL; cnt = min(dimL, params.num_pt - first)
    w = orc.scan_row(srv_tc.db.read_ntt(first, cnt), sv[i, :cnt])
    ok &= bool(np.array_equal(got[i, r], w))
print("TC == oracle on sampled rows:", ok)
print("TC_CHECK_OK" if (eq and ok) else "TC_CHECK_FAIL")
