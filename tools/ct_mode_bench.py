"""Ciphertext-multiplication mode (use_ciphertext_multiplication, database.cpp:202-211) timed through the public Python
API on a bench workload: ProcessRequest of `nq` queries with relinearization keys, host buffers in and out.
Synthetic uniform limbs (every kernel is data-independent).  Not the headline metric — this mode is off by default in the
reference and outside north_star; the number documents that the mode runs at the workload's size.
usage: python tools/ct_mode_bench.py [workload=cfg4] [nq=8] [reps=5]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import pir_b200 as pb

# (items, bytes per item, dimensions, N, plain bits) — bench.py's workloads (not imported: it pulls in torch)
WORKLOADS = {"cfg2": (1 << 16, 288, 2, 4096, 24), "cfg3": (1 << 20, 1024, 2, 8192, 20), "cfg4": (1 << 22, 256, 2, 4096, 20)}


def random_limbs(rng, moduli, shape_prefix, N):
    out = np.empty(tuple(shape_prefix) + (len(moduli), N), dtype=np.uint64)
    for j, q in enumerate(moduli):
        out[..., j, :] = rng.integers(0, int(q), size=tuple(shape_prefix) + (N,), dtype=np.uint64)
    return out


wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
items, size, d, n, bits = WORKLOADS[wl]
out = {"workload": wl, "queries_per_request": nq}
for ct_mult in (True, False):
    params = pb.CreatePIRParameters(items, size, d, pb.GenerateEncryptionParams(n, bits), ct_mult)
    ep = params.encryption_parameters
    N, mods, dims = ep.poly_modulus_degree, ep.coeff_modulus, list(params.dimensions)
    db = pb.PIRDatabase(params)
    db.fill_random(2024)
    server = pb.PIRServer.Create(db, params)
    k = len(mods) - 1
    rng = np.random.default_rng(5)
    q = random_limbs(rng, mods[:k], (nq, sum(dims) // N + 1, 2), N)
    elts = [(N >> i) + 1 for i in range(N.bit_length() - 1)]
    keys = random_limbs(rng, mods, (len(elts), k, 2), N)
    relin = random_limbs(rng, mods, (k, 2), N).reshape(-1) if ct_mult else None
    req = pb.Request(query=[q[i] for i in range(nq)], galois_keys=pb.GaloisKeys(elts, keys.reshape(-1)), relin_keys=relin)
    for _ in range(2):
        resp = server.ProcessRequest(req)
    lat = []
    for _ in range(reps):
        t0 = time.perf_counter()
        resp = server.ProcessRequest(req)
        lat.append(time.perf_counter() - t0)
    ms = float(np.median(lat)) * 1e3
    name = "ct_multiplication" if ct_mult else "reencoder"
    out[name] = {"ms_per_request": round(ms, 3), "queries_per_s": round(nq / ms * 1e3, 2),
                 "reply_limbs_per_query": int(resp.reply[0].size)}
    del server, db
print(json.dumps(out))
