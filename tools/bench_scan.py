#!/usr/bin/env python
"""Sweep the scan kernel's tuning knobs (environment overrides read by launch_scan) on one GPU.

Every variant's output is compared with the first variant's (all are exact mod-q sums, so they must be identical).
Usage: python tools/bench_scan.py [--workload cfg2] [--queries 1] [--iters 20]
"""
import argparse
import itertools
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pir_b200 import sharded  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--queries", type=int, default=1)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--variants", default="")
    ap.add_argument("--shards", type=int, default=1, help="time the scan of shard 0 of this many row shards")
    args = ap.parse_args()
    params = bench.make_params(args.workload)
    ep = params.encryption_parameters
    N, mods = ep.poly_modulus_degree, ep.coeff_modulus
    k = len(mods) - 1
    dev = torch.device("cuda", 0)
    srv = sharded.ShardServer(params, device=0, shard_index=0, shard_count=args.shards)
    srv.db.fill_random(1)
    dimL = params.dimensions[-1]
    rng = np.random.default_rng(0)
    sv = bench.random_limbs(rng, mods[:k], (args.queries, dimL, 2), N)
    d_sv = sharded.to_device(sv, dev)
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():  # read a buffer larger than L2: evicts everything and leaves only CLEAN lines behind
        return flush.view(torch.int64).sum()
    peak, _ = bench.measured_peak()
    nbytes = srv.scan_bytes(args.queries)
    if args.variants:
        variants = [dict(kv.split("=") for kv in v.split(",")) for v in args.variants.split(";")]
    else:
        variants = [{"PIRB_SCAN_MODE": "0", "PIRB_MAC_MODE": "0", "PIRB_SCAN_R": "1", "PIRB_SCAN_U": "1"},
                    {"PIRB_SCAN_MODE": "0", "PIRB_MAC_MODE": "2", "PIRB_SCAN_R": "2", "PIRB_SCAN_U": "2"}]
        for r, g, st in [(2, 1, 8), (2, 2, 8), (4, 1, 8), (4, 2, 6), (4, 2, 8), (4, 4, 8), (8, 4, 4), (8, 4, 6),
                         (8, 4, 8), (8, 2, 6)]:
            for cps in (1, 2, 4):
                variants.append({"PIRB_SCAN_MODE": "1", "PIRB_MAC_MODE": "2", "PIRB_SCAN_R": str(r),
                                 "PIRB_SCAN_G": str(g), "PIRB_SCAN_U": str(st), "PIRB_SCAN_CTAS_PER_SM": str(cps)})
    ref = None
    srv.set_profiling(True)
    knobs = ["PIRB_SCAN_BATCH_V", "PIRB_B2_R", "PIRB_B2_QB", "PIRB_B2_RG", "PIRB_B2_U", "PIRB_MAC_MODE", "PIRB_SCAN_MINB", "PIRB_SCAN_G", "PIRB_SCAN_QB", "PIRB_SCAN_RB", "PIRB_SCAN_BATCH_MIN", "PIRB_SCAN_R", "PIRB_SCAN_U", "PIRB_SCAN_CTAS_PER_SM", "PIRB_SCAN_SPLIT",
             "PIRB_SCAN_MODE"]
    for v in variants:
        for kk in knobs:
            os.environ.pop(kk, None)
        os.environ.update(v)
        try:
            out = srv.scan(d_sv)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"variant": v, "error": str(e)}))
            continue
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
            same = True
        else:
            same = bool(torch.equal(ref, out))
        ts = []
        for _ in range(args.iters):
            flush_l2()
            srv.scan(d_sv, want_rows=False)
            ts.append(srv.last_scan_ms())  # CUDA events recorded around the kernel on its launching stream
        ts.sort()
        med = ts[len(ts) // 2]
        print(json.dumps({"variant": v, "ms_med": round(med, 4), "ms_min": round(ts[0], 4),
                          "GBps_med": round(nbytes / med / 1e6, 1), "frac": round(nbytes / med / 1e6 / peak, 3),
                          "identical": same}))


if __name__ == "__main__":
    main()
