"""Multi-process tests of the row-sharded path.

* CPU (gloo, world_size 2): the sharding arithmetic the product uses (pir_b200.sharded.shard_rows + selection-vector
  slicing + mod-q add of partial replies) reproduces the unsharded oracle answer.
* GPU (nccl, needs >= 2 devices, skipped otherwise): tests/helpers/dist_parity.py under torchrun.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import binding as ob
    from oracle import client as oc
    from pir_b200.sharded import shard_rows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = oc.create_pir_parameters(82, 0, 2, 4096, 20)     # dims [10, 9], short last row
        cl = oc.HarnessClient(p, seed=5)
        rng = np.random.default_rng(42)
        items = [rng.integers(0, 256, p.bytes_per_item, dtype=np.uint8).tobytes() for _ in range(p.num_items)]
        db = oc.db_to_ntt(cl.orc, oc.encode_string_db(p, items))
        q = cl.create_query(42)
        sv = cl.orc.expand(q, p.dim_sum, cl.elts, cl.galois)          # replicated expansion
        d0, d1 = p.dimensions
        lo, hi = shard_rows(d0, world)[rank]
        # this rank's rows: sub-database + the matching slice of the first-dimension selection vector
        sub_db = db[lo * d1:min(hi * d1, p.num_pt)]
        if hi > lo and len(sub_db):
            sub_sv = np.concatenate([sv[lo:hi], sv[d0:]])
            part, _ = cl.orc.db_multiply(sub_db, [hi - lo, d1], sub_sv)   # coefficient form; the sum is linear
        else:
            part = np.zeros((2 * cl.orc.ER, 2, cl.orc.k, cl.orc.N), dtype=np.uint64)
        t = torch.from_numpy(part.view(np.int64).copy())
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        mods = np.array(cl.orc.moduli[:cl.orc.k], dtype=np.uint64)[None, None, :, None]
        total = np.zeros_like(part)
        for g in gathered:                                              # mod-q add, never a plain integer sum
            total = (total + g.numpy().view(np.uint64)) % mods
        want = cl.orc.process_query(db, p.dimensions, cl.elts, cl.galois, q)
        ok = np.array_equal(total, want) and cl.process_response_strings([42], [total])[0] == items[42]
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_row_sharding_math_gloo_world2():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, ret)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(300)
        assert pr.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


def _gloo_worker_d1(rank, world, port, ret):
    """d=1 over two query ciphertexts: a shard expands only the trees that cover its own plaintexts (what
    run_answer does for a one-dimensional shard), scans them, and the NTT-form partials add up mod q."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import client as oc
    from pir_b200.sharded import shard_rows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_items, N = 4200, 4096
        p = oc.create_pir_parameters(n_items, 0, 1, N, 20)       # dims [4200]: trees of 4096 and 104 items
        cl = oc.HarnessClient(p, seed=8)
        rng = np.random.default_rng(3)
        # a random NTT-form database is enough here: the property is linearity over the plaintext index
        mods = [int(q) for q in cl.orc.moduli[:cl.orc.k]]
        db = np.ascontiguousarray(np.stack([rng.integers(0, mods[j], (n_items, N), dtype=np.uint64)
                                            for j in range(cl.orc.k)], axis=1))
        idx = 4150                                                 # lives in the second tree
        q = cl.create_query(idx)                                   # [2][2][k][N]
        lo, hi = shard_rows(n_items, world)[rank]
        t_first, t_last = lo // N, (hi - 1) // N
        sv = {}
        for t in range(t_first, t_last + 1):                       # only this shard's trees
            items = min(N, n_items - t * N)
            out = cl.orc.expand(q[t], items, cl.elts, cl.galois, single=True)
            for i in range(items):
                sv[t * N + i] = out[i]
        assert all(i in sv for i in range(lo, hi))
        sub_sv = np.stack([cl.orc.ct_to_ntt(sv[i]) for i in range(lo, hi)])
        part = cl.orc.scan_row(db[lo:hi], sub_sv)                  # NTT-form partial reply of this shard
        t = torch.from_numpy(part.view(np.int64).copy())
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        m = np.array(mods, dtype=np.uint64)[None, :, None]
        total = np.zeros_like(part)
        for g in gathered:
            total = (total + g.numpy().view(np.uint64)) % m
        got = cl.orc.ct_from_ntt(total)
        want = cl.orc.process_query(db, p.dimensions, cl.elts, cl.galois, q)[0]
        ret[rank] = bool(np.array_equal(got, want))
    finally:
        dist.destroy_process_group()


def test_one_dimensional_tree_sharding_math_gloo_world2():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_gloo_worker_d1, args=(r, world, port, ret)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(600)
        assert pr.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_shard_rows_cover_and_match_library_split():
    sys.path.insert(0, ROOT)
    from pir_b200.sharded import shard_rows
    for d0 in (1, 2, 3, 10, 41, 333, 665):
        for s in (1, 2, 3, 4, 8):
            r = shard_rows(d0, s)
            assert r[0][0] == 0 and r[-1][1] == d0
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(0 <= hi - lo <= -(-d0 // s) for lo, hi in r)


@pytest.mark.gpu
def test_distributed_parity_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "helpers", "dist_parity.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "DIST_PARITY_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
