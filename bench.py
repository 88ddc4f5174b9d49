#!/usr/bin/env python
"""bench.py — PIR server answer path on B200 (BASELINE.json metric: queries/sec + p50 latency + DB-scan GB/s).

Workload (config.workload): BASELINE.json configs[1] — d=2, 2^16 elements x 288 B, BFV N=4096, 24-bit plain modulus
(pir/cpp/benchmark.cpp:17-23 shape: 1639 plaintexts, dims [41,40], 127 key switches and 8 reply ciphertexts per
query).  A step = one pass of the hot path (oblivious expansion + database multiply) over one batch of queries:
  N=1  : one query on one GPU (the configuration the metric is quoted on);
  N>1  : one query per GPU per step (weak scaling).  The database is row-sharded across the N GPUs; each rank
         expands its own query, NTT-form selection vectors are all-gathered (NCCL), every rank multiplies all N
         queries against its rows, partial replies are all-gathered and added mod q (fused into the final inverse NTT).
Synthetic data: uniform limbs in [0,q_j) for the NTT-form database, query ciphertexts and Galois keys (every kernel
on the path is data-independent).  `value` is device-resident throughput (inputs already in HBM); `e2e` goes through
the reference-facing call (PIRServer.ProcessRequest -> C ABI pirb_answer) with pinned HOST buffers, H2D/D2H inside.

--impl reference times the CPU restatement of the reference path (oracle/, SEAL-3.5.6-equivalent algorithms; SEAL itself
cannot be built here) on all host threads for the same workload.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (num_items, bytes_per_item, dims, N, plain_bits, description)
    "cfg2": (1 << 16, 288, 2, 4096, 24, "d=2, 2^16 x 288 B, N=4096, t 24-bit (BASELINE configs[1])"),
    "cfg1": (4096, 64, 1, 4096, 20, "d=1, 4096 x 64 B, N=4096 (BASELINE configs[0])"),
    "cfg4": (1 << 22, 256, 2, 4096, 20, "d=2, 2^22 x 256 B, N=4096 (BASELINE configs[3] database)"),
    "cfg3": (1 << 20, 1024, 2, 8192, 20, "d=2, 2^20 x 1 KiB, N=8192 (BASELINE configs[2] database)"),
    "cfg5a": (1 << 24, 256, 1, 4096, 20, "d=1, 2^24 x 256 B, N=4096 (BASELINE configs[4], d=1 arm)"),
    "cfg5b": (1 << 24, 256, 2, 4096, 20, "d=2, 2^24 x 256 B, N=4096 (BASELINE configs[4], d=2 arm)"),
}


def make_params(name):
    import pir_b200 as pb
    items, size, d, n, bits, _ = WORKLOADS[name]
    return pb.CreatePIRParameters(items, size, d, pb.GenerateEncryptionParams(n, bits))


def random_limbs(rng, moduli, shape_prefix, N):
    """uniform [*shape_prefix][len(moduli)][N] with limb j < moduli[j]"""
    cols = [rng.integers(0, int(q), size=tuple(shape_prefix) + (N,), dtype=np.uint64) for q in moduli]
    return np.ascontiguousarray(np.stack(cols, axis=len(shape_prefix)))


def synth_inputs(params, n_queries, seed):
    ep = params.encryption_parameters
    N, mods = ep.poly_modulus_degree, ep.coeff_modulus
    k = len(mods) - 1
    rng = np.random.default_rng(seed)
    n_ct = sum(params.dimensions) // N + 1
    queries = random_limbs(rng, mods[:k], (n_queries, n_ct, 2), N)                 # [Q][n_ct][2][k][N]
    elts = [(N >> i) + 1 for i in range(N.bit_length() - 1)]
    keys = random_limbs(rng, mods, (len(elts), k, 2), N)                           # [n][k][2][k+1][N]
    return queries, elts, keys


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region, from the enqueuing thread itself: all timed steps
    are enqueued first (the host runs ahead of the GPU), then NVML is polled until the last step's end event has
    completed.  NVML queries take a driver lock that can stall kernel launches for milliseconds, so they must not run
    concurrently with the enqueue loop (a sampling thread did exactly that and added 4-9 ms to single steps)."""

    def __init__(self, gpu_index):
        # NVML enumerates physical devices; honour a CUDA_VISIBLE_DEVICES remap of numeric indices
        vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip().isdigit()]
        self.idx = int(vis[gpu_index]) if gpu_index < len(vis) else gpu_index
        self.sm, self.mx, self.reasons = [], None, set()
        self.nv = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                          "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                          "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                          "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            self.get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self._sample()      # resolves every NVML entry point now, outside the timed region
            self.sm = []
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.reasons.add("nvml unavailable: %s" % type(e).__name__)

    def _sample(self):
        self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
        r = self.get_reasons(self.h)
        for nm, bit in self.names.items():
            if r & bit:
                self.reasons.add(nm)

    def sample_until(self, done_event):
        """Poll while the GPU is still working through the enqueued timed steps (at least one sample)."""
        if self.nv is None:
            return
        try:
            while True:
                self._sample()
                if done_event.query():
                    break
                time.sleep(0.0005)
        except Exception as e:  # noqa: BLE001
            self.reasons.add("nvml unavailable: %s" % type(e).__name__)

    def result(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# =====================================================================================================
def run_reference(args):
    """Reference arm: the CPU restatement of the reference's path on all host threads (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from oracle import binding as ob
    params = make_params(args.workload)
    ep = params.encryption_parameters
    N, mods = ep.poly_modulus_degree, ep.coeff_modulus
    k = len(mods) - 1
    cores = os.cpu_count() or 1
    orc = ob.Oracle(N, mods, ep.plain_modulus)
    rng = np.random.default_rng(1234)
    # bounded sample of the workload when the database is large: keep the first rows (whole rows of the hypercube)
    max_pt = min(params.num_pt, max(params.dimensions[-1], (2 << 30) // (k * N * 8)))
    dims = list(params.dimensions)
    sample = "full workload"
    if max_pt < params.num_pt:
        rows = max(1, max_pt // dims[-1])
        max_pt = rows * dims[-1]
        sample = "first %d of %d rows of the database (expansion in full); scaled to the full scan by rows" % (
            rows, -(-params.num_pt // dims[-1]))
    db = random_limbs(rng, mods[:k], (max_pt,), N)
    queries, elts, keys = synth_inputs(params, cores, 99)
    keys = keys.reshape(-1)

    def one(i):
        return orc.process_query(db, dims, elts, keys, queries[i % len(queries)])

    pool = ThreadPoolExecutor(cores)
    t_one0 = time.perf_counter(); one(0); t_one = time.perf_counter() - t_one0
    steps = args.steps
    # bound the whole run to a few minutes
    while steps > 1 and (steps + args.warmup) * t_one * 1.3 > 240:
        steps -= 1
    for _ in range(max(1, min(args.warmup, 2))):
        list(pool.map(one, range(cores)))
    lat = []
    t0 = time.perf_counter()
    for _ in range(steps):
        s0 = time.perf_counter()
        list(pool.map(one, range(cores)))
        lat.append(time.perf_counter() - s0)
    total = time.perf_counter() - t0
    qps = cores * steps / total
    line = {
        "impl": "reference", "metric": "pir_queries_per_sec", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][5], "queries_per_step": cores, "num_pt": params.num_pt,
                   "dims": dims},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": sample + "; %d concurrent queries per step, one per thread" % cores,
                         "single_query_latency_ms": 1e3 * t_one},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference path (SEAL 3.5.6 algorithms; SEAL itself is not buildable offline)",
    }
    print(json.dumps(line))


# =====================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist

    import pir_b200 as pb
    from pir_b200 import _lib, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    params = make_params(args.workload)
    ep = params.encryption_parameters
    N, mods = ep.poly_modulus_degree, ep.coeff_modulus
    k = len(mods) - 1
    ctL = 2 * k * N
    srv = sharded.ShardServer(params, device=local_rank, shard_index=rank, shard_count=world)
    srv.db.fill_random(2024)
    ql = args.queries_per_gpu
    queries, elts, keys = synth_inputs(params, world * ql, 99)
    gk = pb.GaloisKeys(elts, keys.reshape(-1))
    srv.set_keys(gk)
    n_ct = queries.shape[1]
    q_local = queries[rank * ql:(rank + 1) * ql]
    d_q = sharded.to_device(q_local, dev)
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def flush_l2():
        # read (not write) a buffer larger than L2: everything of ours is evicted and only CLEAN lines are left behind,
        # so the timed step does not also pay for writing back 126 MB of dirty flush data
        return flush.view(torch.int64).sum()

    if world > 1 and args.p2p:
        # peer-memory exchange needs CUDA IPC between the ranks' devices; if any rank cannot set it up, every rank
        # falls back to the NCCL gather so the run still produces a number (and says so in config.parallelism)
        ok = 1
        try:
            srv.setup_peer_exchange(max_queries=world * ql)
        except Exception as e:  # noqa: BLE001
            ok = 0
            sys.stderr.write("rank %d: peer exchange unavailable (%s); using NCCL gather\n" % (rank, e))
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        args.p2p = int(flag.item())

    def step_dev():
        if world == 1:
            return srv.answer(d_q)
        if args.p2p:
            return srv.answer_batch_distributed_p2p(d_q)
        return srv.answer_batch_distributed(d_q)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident throughput ----------------
    out = None
    for _ in range(max(3, args.warmup)):
        flush_l2()
        out = step_dev()  # held like in the timed loop, so the caching allocator already owns both reply buffers
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush_l2()  # L2 flush between timed iterations, outside the per-step events
        ev[i][0].record()
        out = step_dev()
        ev[i][1].record()
    if sampler is not None:
        sampler.sample_until(ev[-1][1])  # clocks under load: the queue of timed steps is still executing
    barrier()
    clocks = sampler.result() if sampler is not None else None
    launches = srv.launch_count() * args.steps
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    # per-stage device times (CUDA events recorded by the library on its launching stream), measured live on a few
    # extra steps: profiling inserts events between kernels, so those steps run eagerly instead of as one graph
    srv.set_profiling(True)
    stage_acc = {}
    for i in range(min(args.steps, 10)):
        flush_l2()
        step_dev()
        torch.cuda.synchronize()
        for nm, v in srv.stage_ms().items():
            stage_acc.setdefault(nm, []).append(v)
    srv.set_profiling(False)  # back to graph replay: the end-to-end leg below must run the product's normal path
    qps = world * ql * args.steps / (total_ms / 1e3)

    # ---------------- end to end through the reference-facing call, host buffers ----------------
    q_pin = torch.empty(q_local.shape, dtype=torch.int64).pin_memory()
    q_pin.copy_(torch.from_numpy(q_local.view(np.int64)))
    out_pin = torch.empty((ql, srv.ctx.reply_cts, 2, k, N), dtype=torch.int64).pin_memory()
    q_np = q_pin.numpy().view(np.uint64)
    out_np = out_pin.numpy().view(np.uint64)
    if world == 1:
        server = pb.PIRServer(srv.db, params)
        req = pb.Request([q_np[i] for i in range(ql)], gk)

        def step_e2e():
            server.ProcessRequest(req, out=out_np)  # H2D + kernels + D2H + sync inside pirb_answer
    else:
        def step_e2e():
            d = q_pin.to(dev, non_blocking=True)
            r = srv.answer_batch_distributed_p2p(d) if args.p2p else srv.answer_batch_distributed(d)
            out_pin.copy_(r, non_blocking=True)
            torch.cuda.synchronize()
    def time_e2e(n):
        for _ in range(3):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            step_e2e()
        barrier()
        return time.perf_counter() - t0

    # single GPU: also the explicitly staged variant (cudaMemcpyAsync H2D + D2H around the kernels) for comparison
    staged_qps = None
    if world == 1:
        prev = os.environ.get("PIRB_ZERO_COPY")
        os.environ["PIRB_ZERO_COPY"] = "0"
        staged_qps = ql * args.steps / time_e2e(args.steps)
        if prev is None:
            del os.environ["PIRB_ZERO_COPY"]
        else:
            os.environ["PIRB_ZERO_COPY"] = prev
    for _ in range(3):
        step_e2e()
    barrier()
    lat = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s0 = time.perf_counter()
        step_e2e()
        lat.append(time.perf_counter() - s0)
    barrier()
    e2e_total = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
    e2e_qps = world * ql * args.steps / float(e2e_total.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the scan kernel (HBM-bound) ----------------
    peak, peak_src = measured_peak()
    scan_ms = statistics.mean(stage_acc["scan"])
    nq_scan = ql if world == 1 else world * ql
    if world > 1 and len(params.dimensions) == 1:
        nq_scan = getattr(srv, "last_partial_batch", nq_scan)  # d=1: the stage times describe the last chunk of queries
    scan_bytes = srv.scan_bytes(nq_scan)
    achieved = scan_bytes / (scan_ms * 1e-3) / 1e9
    stage_mean = {nm: statistics.mean(v) for nm, v in stage_acc.items()}

    # ---------------- CPU baseline: the oracle port on one host core, bounded sample ----------------
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import binding as ob
        orc = ob.Oracle(N, mods, ep.plain_modulus)
        db_host = srv.db.read_ntt(0, params.num_pt) if params.num_pt * k * N * 8 <= (4 << 30) else None
        if db_host is not None:
            t0 = time.perf_counter()
            want = orc.process_query(db_host, params.dimensions, elts, keys.reshape(-1), q_local[0])
            t_one = time.perf_counter() - t0
            n_more = int(max(0, min(8, 15.0 / max(t_one, 1e-3) - 1)))
            ts = [t_one]
            for _ in range(n_more):
                t0 = time.perf_counter()
                orc.process_query(db_host, params.dimensions, elts, keys.reshape(-1), q_local[0])
                ts.append(time.perf_counter() - t0)
            med = statistics.median(ts)
            got = sharded.to_host(step_dev())[0]
            parity = bool(np.array_equal(got, want))
            cpu = {"value": 1.0 / med, "unit": "queries/s", "cores": 1, "kind": "port",
                   "sample": "%d full queries of this workload on one host thread (median); reply compared limb-for-limb "
                             "with the GPU reply: %s" % (len(ts), "identical" if parity else "MISMATCH"),
                   "p50_latency_ms": 1e3 * med, "host_cpus": os.cpu_count()}

    # ---------------- expansion (FP64-pipe-bound): algorithmic FP64 operations of the key switches / stage time -----
    # per key switch: k(k+1) forward + 2(k+1) inverse transforms of N/2*log2(N) butterflies at 8 FP64 ops, 2k(k+1)N
    # digit-key products at 8, canonicalisation of the k(k+1) digit transforms at 4 per coefficient, the final
    # psi^-i/N scaling of the 2(k+1) inverse transforms at 7, the mod-down of 2k polynomials at 20
    logn = N.bit_length() - 1
    ks_per_query = sum(int(pb.next_power_two(min(N, max(0, sum(params.dimensions) - t * N)))) - 1 for t in range(n_ct))
    ops_per_ks = ((k * (k + 1) + 2 * (k + 1)) * (N // 2) * logn * 8 + 2 * k * (k + 1) * N * 8 + k * (k + 1) * N * 4
                  + 2 * (k + 1) * N * 7 + 2 * k * N * 20)
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp64_peak = 57.7 * torch.cuda.get_device_properties(dev).multi_processor_count * sm_mhz * 1e6 / 1e12
    exp_ms = stage_mean["expand"]
    exp_ach = ql * ks_per_query * ops_per_ks / (exp_ms * 1e-3) / 1e12
    expansion = {"kernel": "k_ks_level_cluster", "bound": "fp64 pipe", "key_switches_per_launch_set": ql * ks_per_query,
                 "fp64_ops_per_key_switch": ops_per_ks, "achieved": exp_ach, "peak": fp64_peak, "unit": "T FP64 op/s",
                 "frac": exp_ach / fp64_peak, "ms": exp_ms, "share_of_step": exp_ms / stage_mean["total"],
                 "peak_source": "57.7 DFMA/clk/SM measured with tools/pipe_bench.cu x SM count x SM clock under load",
                 "note": "includes the selection-vector exchange when n_gpus > 1"}

    traffic = args.scan_traffic
    if traffic is None and world == 1 and ql == 1:
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of one scan launch, from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "r1_scan_traffic.json")) as f:
                t = json.load(f)
            if t.get("workload") == args.workload:
                traffic = t["traffic_bytes"]
        except Exception:  # noqa: BLE001
            traffic = None
    line = {
        "metric": "pir_queries_per_sec", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][5], "queries_per_step": world * ql, "num_pt": params.num_pt,
                   "dims": list(params.dimensions), "key_switches_per_query": sum(
                       int(pb.next_power_two(min(N, max(0, sum(params.dimensions) - t * N)))) - 1 for t in range(n_ct)),
                   "reply_cts": srv.ctx.reply_cts, "parallelism": "rows sharded x%d, expansion split by query%s" % (
                       world, "" if world == 1 else (", partial replies reduced over NVLink peer loads" if args.p2p
                                                    else ", partial replies gathered with NCCL")),
                   "l2": "256 MiB buffer read between timed iterations (evicts L2, leaves clean lines; outside the per-step CUDA events)",
                   "galois_keys": "resident in HBM, uploaded once per client"},
        "p50_latency_ms": statistics.median(step_ms),
        "step_ms": [round(x, 4) for x in step_ms],
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": int(world * ql * n_ct * ctL * 8),
                "d2h_bytes_per_step": int(world * ql * srv.ctx.reply_cts * ctL * 8),
                "p50_latency_ms": 1e3 * statistics.median(lat),
                "staged_copies_value": staged_qps,
                "transfer": ("pinned host buffers; the first kernel reads the queries and the last one writes the "
                             "replies over PCIe in place, inside the timed region (PIRB_ZERO_COPY=0: staged "
                             "cudaMemcpyAsync both ways)") if world == 1 else
                            "pinned host buffers, cudaMemcpyAsync H2D of the queries and D2H of the replies every step"},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "k_scan", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "bytes_per_launch": scan_bytes, "ms_per_launch": scan_ms,
                     "share_of_step": scan_ms / stage_mean["total"]},
        "expansion": expansion,
        "stages_ms": stage_mean,
        "cpu_baseline": cpu,
        "parity_vs_oracle": parity,
        "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--queries-per-gpu", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--p2p", type=int, default=1,
                    help="N>1: combine partial replies with peer-memory loads in the reduce kernel (1) or an NCCL gather (0)")
    ap.add_argument("--scan-traffic", type=float, default=None,
                    help="dram bytes per scan launch; default: the committed ncu capture in profiles/ for this workload")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself as one rank per GPU (the driver's own launch line)
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                                   "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
