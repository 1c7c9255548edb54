#!/usr/bin/env python
"""bench.py — IMLE kNN matching throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c1|small] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One step = one pass of the hot path over one batch: all Q queries of the workload matched against the
resident generated pool (exact kNN).  Prints ONE JSON line (rank 0).

  value     queries/s with the pool index built and the query matrix already resident in HBM
            (C-ABI device entry points, CUDA-event timed on the launching stream, max over ranks).
  e2e       the same metric through the host-buffer C-ABI call the DCI Python class makes
            (b200knn_query): pinned HOST float64 queries in, HOST results out, H2D/D2H inside the timed region.
  roofline  the tcgen05 distance kernel: 2*Q*N*d FLOPs per launch / its CUDA-event time, vs MEASURED_PEAKS.json.
  cpu_baseline  the UNMODIFIED reference DCI (oracle/_ref/_dci.so) on this box's host cores, bounded sample.

torch is used for device memory, streams, events and torch.distributed only; no torch op is on the path.
Multi-GPU: pool row-sharded over ranks, queries replicated, local exact top-k per rank, NCCL all-gather of
(index, distance) lists, k-way merge kernel on every rank.  Total work is fixed as N grows -> "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N pool, Q queries, dim, k, description)   — BASELINE.json configs
    "c3": (300000, 30000, 3072, 1, "CelebA-128 IMLE match: 30k queries vs 300k pool, d=3072, k=1 (BASELINE.json configs[2]; north_star target shape)"),
    "c2": (240000, 24000, 3072, 1, "Stacked MNIST IMLE match: 24k queries vs 240k pool, d=3072, k=1 (configs[1])"),
    "c4": (50000, 50000, 2048, 4, "precision/recall self-kNN: 50k vs 50k, d=2048, k=3(+self) (configs[3])"),
    "c1": (10000, 100, 5000, 10, "dci_code/example.py shape: 10k pool, 100 queries, d=5000, k=10 (configs[0])"),
    "small": (20000, 2048, 512, 1, "smoke-sized"),
    # configs[4]: 98 GB of BF16 pool + 197 GB of float32 originals -> needs >= 2 GPUs (>= 4 recommended); "c5s" is the
    # 125k-row share one rank holds in the 8-GPU run, for single-GPU measurements
    "c5": (1000000, 30000, 49152, 10, "scale sweep: 1M pool, d=49152 (128x128x3 raw pixels), 30k queries, k=10 (configs[4])"),
    "c5s": (125000, 30000, 49152, 10, "one rank's 1/8 share (125k rows) of the configs[4] scale sweep: d=49152, 30k queries, k=10"),
}
IMAGE_LIKE = ("c5", "c5s")       # float32 image-like features; the others: float64 N(0,1)
IMAGE_LATENT = 16


def synth_rows(workload, a, b, d, dev, seed_base):
    """Rows [a, b) of a workload's synthetic feature matrix, generated on the device; a row's values depend only on
    (seed_base, global row block), so every sharding sees the same matrix."""
    import torch
    if workload not in IMAGE_LIKE:
        raise ValueError(workload)
    # image-like: a 16-dimensional latent through a fixed random basis + pixel noise, clipped to [-1, 1] (the
    # trainer's pixel range, training_loop.py:362-365), float32 like the generator's output
    gb = torch.Generator(device=dev)
    gb.manual_seed(77)
    basis = torch.randn(IMAGE_LATENT, d, device=dev, dtype=torch.float32, generator=gb) * 0.125
    out = torch.empty(b - a, d, device=dev, dtype=torch.float32)
    SB = 4096
    for sb in range(a // SB, (b + SB - 1) // SB):
        g = torch.Generator(device=dev)
        g.manual_seed(seed_base + sb)
        lat = torch.randn(SB, IMAGE_LATENT, device=dev, dtype=torch.float32, generator=g)
        x = lat @ basis
        x += 0.05 * torch.randn(SB, d, device=dev, dtype=torch.float32, generator=g)
        x.clamp_(-1.0, 1.0)
        lo, hi = max(a, sb * SB), min(b, (sb + 1) * SB)
        out[lo - a:hi - a] = x[lo - sb * SB:hi - sb * SB]
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(object):
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe), through NVML
    (the library nvidia-smi itself reads) from a thread, every 10 ms; mark() delimits the timed region."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False
        self.thread = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.gpu]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self.stop_flag:
                    try:
                        try:
                            rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.rows.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                          pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, int(rs)))
                    except Exception:
                        pass
                    time.sleep(0.01)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is None:
            return out
        self.stop_flag = True
        self.thread.join(timeout=2)
        rows = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or 1e300)] or self.rows
        if rows:
            mask = 0
            for r in rows:
                mask |= r[3]
            out = {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": self.max_mhz,
                   "reasons": [n for n, b in self.REASONS if mask & b], "samples": len(rows),
                   "power_w_median": float(np.median([r[2] for r in rows])), "power_w_max": float(max(r[2] for r in rows))}
        return out


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference DCI on host cores, bounded sample of the workload
# ---------------------------------------------------------------------------------------------------------
def reference_sample(workload, steps, warmup, budget_s=25.0):
    """Times oracle/_ref (reference DCI, training hyper-parameters training_loop.py:197,368,398).

    Sample: the workload's dim and k, a pool subsample sized for the box's core count and a query
    subsample per step (per-query cost does not depend on Q: dci.c:801 parallelises over queries)."""
    from oracle import ref_dci
    n, q, d, k, _ = WORKLOADS[workload]
    cores = host_cores()
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    if not ref_dci.available():
        return None
    ns = int(min(n, 120000, max(2048, int(3.5e9) // (8 * d))))        # pool subsample bounded to 3.5 GB of float64
    qs = int(min(q, max(64, 16 * cores)))
    rng = np.random.default_rng(0)
    pool = rng.standard_normal((ns, d))
    queries = np.random.default_rng(1).standard_normal((qs, d))
    if workload == "c1":
        m, L, levels, cfov, cpr, qfov, qpr = 2, 7, 2, 10, 0.002, 100, 0.05      # dci_code/example.py:44-66
    else:
        m, L, levels, cfov, cpr, qfov, qpr = 3, 15, 3, 10, 0.002, 200, 1.0       # training/training_loop.py:197,368,398
    db = ref_dci.RefDCI(d, m, L)
    t0 = time.perf_counter()
    db.add(pool, num_levels=levels, field_of_view=cfov, prop_to_retrieve=cpr)
    t_add = time.perf_counter() - t0
    times = []
    t_begin = time.perf_counter()
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        db.query(queries, k, field_of_view=qfov, prop_to_retrieve=qpr)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > 4 * budget_s and len(times) >= 1:
            break
    db.clear()
    total = float(np.sum(times))
    return {"qps": qs * len(times) / total, "ms_per_step": 1e3 * total / len(times), "steps_done": len(times), "cores": cores,
            "add_s": t_add,
            "sample": "reference DCI (oracle/_ref, unmodified dci.c) m=%d L=%d levels=%d; pool subsample %d x %d float64 N(0,1) of the "
                      "%d-row workload, %d queries/step, k=%d, OMP threads=%d; add() took %.1f s (not in value)" % (
                          m, L, levels, ns, d, n, qs, k, cores, t_add)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = reference_sample(args.workload, args.steps, max(args.warmup, 1))
    n, q, d, k, desc = WORKLOADS[args.workload]
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/_dci.so missing (build with `make -C oracle ref` where /root/reference exists)"}))
        return 0
    line = {"impl": "reference", "metric": "imle_knn_queries_per_sec", "value": r["qps"], "unit": "queries/s", "n_gpus": args.gpus,
            "steps": r["steps_done"], "warmup": max(args.warmup, 1), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, desc), "pool": n, "queries": q, "dim": d, "k": k},
            "cpu_baseline": {"value": r["qps"], "unit": "queries/s", "cores": r["cores"], "kind": "reference", "sample": r["sample"]},
            "e2e": {"value": r["qps"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from inclusivegan_b200.dci import DeviceKNN, F32, F64, load_library
    import ctypes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, q, d, k, desc = WORKLOADS[args.workload]
    image_like = args.workload in IMAGE_LIKE
    fdt, FT, fbytes = (torch.float32, F32, 4) if image_like else (torch.float64, F64, 8)
    per = (n + world - 1) // world
    r0, r1 = min(n, per * rank), min(n, per * (rank + 1))
    need_gb = (r1 - r0) * d * (fbytes + 2) / 1e9 + q * d * (fbytes + 2) / 1e9
    if need_gb > 150.0:
        raise RuntimeError("workload %s needs %.0f GB per GPU at %d GPU(s): use more GPUs (pool rows are sharded)" % (args.workload, need_gb, world))
    if image_like:
        # ---- image-like float32 features generated on the device per row block (configs[4]) ----
        pool = synth_rows(args.workload, r0, r1, d, dev, 1000)
        queries = synth_rows(args.workload, 0, q, d, dev, 500000)
    else:
        # ---- synthetic features (float64, the dtype of the reference's interface), identical on every rank ----
        # the pool is generated in 8 fixed row blocks, each seeded by its block id, so 1/2/4/8-GPU runs see the same rows
        pool = torch.empty(r1 - r0, d, device=dev, dtype=torch.float64)
        blk = (n + 7) // 8
        for b in range(8):
            b0, b1 = max(r0, b * blk), min(r1, (b + 1) * blk, n)
            if b1 <= b0:
                continue
            g = torch.Generator(device=dev)
            g.manual_seed(1000 + b)
            full = torch.randn(min((b + 1) * blk, n) - b * blk, d, device=dev, dtype=torch.float64, generator=g)
            pool[b0 - r0:b1 - r0] = full[b0 - b * blk:b1 - b * blk]
            del full
        gq = torch.Generator(device=dev)
        gq.manual_seed(1)
        queries = torch.randn(q, d, device=dev, dtype=torch.float64, generator=gq)

    stream = torch.cuda.current_stream()
    ix = DeviceKNN(d, local_rank)
    ix.set_stream(stream.cuda_stream)
    ix.add(pool.data_ptr(), FT, r1 - r0, index_base=r0)
    torch.cuda.synchronize()
    kk = min(k, n)
    loc_i = torch.empty(q, kk, device=dev, dtype=torch.int32)
    loc_d = torch.empty(q, kk, device=dev, dtype=torch.float64)
    out_i = torch.empty(q, kk, device=dev, dtype=torch.int32)
    out_d = torch.empty(q, kk, device=dev, dtype=torch.float64)
    if world > 1:
        all_i = torch.empty(world, q, kk, device=dev, dtype=torch.int32)
        all_d = torch.empty(world, q, kk, device=dev, dtype=torch.float64)

    # ---- multi-GPU exchange: NVLink peer-memory all-gather + merge (no NCCL on the path); NCCL only as a fallback ----
    exchange = None
    exchange_kind = "none"
    if world > 1:
        from inclusivegan_b200.dci import PeerExchange
        try:
            if args.exchange == "nccl":
                raise RuntimeError("NCCL exchange requested")
            exchange = PeerExchange(local_rank, rank, world, q, kk)
            handles = [None] * world
            dist.all_gather_object(handles, exchange.handle())       # once, at start-up: 64 bytes per rank
            exchange.connect(handles)
            ok = torch.tensor([1], device=dev)
        except Exception as e:
            exchange = None
            ok = torch.tensor([0], device=dev)
            if rank == 0 and args.exchange != "nccl":
                sys.stderr.write("peer-memory exchange unavailable (%r); falling back to NCCL all-gather\n" % (e,))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)                     # all ranks or none
        if int(ok.item()) == 0:
            exchange = None
        exchange_kind = "nvlink-peer-stores" if exchange is not None else "nccl-allgather"

    def exchange_and_merge():
        if exchange is not None:
            # publish kernel (P2P stores into every peer's buffer + step flag) and flag-waiting merge kernel
            exchange.allgather_merge(loc_i.data_ptr(), loc_d.data_ptr(), q, kk, out_i.data_ptr(), out_d.data_ptr(), stream.cuda_stream)
            return
        # NCCL all-gather of the per-shard (index, distance) lists over NVLink, then the k-way merge kernel.
        # The device-wide synchronisation keeps the collective's kernels (which may run on NCCL's own stream and
        # spin until the peer arrives) from sharing the GPU with the next step's persistent distance kernel, which
        # wants every SM: measured at N=2, overlapping them costs +45 % on the distance kernel.
        dist.all_gather_into_tensor(all_i.view(world * q, kk), loc_i)
        dist.all_gather_into_tensor(all_d.view(world * q, kk), loc_d)
        torch.cuda.synchronize()
        ix.merge(all_i.data_ptr(), all_d.data_ptr(), world, q, kk, out_i.data_ptr(), out_d.data_ptr(), stream.cuda_stream)

    def step_device():
        ix.query(queries.data_ptr(), FT, q, k, loc_i.data_ptr(), loc_d.data_ptr())
        if world > 1:
            exchange_and_merge()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm ------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    ix.reset_stats()
    ix.set_profiling(True)
    sampler.mark_begin()
    total_ms = timed(step_device, args.steps)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    st = ix.stats()
    ix.set_profiling(False)
    launches = st["kernel_launches"] + ((2 if exchange is not None else 1) * args.steps if world > 1 else 0)

    # ---- end-to-end arm: host-buffer C-ABI call (what DCI.query makes), pinned host queries ----------
    lib = load_library()
    hq = torch.empty(q, d, dtype=fdt).pin_memory()
    hq.copy_(queries)
    torch.cuda.synchronize()
    hx = ctypes.c_void_p()
    ids = (ctypes.c_int * 1)(local_rank)
    assert lib.b200knn_create(d, 1, ids, ctypes.byref(hx)) == 0
    assert lib.b200knn_set_stream(hx, ctypes.c_void_p(stream.cuda_stream)) == 0
    assert lib.b200knn_add_device(hx, ctypes.c_void_p(pool.data_ptr()), FT, r1 - r0, d, r0) == 0, lib.b200knn_last_error()
    h_i = torch.empty(q, kk, dtype=torch.int32).pin_memory()
    h_d = torch.empty(q, kk, dtype=torch.float64).pin_memory()
    res_i = torch.empty(q, kk, dtype=torch.int32).pin_memory()
    res_d = torch.empty(q, kk, dtype=torch.float64).pin_memory()

    # N > 1: the replicated query matrix is uploaded ONCE per box — every rank copies its 1/N row slice from pinned host
    # memory and the slices are all-gathered over NVLink ("queries are broadcast", north_star) — instead of N full
    # uploads competing for the host's memory bandwidth (measured at N=8: 34 ms/step that way).
    q_pad = (q + world - 1) // world * world
    q_per = q_pad // world
    if world > 1:
        qa, qb = min(q, rank * q_per), min(q, (rank + 1) * q_per)
        hq_slice = torch.zeros(q_per, d, dtype=fdt).pin_memory()
        hq_slice[:qb - qa].copy_(queries[qa:qb])
        dq_full = torch.empty(q_pad, d, device=dev, dtype=fdt)
        torch.cuda.synchronize()

    sliced = world >= 4          # N <= 2: every rank runs the pipelined host-buffer call (upload hidden behind compute)

    def step_e2e():
        if not sliced:
            rc = lib.b200knn_query(hx, ctypes.c_void_p(hq.data_ptr()), FT, q, d, k, 0, ctypes.c_void_p(h_i.data_ptr()),
                                   ctypes.c_void_p(h_d.data_ptr()), None)
            if rc != 0:
                raise RuntimeError(lib.b200knn_last_error().decode())
            if world > 1:     # shard results back to the device for the NVLink exchange, merged result back to the host
                loc_i.copy_(h_i, non_blocking=True)
                loc_d.copy_(h_d, non_blocking=True)
                exchange_and_merge()
                res_i.copy_(out_i, non_blocking=True)
                res_d.copy_(out_d, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            return
        dq_full[rank * q_per:(rank + 1) * q_per].copy_(hq_slice, non_blocking=True)          # H2D of this rank's slice
        dist.all_gather_into_tensor(dq_full, dq_full[rank * q_per:(rank + 1) * q_per])       # NVLink broadcast of the slices
        torch.cuda.synchronize()       # keep NCCL's kernels off the SMs the persistent distance kernel wants
        ix.query(dq_full.data_ptr(), FT, q, k, loc_i.data_ptr(), loc_d.data_ptr())
        exchange_and_merge()
        res_i.copy_(out_i, non_blocking=True)                                                 # D2H of the merged result
        res_d.copy_(out_d, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    e2e_ms = timed(step_e2e, e2e_steps)
    lib.b200knn_destroy(hx)

    # ---- self-check of the last device result against a float64 torch brute force on a query subsample ----
    nchk = min(q, 64)
    sub = queries[:nchk].double()
    d2 = torch.empty(nchk, r1 - r0, device=dev, dtype=torch.float64)
    for c0 in range(0, r1 - r0, 8192):                # float64 whatever the feature dtype
        pc = pool[c0:c0 + 8192].double()
        d2[:, c0:c0 + 8192] = (sub * sub).sum(1, keepdim=True) + (pc * pc).sum(1)[None, :] - 2.0 * sub @ pc.T
    del pc
    tk = torch.topk(d2, min(kk, r1 - r0), dim=1, largest=False)
    if world == 1:
        check = bool((tk.indices.to(torch.int32) == loc_i[:nchk]).all().item()) if rank == 0 else None
    else:
        # global answer of the subsample from per-shard torch answers (float64), compared with the merged result
        step_device()
        torch.cuda.synchronize()
        cand_d = torch.full((nchk, kk), float("inf"), device=dev, dtype=torch.float64)
        cand_i = torch.full((nchk, kk), -1, device=dev, dtype=torch.int64)
        cand_d[:, :tk.values.shape[1]] = tk.values
        cand_i[:, :tk.indices.shape[1]] = tk.indices + r0
        g_d = torch.empty(world * nchk, kk, device=dev, dtype=torch.float64)
        g_i = torch.empty(world * nchk, kk, device=dev, dtype=torch.int64)
        dist.all_gather_into_tensor(g_d, cand_d)
        dist.all_gather_into_tensor(g_i, cand_i)
        torch.cuda.synchronize()
        g_d = g_d.view(world, nchk, kk).permute(1, 0, 2).reshape(nchk, world * kk)
        g_i = g_i.view(world, nchk, kk).permute(1, 0, 2).reshape(nchk, world * kk)
        best = torch.topk(g_d, kk, dim=1, largest=False).indices
        ref = torch.gather(g_i, 1, best).to(torch.int32)
        check = bool((ref == out_i[:nchk]).all().item()) if rank == 0 else None

    if rank == 0:
        peaks = load_peaks()
        ms_step = total_ms / args.steps
        dist_ms = st["ms_distance"] / max(st["distance_launches"], 1)
        ach = st["distance_flops"] / max(st["ms_distance"], 1e-9) / 1e9     # TFLOP/s per GPU (this rank)
        traffic = None
        prof = os.path.join(ROOT, "profiles", "dist_kernel_ncu.json")
        if os.path.exists(prof):
            try:
                with open(prof) as fh:
                    traffic = json.load(fh).get(args.workload, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "imle_knn_queries_per_sec", "value": q / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16 (tensor pass) + f64 (exact re-rank)", "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, desc), "pool": n, "queries": q, "dim": d, "k": k,
                       "features": ("float32 image-like (16-d latent through a fixed basis + 0.05 pixel noise, clipped to [-1,1]), seeded per 4096-row block, generated on device" if image_like else "float64 N(0,1), seeded"), "parallelism": "pool row-sharded x%d, queries replicated, exchange=%s + k-way merge kernel" % (world, exchange_kind)
                       if world > 1 else "single GPU", "l2": "inputs exceed L2 (BF16 pool shard %.2f GB > 126 MB); no explicit flush" % ((r1 - r0) * d * 2 / 1e9),
                       "timing": "CUDA events on the launching stream, max over ranks"},
            "tensor_peak_frac": 2.0 * q * n * d / (ms_step * 1e-3) / 1e12 / (peaks["bf16_tflops"] * world),
            "roofline": {"bound": "tensor", "kernel": "dist_topc_kernel (tcgen05 BF16 distance GEMM + fused top-C)",
                         "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                         "frac_of_sustained_peak": (ach / peaks["bf16_tflops_sustained"]) if peaks.get("bf16_tflops_sustained") else None,
                         "peak_source": peaks["source"] + ", burst figure", "ms_per_launch": dist_ms, "launches": st["distance_launches"],
                         "flops_per_launch": st["distance_flops"] / max(st["distance_launches"], 1), "traffic": traffic},
            "kernel_ms_per_step": {"convert": st["ms_convert"] / args.steps, "distance": st["ms_distance"] / args.steps,
                                   "rerank": st["ms_rerank"] / args.steps, "second_pass": st["ms_scan"] / args.steps},
            "uncertified_per_step": st["uncertified"] / args.steps,
            "e2e": {"value": q / (e2e_ms / e2e_steps * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "h2d_bytes_per_step": q * d * fbytes, "d2h_bytes_per_step": q * kk * 12,
                    "api": ("b200knn_query (host buffers; the call inclusivegan_b200.dci.DCI.query makes)" + ("" if world == 1 else " per rank + peer exchange + D2H of the merged result")) if not sliced else
                           "per rank: H2D of a 1/N query slice from pinned host memory, NVLink all-gather of the slices, b200knn_query_device, peer exchange, D2H of the merged result"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "self_check_top%d_vs_torch_f64" % kk: check,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                r = reference_sample(args.workload, steps=2, warmup=1)
                if r is not None:
                    line["cpu_baseline"] = {"value": r["qps"], "unit": "queries/s", "cores": r["cores"], "kind": "reference", "sample": r["sample"]}
                else:
                    line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": host_cores(), "kind": "reference",
                                            "sample": "unavailable: oracle/_ref/_dci.so missing"}
            except Exception as e:   # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": host_cores(), "kind": "reference", "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="multi-GPU result exchange: NVLink peer-memory kernels (default) or NCCL all-gather")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
