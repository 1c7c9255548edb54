#!/usr/bin/env python
"""bench.py — IMLE kNN matching throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c1|c5|c5s|small] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One step = one pass of the hot path over one batch: all Q queries of the workload matched against the
resident generated pool (exact kNN).  Prints ONE JSON line (rank 0).

  value     queries/s with the pool index built and the query matrix already resident in HBM
            (C-ABI device entry points, CUDA-event timed on the launching stream, max over ranks).
  e2e       the same metric through the host-buffer C-ABI call the DCI Python class makes
            (b200knn_query): pinned HOST queries in, HOST results out, H2D/D2H inside the timed region.
  roofline  the tcgen05 distance kernel: 2*Q*N*d FLOPs per launch / its CUDA-event time, vs MEASURED_PEAKS.json.
  add_s     DCI.add() of the workload's pool from pageable host memory (H2D + centre + BF16 convert + norms).
  small_call  the trainer's call granularity (training_loop.py:374-403): 24-row b200knn_query calls, back to back.
  e2e_pageable  the e2e call again with pageable NumPy buffers (what DCI.query hands over), N = 1.
  cpu_baseline / --impl reference
            the UNMODIFIED reference DCI (oracle/_ref/_dci.so) on this box's host cores with the trainer's
            hyper-parameters, FULL pool, bounded query sample per step; its approximate answers are scored as
            recall@k against the exact float64 answer on the same queries.

torch is used for device memory, streams, events and torch.distributed only; no torch op is on the path.
Multi-GPU: pool row-sharded over ranks, queries replicated, local exact top-k per rank, NVLink peer-store
exchange of (index, distance) lists, k-way merge kernel on every rank.  Total work is fixed as N grows -> "strong".
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "imle_knn_queries_per_sec"

# name: (N pool, Q queries, dim, k, feature generator, description)   — BASELINE.json configs, generators per SURVEY.md 8d
WORKLOADS = {
    "c3": (300000, 30000, 3072, 1, "gauss", "CelebA-128 IMLE match: 30k queries vs 300k pool, d=3072, k=1 (BASELINE.json configs[2]; north_star target shape)"),
    "c2": (240000, 24000, 3072, 1, "pixels", "Stacked MNIST IMLE match: 24k queries vs 240k pool, d=3072, k=1 (configs[1])"),
    "c4": (50000, 50000, 2048, 4, "relu", "precision/recall metric: self-kNN radii of two 50k x 2048 feature sets, k=3 (+self) (configs[3])"),
    "c1": (10000, 100, 5000, 10, "lowrank", "dci_code/example.py shape: 10k pool, 100 queries, d=5000, k=10 (configs[0])"),
    "small": (20000, 2048, 512, 1, "gauss", "smoke-sized"),
    # configs[4]: 98 GB of BF16 pool + 197 GB of float32 originals -> needs >= 4 GPUs; "c5s" is the 125k-row share one
    # rank holds in the 8-GPU run, for single-GPU measurements
    "c5": (1000000, 30000, 49152, 10, "image", "scale sweep: 1M pool, d=49152 (128x128x3 raw pixels), 30k queries, k=10 (configs[4])"),
    "c5s": (125000, 30000, 49152, 10, "image", "one rank's 1/8 share (125k rows) of the configs[4] scale sweep: d=49152, 30k queries, k=10"),
}
# feature generators.  dtype = what the caller hands to the library.
GENERATORS = {
    "gauss":   ("float64", "N(0,1) drawn in float32 and widened to float64 like the trainer's .astype(float64) (training_loop.py:363,379)"),
    "pixels":  ("float64", "clip(N(0,0.5),-1,1) drawn in float32 ([-1,1] image range) and widened to float64 (training_loop.py:363,379)"),
    "relu":    ("float32", "relu(N(0,1)) float32, Inception-pool-like features (metrics/precision_recall.py:184-216 hands float32)"),
    "lowrank": ("float64", "dci_code/example.py:36-40: (2U-1)[rows x 50] @ (2U-1)[50 x d], float64"),
    "image":   ("float32", "float32 image-like (16-d latent through a fixed basis + 0.05 pixel noise, clipped to [-1,1]), generated on device"),
}
IMAGE_LATENT = 16
ROW_BLOCK = 4096           # rows are generated per 4096-row block seeded by (seed base + block id): any sharding sees the same matrix
POOL_SEED, QUERY_SEED, SET_B_SEED = 1000, 500000, 900000
SMALL_CALL_ROWS = 24       # 2 * minibatch with the README settings (run_training.py:67-68,200)
SMALL_CALLS = 1250         # calls per refresh at config 3 (30000 / 24)


def workload_config(name):
    """The `config` object of the JSON line — identical in both arms (the driver compares them)."""
    n, q, d, k, gen, desc = WORKLOADS[name]
    return {"workload": "%s: %s" % (name, desc), "pool": n, "queries": q, "dim": d, "k": k,
            "features": "%s; seeded per %d-row block" % (GENERATORS[gen][1], ROW_BLOCK)}


def _gen_block(gen, rows, d, seed, dev, low_t=None, basis=None):
    """One row block of a generator, on torch device `dev` (CPU works too: used by the CPU tests)."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    if gen == "gauss":
        return torch.randn(rows, d, device=dev, dtype=torch.float32, generator=g).double()
    if gen == "pixels":
        return (torch.randn(rows, d, device=dev, dtype=torch.float32, generator=g) * 0.5).clamp_(-1.0, 1.0).double()
    if gen == "relu":
        return torch.randn(rows, d, device=dev, dtype=torch.float32, generator=g).clamp_(min=0.0)
    if gen == "lowrank":
        lat = torch.rand(rows, 50, device=dev, dtype=torch.float64, generator=g) * 2.0 - 1.0
        return lat @ low_t
    if gen == "image":
        lat = torch.randn(rows, IMAGE_LATENT, device=dev, dtype=torch.float32, generator=g)
        x = lat @ basis
        x += 0.05 * torch.randn(rows, d, device=dev, dtype=torch.float32, generator=g)
        return x.clamp_(-1.0, 1.0)
    raise ValueError(gen)


def synth_rows(workload, a, b, d, dev, seed_base):
    """Rows [a, b) of a workload's synthetic feature matrix, generated on `dev`; a row's values depend only on
    (seed_base, global row block), so every sharding sees the same matrix."""
    import torch
    gen = WORKLOADS[workload][4]
    dt = torch.float64 if GENERATORS[gen][0] == "float64" else torch.float32
    low_t = basis = None
    if gen == "lowrank":
        g = torch.Generator(device=dev)
        g.manual_seed(77)
        low_t = torch.rand(50, d, device=dev, dtype=torch.float64, generator=g) * 2.0 - 1.0
    if gen == "image":
        g = torch.Generator(device=dev)
        g.manual_seed(77)
        basis = torch.randn(IMAGE_LATENT, d, device=dev, dtype=torch.float32, generator=g) * 0.125
    out = torch.empty(b - a, d, device=dev, dtype=dt)
    for sb in range(a // ROW_BLOCK, (b + ROW_BLOCK - 1) // ROW_BLOCK):
        x = _gen_block(gen, ROW_BLOCK, d, seed_base + sb, dev, low_t, basis)
        lo, hi = max(a, sb * ROW_BLOCK), min(b, (sb + 1) * ROW_BLOCK)
        out[lo - a:hi - a] = x[lo - sb * ROW_BLOCK:hi - sb * ROW_BLOCK]
    return out


def synth_rows_host(workload, a, b, d, seed_base, threads):
    """The same distributions on the HOST with NumPy (reference arm: no GPU, no torch), float64 out; values differ
    from the device generator's (another RNG), the distribution and the block seeding do not."""
    gen = WORKLOADS[workload][4]
    if gen == "image":
        raise ValueError("image-like rows are generated on the device only")
    low_t = None
    if gen == "lowrank":
        low_t = np.random.default_rng(77).random((50, d)) * 2.0 - 1.0
    out = np.empty((b - a, d), dtype=np.float64)

    def fill(sb):
        rng = np.random.default_rng(seed_base + sb)
        if gen == "lowrank":
            x = (rng.random((ROW_BLOCK, 50)) * 2.0 - 1.0) @ low_t
        else:
            x = rng.standard_normal((ROW_BLOCK, d), dtype=np.float32)
            if gen == "pixels":
                x = np.clip(x * np.float32(0.5), -1.0, 1.0)
            elif gen == "relu":
                x = np.maximum(x, 0.0)
        lo, hi = max(a, sb * ROW_BLOCK), min(b, (sb + 1) * ROW_BLOCK)
        out[lo - a:hi - a] = x[lo - sb * ROW_BLOCK:hi - sb * ROW_BLOCK]

    with ThreadPoolExecutor(max(1, threads)) as ex:
        list(ex.map(fill, range(a // ROW_BLOCK, (b + ROW_BLOCK - 1) // ROW_BLOCK)))
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
def force_omp_threads(cores):
    """torchrun exports OMP_NUM_THREADS=1 to its children when nproc > 1; the reference DCI parallelises its query loop
    with OpenMP (dci.c:801), so the thread count is ASSIGNED here (before libgomp is loaded with oracle/_ref/_dci.so)
    and then read back from the OpenMP runtime: the count reported is the one in effect."""
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ["OMP_STACKSIZE"] = "256M"          # dci.c:572-573 keeps ~1.3 MB VLAs on worker stacks
    os.environ.pop("OMP_THREAD_LIMIT", None)
    from oracle import ref_dci
    if not ref_dci.available():
        return None
    ref_dci.ext()                                # loads _dci.so (and libgomp with it)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(cores))     # also covers a libgomp that something loaded earlier
        gomp.omp_get_max_threads.restype = ctypes.c_int
        return int(gomp.omp_get_max_threads())
    except OSError:
        return None


def reference_params(workload):
    if workload == "c1":     # dci_code/example.py:44-66
        return dict(m=2, L=7, levels=2, cfov=10, cpr=0.002, qfov=100, qpr=0.05, source="dci_code/example.py:60-66")
    return dict(m=3, L=15, levels=3, cfov=10, cpr=0.002, qfov=200, qpr=1.0, source="training/training_loop.py:197,368,398")


def reference_sample(workload, steps, warmup, wall_budget_s=240.0, max_pool_gb=24.0):
    """Times oracle/_ref (the reference DCI) with the reference's own hyper-parameters on the FULL pool of the workload
    (bounded only by host memory), `steps` timed query calls of a bounded query sample each (per-query cost does not
    depend on Q: dci.c:801 parallelises over queries), and scores the approximate answers as recall@k against the
    exact float64 answer on the same queries."""
    from oracle import ref_dci
    from oracle import knn_oracle as ko
    n, q, d, k, gen, _ = WORKLOADS[workload]
    cores = host_cores()
    threads = force_omp_threads(cores)
    if threads is None:
        return None
    note = ""
    if gen == "image":
        # configs[4] is 197 GB even in float32: infeasible on the host; same dim and k on Gaussian rows
        gen_wl, note = "c3", " (configs[4] pool does not fit host memory: %d x %d subsample, Gaussian rows)" % (min(n, 60000), d)
        ns = min(n, 60000)
    else:
        gen_wl = workload
        ns = int(min(n, max_pool_gb * 1e9 // (8 * d)))
    qs = int(min(q, max(64, 16 * cores), 512))
    t0 = time.perf_counter()
    if gen_wl == workload:
        pool = synth_rows_host(workload, 0, ns, d, POOL_SEED, cores)
        queries = synth_rows_host(workload, 0, qs, d, QUERY_SEED, cores)
    else:
        rng = np.random.default_rng(0)
        pool = rng.standard_normal((ns, d), dtype=np.float32).astype(np.float64)
        queries = np.random.default_rng(1).standard_normal((qs, d), dtype=np.float32).astype(np.float64)
    if workload == "c4":
        queries = pool[:qs].copy()            # the precision/recall metric queries every row against its own set
    t_gen = time.perf_counter() - t0
    p = reference_params(workload)
    db = ref_dci.RefDCI(d, p["m"], p["L"])
    t0 = time.perf_counter()
    db.add(pool, num_levels=p["levels"], field_of_view=p["cfov"], prop_to_retrieve=p["cpr"])
    t_add = time.perf_counter() - t0

    def one(nq_):
        t0_ = time.perf_counter()
        r = db.query(queries[:nq_], k, field_of_view=p["qfov"], prop_to_retrieve=p["qpr"])
        return time.perf_counter() - t0_, r

    # the requested number of steps is kept; what shrinks, if the box is slow, is the query sample per step
    t_probe, _ = one(min(qs, 32))
    per_q = t_probe / min(qs, 32)
    total_steps = warmup + steps
    if per_q * qs * total_steps > wall_budget_s:
        qs = int(max(16, min(qs, wall_budget_s / (per_q * total_steps))))
    times = []
    res = None
    for s in range(total_steps):
        dt, res = one(qs)
        if s >= warmup:
            times.append(dt)
    # recall@k of the approximate answer against the exact float64 answer on the same queries
    flat_idx, _flat_dist, counts = res
    exact_idx, _ = ko.exact_knn_numpy(pool, queries[:qs], k)
    off = np.concatenate([[0], np.cumsum(counts)])
    hits = 0
    top1 = 0
    for i in range(qs):
        got = flat_idx[off[i]:off[i + 1]]
        hits += len(set(got.tolist()) & set(exact_idx[i].tolist()))
        top1 += int(len(got) > 0 and got[0] == exact_idx[i, 0])
    db.clear()
    total = float(np.sum(times))
    kk = min(k, ns)
    return {"qps": qs * len(times) / total, "ms_per_step": 1e3 * total / len(times), "steps_done": len(times), "cores": cores,
            "threads": threads, "add_s": t_add, "recall_at_k": hits / float(qs * kk), "top1_recall": top1 / float(qs),
            "queries_per_step": qs, "pool_rows": ns,
            "sample": "reference DCI (oracle/_ref, unmodified dci.c) m=%d L=%d levels=%d build fov=%d p_retr=%g, query fov=%d p_retr=%g (%s); "
                      "pool %d of %d rows x %d float64%s, %d queries/step, k=%d, OpenMP threads in effect=%d on %d cores; "
                      "add() %.1f s and data generation %.1f s are not in value; recall@%d vs exact float64 = %.3f" % (
                          p["m"], p["L"], p["levels"], p["cfov"], p["cpr"], p["qfov"], p["qpr"], p["source"], ns, n, d, note, qs, k,
                          threads, cores, t_add, t_gen, kk, hits / float(qs * kk))}


def cpu_baseline_object(r):
    return {"value": r["qps"], "unit": "queries/s", "cores": r["cores"], "threads": r["threads"], "kind": "reference",
            "sample": r["sample"], "recall_at_k": r["recall_at_k"], "top1_recall": r["top1_recall"], "add_s": r["add_s"],
            "pool_rows": r["pool_rows"], "queries_per_step": r["queries_per_step"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = reference_sample(args.workload, args.steps, args.warmup)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/_dci.so missing (build with `make -C oracle ref` where /root/reference exists)"}))
        return 0
    line = {"impl": "reference", "metric": METRIC, "value": r["qps"], "unit": "queries/s", "n_gpus": args.gpus,
            "steps": r["steps_done"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload),
            "cpu_baseline": cpu_baseline_object(r),
            "recall_at_k": r["recall_at_k"], "add_s": r["add_s"],
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

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, q, d, k, gen, desc = WORKLOADS[args.workload]
    f32_features = GENERATORS[gen][0] == "float32"
    fdt, FT, fbytes = (torch.float32, F32, 4) if f32_features else (torch.float64, F64, 8)
    self_knn = args.workload == "c4"           # the precision/recall metric: self-kNN radii of TWO feature sets per step
    if self_knn and world > 1:
        raise RuntimeError("workload c4 (two-set self-kNN) is measured on one GPU; use --gpus 1")
    per = (n + world - 1) // world
    r0, r1 = min(n, per * rank), min(n, per * (rank + 1))
    need_gb = (r1 - r0) * d * (fbytes + 2) / 1e9 + q * d * (fbytes + 2) / 1e9
    if need_gb > 150.0:
        raise RuntimeError("workload %s needs %.0f GB per GPU at %d GPU(s): use more GPUs (pool rows are sharded)" % (args.workload, need_gb, world))
    pool = synth_rows(args.workload, r0, r1, d, dev, POOL_SEED)
    if self_knn:
        queries = synth_rows(args.workload, 0, q, d, dev, SET_B_SEED)      # the second feature set
    else:
        queries = synth_rows(args.workload, 0, q, d, dev, QUERY_SEED)

    # N > 1: a tie batch rides along: the last row of every shard is duplicated as the first row of the next
    # shard, and the first `world` queries ARE those rows — distance 0 on two shards at once, resolved to the lower index
    n_tie = 0
    if world > 1 and r1 > r0:
        lasts = torch.empty(world, d, device=dev, dtype=fdt)
        dist.all_gather_into_tensor(lasts, pool[-1:].contiguous())
        if rank > 0:
            pool[0] = lasts[rank - 1]
        n_tie = min(world, q)
        queries[:n_tie] = lasts[:n_tie]

    stream = torch.cuda.current_stream()
    lib = load_library()
    kk = min(k, n)
    ix = DeviceKNN(d, local_rank)
    ix.set_stream(stream.cuda_stream)
    # ---- multi-GPU: the library's collective protocol over NVLink peer memory (b200knn_exchange_*); torch.distributed only
    # hands the 64-byte IPC handles round, once ----
    exchange = None
    if world > 1:
        from inclusivegan_b200.dci import PeerExchange
        exchange = PeerExchange(local_rank, rank, world, min(q, 32768), kk, dim=d)
        handles = [None] * world
        dist.all_gather_object(handles, exchange.handle())
        exchange.connect(handles)
        exchange.add_device(ix, pool.data_ptr(), FT, r1 - r0, index_base=r0)      # global centring vector: column sums gathered over peer memory
    else:
        ix.add(pool.data_ptr(), FT, r1 - r0, index_base=r0)
    ix_b = None
    if self_knn:
        ix_b = DeviceKNN(d, local_rank)
        ix_b.set_stream(stream.cuda_stream)
        ix_b.add(queries.data_ptr(), FT, q, index_base=0)
    torch.cuda.synchronize()
    out_i = torch.empty(q, kk, device=dev, dtype=torch.int32)
    out_d = torch.empty(q, kk, device=dev, dtype=torch.float64)
    if self_knn:
        self_i = [torch.empty(n, kk, dtype=torch.int32).pin_memory(), torch.empty(q, kk, dtype=torch.int32).pin_memory()]
        self_d = [torch.empty(n, kk, dtype=torch.float64).pin_memory(), torch.empty(q, kk, dtype=torch.float64).pin_memory()]

    def query_self(handle, oi, od):
        rc = lib.b200knn_query_self(handle._handle, k, 0, ctypes.c_void_p(oi.data_ptr()), ctypes.c_void_p(od.data_ptr()), None)
        if rc != 0:
            raise RuntimeError(lib.b200knn_last_error().decode())

    def step_device():
        if self_knn:      # ManifoldEstimator.__init__ of both sets (precision_recall.py:149-150): rows already on the device
            query_self(ix, self_i[0], self_d[0])
            query_self(ix_b, self_i[1], self_d[1])
        elif world > 1:   # tensor pass on the local shard, bound exchange, globally pruned exact re-rank, list exchange + merge
            exchange.query_device(ix, queries.data_ptr(), FT, q, k, out_i.data_ptr(), out_d.data_ptr())
        else:
            ix.query(queries.data_ptr(), FT, q, k, out_i.data_ptr(), out_d.data_ptr())

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

    units_per_step = (n + q) if self_knn else q          # query rows answered per step
    flops_per_step = 2.0 * d * (float(n) * n + float(q) * q) if self_knn else 2.0 * q * n * d

    # ---- device-resident arm ------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    for h in (ix, ix_b):
        if h is not None:
            h.reset_stats()
            h.set_profiling(True)
    sampler.mark_begin()
    total_ms = timed(step_device, args.steps)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    st = ix.stats()
    if ix_b is not None:
        sb = ix_b.stats()
        st = {key: st[key] + sb[key] for key in st}
    for h in (ix, ix_b):
        if h is not None:
            h.set_profiling(False)
    launches = st["kernel_launches"]

    # ---- end-to-end arm: host-buffer C-ABI call (what DCI.query makes), pinned host queries ----------
    hq = torch.empty(q, d, dtype=fdt).pin_memory()
    hq.copy_(queries)
    torch.cuda.synchronize()
    hx = ctypes.c_void_p()
    ids = (ctypes.c_int * 1)(local_rank)
    assert lib.b200knn_create(d, 1, ids, ctypes.byref(hx)) == 0
    assert lib.b200knn_set_stream(hx, ctypes.c_void_p(stream.cuda_stream)) == 0
    if not self_knn and world == 1:
        assert lib.b200knn_add_device(hx, ctypes.c_void_p(pool.data_ptr()), FT, r1 - r0, d, r0) == 0, lib.b200knn_last_error()
    h_i = torch.empty(q, kk, dtype=torch.int32).pin_memory()
    h_d = torch.empty(q, kk, dtype=torch.float64).pin_memory()
    if self_knn:
        hp = torch.empty(n, d, dtype=fdt).pin_memory()
        hp.copy_(pool)
        torch.cuda.synchronize()

    def check_rc(rc):
        if rc != 0:
            raise RuntimeError(lib.b200knn_last_error().decode())

    def step_e2e():
        if self_knn:
            # per feature set: ManifoldEstimator(features) = add(host rows) + self-kNN radii (precision_recall.py:60-90)
            for rows, nrows, oi, od in ((hp, n, self_i[0], self_d[0]), (hq, q, self_i[1], self_d[1])):
                check_rc(lib.b200knn_clear(hx))
                check_rc(lib.b200knn_add(hx, ctypes.c_void_p(rows.data_ptr()), FT, nrows, d))
                check_rc(lib.b200knn_query_self(hx, k, 0, ctypes.c_void_p(oi.data_ptr()), ctypes.c_void_p(od.data_ptr()), None))
        elif world > 1:
            # b200knn_exchange_query: every rank uploads 1/N of every query chunk from the (same) pinned host matrix and broadcasts
            # the BF16 rows by peer stores; original rows are read from their owner by the re-rank; merged result lands in host memory
            exchange.query_host(ix, hq.data_ptr(), FT, q, k, h_i.data_ptr(), h_d.data_ptr())
        else:
            check_rc(lib.b200knn_query(hx, ctypes.c_void_p(hq.data_ptr()), FT, q, d, k, 0, ctypes.c_void_p(h_i.data_ptr()),
                                       ctypes.c_void_p(h_d.data_ptr()), None))

    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    e2e_ms = timed(step_e2e, e2e_steps)

    # ---- secondaries: the trainer's 24-row calls, and add() from pageable host memory ----
    small_call = None
    add_s = None
    if not self_knn and args.workload not in ("c5", "c5s") and not args.no_secondaries:
        nsc = min(SMALL_CALLS, max(1, q // SMALL_CALL_ROWS))
        hq_np = hq.numpy()
        sc_i = np.empty((SMALL_CALL_ROWS, kk), dtype=np.int32)
        sc_d = np.empty((SMALL_CALL_ROWS, kk), dtype=np.float64)

        def small(i):
            rows = hq_np[i * SMALL_CALL_ROWS:(i + 1) * SMALL_CALL_ROWS]
            if world > 1:
                exchange.query_host(ix, rows.ctypes.data, FT, SMALL_CALL_ROWS, k, sc_i.ctypes.data, sc_d.ctypes.data)
            else:
                check_rc(lib.b200knn_query(hx, ctypes.c_void_p(rows.ctypes.data), FT, SMALL_CALL_ROWS, d, k, 0,
                                           ctypes.c_void_p(sc_i.ctypes.data), ctypes.c_void_p(sc_d.ctypes.data), None))
        for i in range(min(20, nsc)):
            small(i)
        barrier()
        lat = []
        t_all = time.perf_counter()
        for i in range(nsc):
            t0 = time.perf_counter()
            small(i)
            lat.append(time.perf_counter() - t0)
        t_all = time.perf_counter() - t_all
        tl = torch.tensor([t_all], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        small_call = {"rows_per_call": SMALL_CALL_ROWS, "calls": nsc, "latency_ms_median": 1e3 * float(np.median(lat)),
                      "latency_ms_p95": 1e3 * float(np.percentile(lat, 95)), "calls_per_s": nsc / float(tl.item()),
                      "queries_per_s": nsc * SMALL_CALL_ROWS / float(tl.item()),
                      "api": "%s, host rows, one call per %d rows, back to back (training_loop.py:374-403)" % (
                          "b200knn_query" if world == 1 else "b200knn_exchange_query (collective: slice upload, broadcast, bound + list exchange)", SMALL_CALL_ROWS)}
    # ---- the same call from PAGEABLE memory (what a NumPy caller of DCI.query hands over): uploads go through the library's
    # pinned ring, slower than the DMA from page-locked memory, so the cut of the call into chunks matters more ----
    e2e_pageable = None
    if not self_knn and world == 1 and args.workload not in ("c5", "c5s") and not args.no_secondaries:
        q_np = np.array(hq.numpy(), copy=True)                  # fresh pageable allocation
        o_i = np.empty((q, kk), dtype=np.int32)
        o_d = np.empty((q, kk), dtype=np.float64)

        def numpy_call():
            check_rc(lib.b200knn_query(hx, ctypes.c_void_p(q_np.ctypes.data), FT, q, d, k, 0, ctypes.c_void_p(o_i.ctypes.data),
                                       ctypes.c_void_p(o_d.ctypes.data), None))
        numpy_call()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            numpy_call()
            ts.append(time.perf_counter() - t0)
        e2e_pageable = {"ms_per_step": 1e3 * float(np.median(ts)), "value": q / float(np.median(ts)), "unit": "queries/s", "steps": 3,
                        "api": "b200knn_query, pageable NumPy query rows and results (host wall clock, median of 3)"}
        del q_np
    lib.b200knn_destroy(hx)
    if not self_knn and args.workload not in ("c5", "c5s") and (r1 - r0) * d * fbytes < 40e9 and not args.no_secondaries:
        pool_np = pool.cpu().numpy()             # pageable, like the trainer's np.zeros + fill (training_loop.py:358-365)
        ha = DeviceKNN(d, local_rank)
        ts = []
        for _ in range(3):
            ha.clear()
            barrier()
            t0 = time.perf_counter()
            if world > 1:
                exchange.add(ha, pool_np, index_base=r0)
            else:
                check_rc(lib.b200knn_add(ha._handle, ctypes.c_void_p(pool_np.ctypes.data), FT, r1 - r0, d))
            ts.append(time.perf_counter() - t0)
        del ha
        del pool_np
        ta = torch.tensor([min(ts)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        add_s = {"value": float(ta.item()), "unit": "s", "rows_per_rank": r1 - r0, "bytes_per_rank": (r1 - r0) * d * fbytes,
                 "api": "%s from pageable host memory (H2D through the pinned ring + column means + BF16 convert + norms), best of 3, max over ranks" % (
                     "b200knn_add" if world == 1 else "b200knn_exchange_add (each rank its shard; column sums gathered over peer memory)")}

    # ---- self-check of the last device result against a float64 torch brute force on a query subsample ----
    nchk = min(units_per_step if self_knn else q, args.check_queries)

    def brute(sub, base, lo, hi):
        """exact float64 top-kk of `sub` rows against base rows [lo, hi): direct differences, no cancellation"""
        cand_d = torch.full((sub.shape[0], kk), float("inf"), device=dev, dtype=torch.float64)
        cand_i = torch.full((sub.shape[0], kk), -1, device=dev, dtype=torch.int64)
        for s0 in range(0, sub.shape[0], 256):
            sq = sub[s0:s0 + 256].double()
            d2 = torch.empty(sq.shape[0], hi - lo, device=dev, dtype=torch.float64)
            for c0 in range(lo, hi, 16384):
                pc = base[c0:min(c0 + 16384, hi)].double()
                d2[:, c0 - lo:c0 - lo + pc.shape[0]] = (sq * sq).sum(1, keepdim=True) + (pc * pc).sum(1)[None, :] - 2.0 * sq @ pc.T
            kt = min(kk + 4, hi - lo)
            tk = torch.topk(d2, kt, dim=1, largest=False)
            # re-evaluate the shortlist by direct differences (the GEMM form cancels), then order by (distance, index)
            rows = base[(tk.indices + lo).reshape(-1)].double().view(sq.shape[0], kt, -1)
            ex = ((rows - sq[:, None, :]) ** 2).sum(2)
            key = torch.argsort(ex + 0.0, dim=1, stable=True)
            order = torch.gather(tk.indices, 1, key)
            exs = torch.gather(ex, 1, key)
            m = min(kk, kt)
            cand_d[s0:s0 + sq.shape[0], :m] = exs[:, :m]
            cand_i[s0:s0 + sq.shape[0], :m] = order[:, :m] + lo
        return cand_i, cand_d

    def agree(got_i, got_d, ref_i, ref_d):
        """indices equal, or the distance at that rank ties within 1e-6 relative (north_star acceptance)"""
        same = got_i.long() == ref_i
        tie = (got_d - ref_d.sqrt()).abs() <= 1e-6 * ref_d.sqrt().clamp_min(1e-300)
        return bool((same | tie).all().item()) and bool(((got_d - ref_d.sqrt()).abs() <= 1e-5 * ref_d.sqrt().clamp_min(1e-300)).all().item())

    if self_knn:
        sel = torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(3))[:nchk // 2]
        ri, rd = brute(pool[sel], pool, 0, n)
        check = agree(self_i[0].to(dev)[sel], self_d[0].to(dev)[sel], ri, rd)
        ri, rd = brute(queries[sel], queries, 0, q)
        check = check and agree(self_i[1].to(dev)[sel], self_d[1].to(dev)[sel], ri, rd)
    else:
        # a tie batch rides along at N > 1: queries that ARE pool rows sitting on shard boundaries (distance 0 on one
        # shard, duplicates resolved to the lower index by the merge)
        step_device()
        torch.cuda.synchronize()
        sel = torch.randperm(q, device=dev, generator=torch.Generator(device=dev).manual_seed(3))[:nchk]
        sel[:n_tie] = torch.arange(n_tie, device=dev)        # the tie batch is always checked
        ci, cd = brute(queries[sel], pool, 0, r1 - r0)
        ci = torch.where(ci >= 0, ci + r0, ci)
        if world == 1:
            check = agree(out_i[sel], out_d[sel], ci, cd)
        else:
            g_d = torch.empty(world * nchk, kk, device=dev, dtype=torch.float64)
            g_i = torch.empty(world * nchk, kk, device=dev, dtype=torch.int64)
            dist.all_gather_into_tensor(g_d, cd)
            dist.all_gather_into_tensor(g_i, ci)
            torch.cuda.synchronize()
            g_d = g_d.view(world, nchk, kk).permute(1, 0, 2).reshape(nchk, world * kk)
            g_i = g_i.view(world, nchk, kk).permute(1, 0, 2).reshape(nchk, world * kk)
            best = torch.topk(g_d, kk, dim=1, largest=False).indices
            check = agree(out_i[sel], out_d[sel], torch.gather(g_i, 1, best), torch.gather(g_d, 1, best))
            # ties across shards resolve to the LOWER index: query g duplicates the last row of shard g
            want = torch.tensor([min(n, per * (g + 1)) - 1 for g in range(n_tie)], device=dev, dtype=torch.int32)
            check = check and bool((out_i[:n_tie, 0] == want).all().item()) and bool((out_d[:n_tie, 0] == 0).all().item())
            flag = torch.tensor([1 if check else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)      # every rank holds the merged result: all must agree
            check = bool(flag.item())

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
                    ent = json.load(fh).get(args.workload, {})
                # an ncu capture describes ONE shard size: it applies to the GPU count it was taken at only
                if int(ent.get("n_gpus", 1)) == world:
                    traffic = ent.get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": units_per_step / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16 (tensor pass) + f64 (exact re-rank)", "data": "synthetic",
            "config": workload_config(args.workload),
            "run": {"parallelism": ("pool row-sharded x%d, queries replicated; b200knn_exchange_query*: bound exchange + globally pruned exact re-rank, "
                                    "list all-gather by NVLink peer stores + k-way merge kernel (no collective library on the path)" % world)
                    if world > 1 else "single GPU",
                    "l2": "inputs exceed L2 (BF16 pool shard %.2f GB > 126 MB); no explicit flush" % ((r1 - r0) * d * 2 / 1e9),
                    "timing": "CUDA events on the launching stream, max over ranks"},
            "tensor_peak_frac": flops_per_step / (ms_step * 1e-3) / 1e12 / (peaks["bf16_tflops"] * world),
            "roofline": {"bound": "tensor", "kernel": "dist_topc_kernel (tcgen05 BF16 distance GEMM + fused top-C)",
                         "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                         "frac_of_sustained_peak": (ach / peaks["bf16_tflops_sustained"]) if peaks.get("bf16_tflops_sustained") else None,
                         "peak_source": peaks["source"] + ", burst figure", "ms_per_launch": dist_ms, "launches": st["distance_launches"],
                         "flops_per_launch": st["distance_flops"] / max(st["distance_launches"], 1), "traffic": traffic},
            "kernel_ms_per_step": {"convert": st["ms_convert"] / args.steps, "distance": st["ms_distance"] / args.steps,
                                   "rerank": st["ms_rerank"] / args.steps, "second_pass": st["ms_scan"] / args.steps,
                                   "wait_for_peers": st["ms_wait"] / args.steps},
            "uncertified_per_step": st["uncertified"] / args.steps,
            "e2e": {"value": units_per_step / (e2e_ms / e2e_steps * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "h2d_bytes_per_step": (n + q) * d * fbytes if self_knn else q * d * fbytes,
                    "d2h_bytes_per_step": units_per_step * kk * 12,
                    "h2d_bytes_per_rank_per_step": ((n + q) * d * fbytes if self_knn else (q * d * fbytes + world - 1) // world),
                    "api": "per feature set: b200knn_clear + b200knn_add (host rows) + b200knn_query_self (what ManifoldEstimator.__init__ does)" if self_knn else
                           ("b200knn_query (host buffers; the call inclusivegan_b200.dci.DCI.query makes)" if world == 1 else
                            "b200knn_exchange_query (host buffers, collective): per chunk every rank uploads 1/N of the rows from pinned host memory, converts them, "
                            "broadcasts the BF16 rows + norms by peer stores over NVLink; the exact re-rank reads original rows from the rank that uploaded them; "
                            "merged result copied to the host")},
            "add_s": add_s,
            "small_call": small_call,
            "e2e_pageable": e2e_pageable,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "self_check": check, "self_check_queries": int(nchk), "self_check_rule": "top-%d vs torch float64 brute force (direct differences): index equal or distance tie <= 1e-6 rel, distances <= 1e-5 rel" % kk,
        }
        if not args.no_cpu_baseline and world == 1 and args.workload != "small":
            try:
                r = reference_sample(args.workload, steps=2, warmup=1, wall_budget_s=30.0)
                if r is not None:
                    line["cpu_baseline"] = cpu_baseline_object(r)
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
    ap.add_argument("--no-secondaries", action="store_true", help="skip add_s and small_call (development runs)")
    ap.add_argument("--check-queries", type=int, default=2048, help="queries of the last step checked against a float64 brute force")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
