"""Multi-GPU path (b200knn_exchange_*): two processes, two GPUs — skipped on a single-GPU box.

Covers the result exchange alone (all-gather by peer stores + merge) and the whole collective protocol: global
centring at add, BF16 query slices broadcast over peer memory, bound exchange + globally pruned exact re-rank, padded
short shards, the scan path, host- and device-resident queries, several chunks and calls in a row."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(rank, world, port):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # control plane only: hands the IPC handles round
    return torch, dist, torch.device("cuda", rank)


def _finish(dist, torch, rank, ok, msgs, out_path):
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, msgs)
    if rank == 0:
        with open(out_path, "w") as fh:
            fh.write("ok" if int(flag.item()) == 1 else "mismatch: %r" % (gathered,))
    dist.barrier()


def _worker(rank, world, port, out_path):
    torch, dist, dev = _setup(rank, world, port)
    from inclusivegan_b200.dci import DeviceKNN, PeerExchange, F64
    from inclusivegan_b200.sharding import shard_range
    from oracle import knn_oracle as ko
    n, q, d, k = 20011, 700, 160, 5
    rng = np.random.default_rng(5)
    pool = rng.standard_normal((n, d)); queries = rng.standard_normal((q, d))
    a, b = shard_range(n, world, rank)
    tx = torch.from_numpy(pool[a:b]).to(dev); ty = torch.from_numpy(queries).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    ix = DeviceKNN(d, rank); ix.set_stream(st); ix.add(tx.data_ptr(), F64, b - a, index_base=a)
    li = torch.empty(q, k, dtype=torch.int32, device=dev); ld = torch.empty(q, k, dtype=torch.float64, device=dev)
    oi = torch.empty_like(li); od = torch.empty_like(ld)
    ex = PeerExchange(rank, rank, world, q, k)
    handles = [None] * world
    dist.all_gather_object(handles, ex.handle())
    ex.connect(handles)
    ok = True
    msgs = []
    for step in range(5):                                            # several steps: flags and double buffering
        qq = ty if step % 2 == 0 else ty.flip(0).contiguous()
        ix.query(qq.data_ptr(), F64, q, k, li.data_ptr(), ld.data_ptr())
        ex.allgather_merge(li.data_ptr(), ld.data_ptr(), q, k, oi.data_ptr(), od.data_ptr(), st)
        torch.cuda.synchronize()
        ri, rd = ko.exact_knn_c(pool, qq.cpu().numpy(), k)
        good, msg = ko.compare_knn(oi.cpu().numpy(), od.cpu().numpy(), ri, rd, pool, qq.cpu().numpy())
        ok = ok and good
        if not good:
            msgs.append(msg)
    _finish(dist, torch, rank, ok, msgs, out_path)
    ex.close()
    dist.destroy_process_group()


def _protocol_worker(rank, world, port, out_path):
    torch, dist, dev = _setup(rank, world, port)
    from inclusivegan_b200.dci import DeviceKNN, PeerExchange, F32, F64, FLAG_FORCE_SCAN
    from inclusivegan_b200.sharding import shard_range
    from oracle import knn_oracle as ko
    st = torch.cuda.current_stream().cuda_stream
    ok = True
    msgs = []

    def connect(ex):
        handles = [None] * world
        dist.all_gather_object(handles, ex.handle())
        ex.connect(handles)

    def check(tag, gi, gd, pool, queries, k):
        nonlocal ok
        ri, rd = ko.exact_knn_c(pool, queries, k)
        good, msg = ko.compare_knn(gi, gd, ri, rd, pool, queries)
        if not good:
            ok = False
            msgs.append("%s: %s" % (tag, msg))

    # ---- case 1: float64, several chunks per call (max_nq 1024), k = 1 and 5, host and device queries, a tie batch ----
    n, q, d = 30011, 2600, 160
    rng = np.random.default_rng(11)
    pool = rng.standard_normal((n, d)) + 2.0          # a common offset: the GLOBAL mean must be the centring vector
    a, b = shard_range(n, world, rank)
    e0 = shard_range(n, world, 0)[1]
    pool[e0] = pool[e0 - 1]                            # the same row on both sides of the shard boundary
    queries = rng.standard_normal((q, d)) + 2.0
    queries[:4] = pool[[e0 - 1, n - 1, 17, 9000]]      # exact hits, the first one duplicated on two shards
    ix = DeviceKNN(d, rank); ix.set_stream(st)
    ex = PeerExchange(rank, rank, world, 1024, 8, dim=d)
    connect(ex)
    ex.add(ix, pool[a:b], index_base=a)
    for k in (1, 5):
        for rep in range(2):
            qq = queries if rep == 0 else np.ascontiguousarray(queries[::-1])
            gi, gd = ex.query(ix, qq, k)
            check("host f64 k=%d rep=%d" % (k, rep), gi, gd, pool, qq, k)
    boundary = shard_range(n, world, 0)[1] - 1
    gi, gd = ex.query(ix, queries[:4], 1)
    if gi[0, 0] != boundary or gd[0, 0] != 0.0:          # ties across shards resolve to the LOWER index
        ok = False
        msgs.append("tie: got %d (d=%r), want %d" % (gi[0, 0], gd[0, 0], boundary))
    tq = torch.from_numpy(queries).to(dev)
    oi = torch.empty(q, 5, dtype=torch.int32, device=dev); od = torch.empty(q, 5, dtype=torch.float64, device=dev)
    for rep in range(3):
        ex.query_device(ix, tq.data_ptr(), F64, q, 5, oi.data_ptr(), od.data_ptr())
        torch.cuda.synchronize()
        check("device f64 rep=%d" % rep, oi.cpu().numpy(), od.cpu().numpy(), pool, queries, 5)
    # scan path (k > 32) and forced scan through the same exchange
    ex2 = PeerExchange(rank, rank, world, 1024, 40, dim=d)
    connect(ex2)
    ix2 = DeviceKNN(d, rank); ix2.set_stream(st)
    ex2.add(ix2, pool[a:b], index_base=a)
    gi, gd = ex2.query(ix2, queries[:300], 40)
    check("host k=40 (scan)", gi, gd, pool, queries[:300], 40)
    gi, gd = ex2.query(ix2, queries[:300], 3, flags=FLAG_FORCE_SCAN)
    check("host forced scan", gi, gd, pool, queries[:300], 3)
    gi, gd = ex2.query(ix2, queries[:300], 3)
    check("host after scans", gi, gd, pool, queries[:300], 3)

    # ---- case 2: float32 rows, clustered queries (tiny gaps), long rows ----
    n, q, d = 9001, 1500, 1030
    rng = np.random.default_rng(12)
    pool = rng.standard_normal((n, d)).astype(np.float32)
    queries = (pool[rng.integers(0, n, q)] + 0.05 * rng.standard_normal((q, d))).astype(np.float32)
    a, b = shard_range(n, world, rank)
    ix3 = DeviceKNN(d, rank); ix3.set_stream(st)
    ex3 = PeerExchange(rank, rank, world, 4096, 10, dim=d)
    connect(ex3)
    ex3.add(ix3, pool[a:b], index_base=a)
    gi, gd = ex3.query(ix3, queries, 10)
    check("host f32 k=10", gi, gd, pool.astype(np.float64), queries.astype(np.float64), 10)

    # ---- case 3: fewer rows per shard than k (lists padded with -1 before the merge) ----
    n, q, d = 4 * world - 1, 300, 64          # every rank holds 3-4 rows
    pool = rng.standard_normal((n, d)); queries = rng.standard_normal((q, d))
    a, b = shard_range(n, world, rank)
    ix4 = DeviceKNN(d, rank); ix4.set_stream(st)
    ex4 = PeerExchange(rank, rank, world, 512, n + 1, dim=d)
    connect(ex4)
    ex4.add(ix4, pool[a:b], index_base=a)
    gi, gd = ex4.query(ix4, queries, 6)
    check("tiny pool k=6", gi, gd, pool, queries, 6)
    gi, gd = ex4.query(ix4, queries, n + 1)      # k > N: min(k, N) = N columns
    if gi.shape != (q, n):
        ok = False
        msgs.append("k > N: shape %r" % (gi.shape,))
    else:
        check("tiny pool k>N", gi, gd, pool, queries, n + 1)
    # ---- case 4: second-pass lists overflow on every rank (two tight clusters, noise far below BF16 resolution): every
    # rank answers its overflowed queries by the exact scan and the list exchange is repeated, collectively ----
    n, q, d = 4000 * world, 64, 256
    sign = np.where(np.arange(n) % 2 == 0, 1.0, -1.0)[:, None]
    pool = np.ascontiguousarray(40.0 * sign + 1e-3 * rng.standard_normal((n, d)))
    queries = np.ascontiguousarray(40.0 * np.where(np.arange(q) % 2 == 0, 1.0, -1.0)[:, None] + 1e-3 * rng.standard_normal((q, d)))
    a, b = shard_range(n, world, rank)
    ix5 = DeviceKNN(d, rank); ix5.set_stream(st)
    ex5 = PeerExchange(rank, rank, world, 512, 8, dim=d)
    connect(ex5)
    ex5.add(ix5, pool[a:b], index_base=a)
    gi, gd = ex5.query(ix5, queries, 3)
    check("overflow -> collective scan fix-up (host rows)", gi, gd, pool, queries, 3)
    tq5 = torch.from_numpy(queries).to(dev)
    oi5 = torch.empty(q, 3, dtype=torch.int32, device=dev); od5 = torch.empty(q, 3, dtype=torch.float64, device=dev)
    ex5.query_device(ix5, tq5.data_ptr(), F64, q, 3, oi5.data_ptr(), od5.data_ptr())
    torch.cuda.synchronize()
    check("overflow -> collective scan fix-up (device rows)", oi5.cpu().numpy(), od5.cpu().numpy(), pool, queries, 3)
    if ix5.stats()["exact_scanned"] == 0:
        ok = False
        msgs.append("case 4 did not reach the exact scan")
    _finish(dist, torch, rank, ok, msgs, out_path)
    for e in (ex, ex2, ex3, ex4, ex5):
        e.close()
    dist.destroy_process_group()


def _run(worker, native_lib, tmp_path, world=2):
    if native_lib.b200knn_device_count() < world:
        pytest.skip("needs >= %d GPUs" % world)
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    mp.spawn(worker, args=(world, _free_port(), out), nprocs=world, join=True)
    with open(out) as fh:
        assert fh.read() == "ok"


def test_peer_exchange_two_gpus(native_lib, tmp_path):
    _run(_worker, native_lib, tmp_path)


@pytest.mark.parametrize("world", [2, 8])
def test_collective_protocol(native_lib, tmp_path, world):
    _run(_protocol_worker, native_lib, tmp_path, world)
