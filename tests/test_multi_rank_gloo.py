"""world_size-2 gloo test of the N>1 host logic (CPU): row sharding + index_base + all-gather layout + merge
contract, as bench.py wires them (there with NCCL and the CUDA merge kernel; here the per-shard answers come
from the oracle and the merge is its numpy restatement — the check is of the sharding protocol, not of speed)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q, d, k, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import knn_oracle as ko
    from inclusivegan_b200.sharding import shard_range, pad_local_topk
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)                       # every rank regenerates the same global data
    pool = rng.standard_normal((n, d))
    queries = rng.standard_normal((q, d))
    a, b = shard_range(n, world, rank)
    kk = min(k, n)
    li, ld = ko.exact_knn_c(pool[a:b], queries, k)         # shard-local exact answer
    li, ld = pad_local_topk(li + np.int32(a), ld, kk)      # index_base = shard start; pad short shards
    ti, td = torch.from_numpy(li.copy()), torch.from_numpy(ld.copy())
    all_i = torch.empty(world, q, kk, dtype=torch.int32)
    all_d = torch.empty(world, q, kk, dtype=torch.float64)
    dist.all_gather_into_tensor(all_i.view(world * q, kk), ti)
    dist.all_gather_into_tensor(all_d.view(world * q, kk), td)
    mi, md = ko.merge_topk_numpy(all_i.numpy(), all_d.numpy())
    gi, gd = ko.exact_knn_c(pool, queries, k)
    ok = bool(np.array_equal(mi, gi) and np.array_equal(md, gd))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        with open(out_path, "w") as fh:
            fh.write("ok" if int(flag.item()) == 1 else "mismatch")
    dist.destroy_process_group()


@pytest.mark.parametrize("n,q,d,k", [(1001, 37, 24, 5), (5, 9, 8, 4)])
def test_two_rank_shard_gather_merge(tmp_path, n, q, d, k):
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), n, q, d, k, out), nprocs=2, join=True)
    with open(out) as fh:
        assert fh.read() == "ok"
