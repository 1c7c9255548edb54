"""NVLink peer-memory exchange (b200knn_exchange_*): two processes, two GPUs — skipped on a single-GPU box."""
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


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from inclusivegan_b200.dci import DeviceKNN, PeerExchange, F64
    from inclusivegan_b200.sharding import shard_range
    from oracle import knn_oracle as ko
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # control plane only: hands the IPC handles round
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
    for step in range(5):                                            # several steps: flags and double buffering
        qq = ty if step % 2 == 0 else ty.flip(0).contiguous()
        ix.query(qq.data_ptr(), F64, q, k, li.data_ptr(), ld.data_ptr())
        ex.allgather_merge(li.data_ptr(), ld.data_ptr(), q, k, oi.data_ptr(), od.data_ptr(), st)
        torch.cuda.synchronize()
        ri, rd = ko.exact_knn_c(pool, qq.cpu().numpy(), k)
        good, msg = ko.compare_knn(oi.cpu().numpy(), od.cpu().numpy(), ri, rd, pool, qq.cpu().numpy())
        ok = ok and good
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        with open(out_path, "w") as fh:
            fh.write("ok" if int(flag.item()) == 1 else "mismatch")
    dist.barrier()
    ex.close()
    dist.destroy_process_group()


def test_peer_exchange_two_gpus(native_lib, tmp_path):
    if native_lib.b200knn_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    with open(out) as fh:
        assert fh.read() == "ok"
