"""Development aid (run under torchrun, 2 ranks): does NCCL presence / traffic change the distance kernel's time?"""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200.dci import DeviceKNN, F64
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
N, Q, d, k = 150000, 30000, 3072, 1
torch.manual_seed(rank)
X = torch.randn(N, d, device=dev, dtype=torch.float64); Y = torch.randn(Q, d, device=dev, dtype=torch.float64)
oi = torch.empty(Q, k, device=dev, dtype=torch.int32); od = torch.empty(Q, k, device=dev, dtype=torch.float64)
ai = torch.empty(world * Q, k, device=dev, dtype=torch.int32); ad = torch.empty(world * Q, k, device=dev, dtype=torch.float64)
ix = DeviceKNN(d, lr); ix.set_stream(torch.cuda.current_stream().cuda_stream); ix.set_profiling(True); ix.add(X.data_ptr(), F64, N, index_base=rank * N)
junk_in = torch.zeros(1024, device=dev); junk_out = torch.zeros(world * 1024, device=dev)
def phase(name, steps, gather):
    ix.query(Y.data_ptr(), F64, Q, k, oi.data_ptr(), od.data_ptr()); torch.cuda.synchronize(); ix.reset_stats()
    t0 = time.time()
    for _ in range(steps):
        ix.query(Y.data_ptr(), F64, Q, k, oi.data_ptr(), od.data_ptr())
        if gather == 1:
            dist.all_gather_into_tensor(ai, oi); dist.all_gather_into_tensor(ad, od)
        elif gather == 6:
            w1 = dist.all_gather_into_tensor(ai, oi, async_op=True); w2 = dist.all_gather_into_tensor(ad, od, async_op=True); w1.wait(); w2.wait()
        elif gather == 7:
            dist.all_gather_into_tensor(ai, oi); dist.all_gather_into_tensor(ad, od); torch.cuda.current_stream().synchronize()
        elif gather == 2:
            dist.all_gather_into_tensor(junk_out, junk_in)
        elif gather == 3:
            dist.all_gather_into_tensor(ai, oi); dist.all_gather_into_tensor(ad, od); torch.cuda.synchronize()
        elif gather == 4:
            dist.all_gather_into_tensor(ai, oi); dist.all_gather_into_tensor(ad, od); torch.cuda.synchronize(); time.sleep(0.005)
        elif gather == 5:
            torch.cuda.synchronize(); time.sleep(0.005)
    torch.cuda.synchronize(); wall = (time.time() - t0) / steps * 1e3
    s = ix.stats()
    print("rank %d %-34s dist %.2f ms/launch, wall %.2f ms/step" % (rank, name, s["ms_distance"] / s["distance_launches"], wall)); sys.stdout.flush()
phase("before NCCL init", 6, False)
dist.init_process_group("nccl", device_id=dev)
phase("after init, no collectives yet", 6, False)
dist.all_gather_into_tensor(ai, oi); torch.cuda.synchronize()
phase("after first collective, none in loop", 6, False)
phase("all_gather in loop", 6, 1)
phase("junk all_gather in loop", 6, 2)
phase("all_gather + device sync", 6, 3)
phase("all_gather async_op + wait()", 6, 6)
phase("all_gather + stream sync", 6, 7)
dist.barrier(); dist.destroy_process_group()
phase("after destroy", 6, False)
