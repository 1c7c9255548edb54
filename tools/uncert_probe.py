"""How often does the first pass fail to certify, by data distribution? (development aid)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200.dci import DeviceKNN, F32
dev = torch.device("cuda:0")
def run(name, X, Y, k):
    N, d = X.shape; Q = Y.shape[0]
    oi = torch.empty(Q, k, device=dev, dtype=torch.int32); od = torch.empty(Q, k, device=dev, dtype=torch.float64)
    ix = DeviceKNN(d, 0); ix.set_stream(torch.cuda.current_stream().cuda_stream); ix.set_profiling(True); ix.add(X.data_ptr(), F32, N)
    ix.query(Y.data_ptr(), F32, Q, k, oi.data_ptr(), od.data_ptr()); ix.reset_stats()
    ix.query(Y.data_ptr(), F32, Q, k, oi.data_ptr(), od.data_ptr()); s = ix.stats()
    print("%-34s N=%d Q=%d d=%d k=%d: uncertified %5.2f%%  scanned %d | dist %.2f rerank %.2f second %.2f ms" % (
        name, N, Q, d, k, 100.0 * s["uncertified"] / Q, s["exact_scanned"], s["ms_distance"], s["ms_rerank"], s["ms_scan"])); sys.stdout.flush()
torch.manual_seed(0)
N, Q, d = 50000, 20000, 2048
g = torch.randn(N, d, device=dev); gq = torch.randn(Q, d, device=dev)
run("gaussian", g, gq, 4)
run("relu(gaussian)", g.clamp_min(0).contiguous(), gq.clamp_min(0).contiguous(), 4)
run("gaussian + 3 (large mean)", (g + 3).contiguous(), (gq + 3).contiguous(), 4)
w = torch.randn(16, d, device=dev) / 4
lat = torch.randn(N, 16, device=dev); latq = torch.randn(Q, 16, device=dev)
run("low-rank(16) relu features", (lat @ w + 0.05 * g).clamp_min(0).contiguous(), (latq @ w + 0.05 * gq).clamp_min(0).contiguous(), 4)
run("low-rank(16) + offset 2", (lat @ w + 0.05 * g + 2).contiguous(), (latq @ w + 0.05 * gq + 2).contiguous(), 4)
img = (0.4 + 0.3 * (lat @ w) + 0.05 * g).clamp(-1, 1).contiguous(); imgq = (0.4 + 0.3 * (latq @ w) + 0.05 * gq).clamp(-1, 1).contiguous()
run("image-like (mean 0.4, lowrank+noise)", img, imgq, 1)
