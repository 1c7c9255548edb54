"""Perf probe (development aid): device-resident pool/queries via torch, per-kernel CUDA-event times."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200.dci import DeviceKNN, F32, F64

def calibrate(n=8192, reps=10):
    """cuBLAS bf16 GEMM on this very box: the yardstick the kernel numbers of this run are read against."""
    a = torch.randn(n, n, device="cuda", dtype=torch.bfloat16); b = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): a @ b
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(60): a @ b
    e1.record(); torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / 60
    print("calibration: cuBLAS bf16 %d^3 burst %.1f TF/s, sustained(60 back-to-back) %.1f TF/s" % (n, 2 * n**3 / best / 1e9, 2 * n**3 / sus / 1e9)); sys.stdout.flush()


def probe(N, Q, d, k, dtype=torch.float32, reps=3, kind="gauss", cg=None):
    if cg: os.environ["B200KNN_CTA_GROUP"] = str(cg)
    else: os.environ.pop("B200KNN_CTA_GROUP", None)
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    X = torch.randn(N, d, device=dev, dtype=dtype)
    if kind == "cluster":
        sel = torch.randint(0, N, (Q,), device=dev)
        Y = X[sel] + 0.1 * torch.randn(Q, d, device=dev, dtype=dtype)
    else:
        Y = torch.randn(Q, d, device=dev, dtype=dtype)
    code = F32 if dtype == torch.float32 else F64
    ix = DeviceKNN(d, 0)
    st = torch.cuda.current_stream()
    ix.set_stream(st.cuda_stream)
    ix.set_profiling(True)
    t0 = time.time(); ix.add(X.data_ptr(), code, N); torch.cuda.synchronize(); t_add = time.time() - t0
    oi = torch.empty(Q, k, device=dev, dtype=torch.int32); od = torch.empty(Q, k, device=dev, dtype=torch.float64)
    ix.query(Y.data_ptr(), code, Q, k, oi.data_ptr(), od.data_ptr()); torch.cuda.synchronize()
    s0 = ix.stats(); ix.reset_stats()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ix.query(Y.data_ptr(), code, Q, k, oi.data_ptr(), od.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    s = ix.stats()
    fl = 2.0 * Q * N * d
    md = s["ms_distance"] / max(s["distance_launches"], 1)
    print("cg=%s N=%d Q=%d d=%d k=%d %s %s: add %.3fs (convert %.2f ms) | query %.2f ms = %.0f q/s | dist %.2f ms/launch x%d = %.1f TF/s (%.1f%% of 1676.7) | convert %.2f rerank %.2f scan %.2f ms/q-call | uncert %d" % (
        cg, N, Q, d, k, str(dtype).split('.')[-1], kind, t_add, s0["ms_convert"], ms, Q / ms * 1e3, md, s["distance_launches"] // reps,
        s["distance_flops"] / max(s["ms_distance"], 1e-9) / 1e9, 100 * s["distance_flops"] / max(s["ms_distance"], 1e-9) / 1e9 / 1676.7,
        s["ms_convert"] / reps, s["ms_rerank"] / reps, s["ms_scan"] / reps, s["uncertified"] // reps))
    if kind == "cluster":
        print("   cluster top-1 hit rate:", float((oi[:, 0].long() == sel).float().mean()))
    sys.stdout.flush()
    del ix

if __name__ == "__main__":
    calibrate()
    probe(60000, 8192, 49152, 10)            # config-5 feature shape, reduced N and Q
    probe(300000, 30000, 3072, 1, dtype=torch.float64)
    probe(240000, 24000, 3072, 1, dtype=torch.float64)
    probe(50000, 50000, 2048, 4)
    probe(10000, 100, 5000, 10, dtype=torch.float64)
    calibrate()
