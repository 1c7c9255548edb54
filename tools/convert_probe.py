"""Convert/norm kernel bandwidth (development aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200.dci import DeviceKNN, F32, F64
dev = torch.device("cuda:0")
for dtype, code, esz in ((torch.float64, F64, 8), (torch.float32, F32, 4)):
    for (n, d) in ((300000, 3072), (100000, 5000), (50000, 2048)):
        X = torch.randn(n, d, device=dev, dtype=dtype)
        best = 1e9
        for rep in range(5):
            ix = DeviceKNN(d, 0); ix.set_stream(torch.cuda.current_stream().cuda_stream); ix.set_profiling(True)
            ix.add(X.data_ptr(), code, n); s = ix.stats(); best = min(best, s["ms_convert"]); del ix
        by = n * d * (esz + 2) + 8 * n
        print("convert %s %d x %d: %.3f ms  %.2f TB/s (%.1f%% of 6532 GB/s measured copy peak)" % (str(dtype).split('.')[-1], n, d, best, by / best / 1e9, 100 * by / best / 1e9 / 6.532))
