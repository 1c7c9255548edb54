import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200.dci import DeviceKNN, F32, F64, FLAG_FORCE_SCAN
dev = torch.device("cuda:0")
for (N, Q, d, k, dt, code) in ((60000, 1024, 3072, 1, torch.float64, F64), (60000, 1024, 3072, 1, torch.float32, F32), (20000, 512, 49152, 10, torch.float32, F32), (10000, 256, 512, 100, torch.float64, F64)):
    X = torch.randn(N, d, device=dev, dtype=dt); Y = torch.randn(Q, d, device=dev, dtype=dt)
    oi = torch.empty(Q, k, device=dev, dtype=torch.int32); od = torch.empty(Q, k, device=dev, dtype=torch.float64)
    ix = DeviceKNN(d, 0); ix.set_stream(torch.cuda.current_stream().cuda_stream); ix.add(X.data_ptr(), code, N)
    ix.query(Y.data_ptr(), code, Q, k, oi.data_ptr(), od.data_ptr(), flags=FLAG_FORCE_SCAN); torch.cuda.synchronize()
    t = time.time(); ix.query(Y.data_ptr(), code, Q, k, oi.data_ptr(), od.data_ptr(), flags=FLAG_FORCE_SCAN); torch.cuda.synchronize(); dt_s = time.time() - t
    print("exact scan N=%d Q=%d d=%d k=%d %s: %.1f ms  %.2f T(sub+fma)/s" % (N, Q, d, k, str(dt).split('.')[-1], dt_s * 1e3, N * Q * d / dt_s / 1e12))
    del ix
