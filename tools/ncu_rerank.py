import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200.dci import DeviceKNN, F32, F64, FLAG_NO_CERTIFY
dev = torch.device("cuda:0"); N, Q, d = 300000, 30000, 3072
for dt, code in ((torch.float32, F32), (torch.float64, F64)):
    torch.manual_seed(0)
    X = torch.randn(N, d, device=dev, dtype=dt); Y = torch.randn(Q, d, device=dev, dtype=dt)
    oi = torch.empty(Q, 1, device=dev, dtype=torch.int32); od = torch.empty(Q, 1, device=dev, dtype=torch.float64)
    ix = DeviceKNN(d, 0); ix.set_stream(torch.cuda.current_stream().cuda_stream); ix.add(X.data_ptr(), code, N)
    ix.query(Y.data_ptr(), code, Q, 1, oi.data_ptr(), od.data_ptr(), flags=FLAG_NO_CERTIFY); torch.cuda.synchronize()
    del ix, X, Y
