"""Target for ncu captures of convert_norm_kernel: one pool-sized (300k x 3072) and one query-sized (30k x 3072) conversion,
float64 then float32."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200.dci import DeviceKNN, F32, F64
dev = torch.device("cuda:0")
for dtype, code in ((torch.float64, F64), (torch.float32, F32)):
    X = torch.randn(300000, 3072, device=dev, dtype=dtype)
    Y = torch.randn(30000, 3072, device=dev, dtype=dtype)
    oi = torch.empty(30000, 1, device=dev, dtype=torch.int32); od = torch.empty(30000, 1, device=dev, dtype=torch.float64)
    ix = DeviceKNN(3072, 0); ix.set_stream(torch.cuda.current_stream().cuda_stream)
    ix.add(X.data_ptr(), code, 300000)
    ix.query(Y.data_ptr(), code, 30000, 1, oi.data_ptr(), od.data_ptr())
    torch.cuda.synchronize()
    del ix, X, Y
