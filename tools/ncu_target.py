"""Target for ncu captures: one C2-shaped query per kernel flavour (1-CTA, then 2-CTA)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200.dci import DeviceKNN, F32, FLAG_NO_CERTIFY
N, Q, d = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (240000, 24000, 3072)))
dev = torch.device("cuda:0")
torch.manual_seed(0)
X = torch.randn(N, d, device=dev); Y = torch.randn(Q, d, device=dev)
oi = torch.empty(Q, 1, device=dev, dtype=torch.int32); od = torch.empty(Q, 1, device=dev, dtype=torch.float64)
for cg in (1, 2):
    os.environ["B200KNN_CTA_GROUP"] = str(cg); os.environ.setdefault("B200KNN_OPT", "0")
    ix = DeviceKNN(d, 0)
    ix.set_stream(torch.cuda.current_stream().cuda_stream)
    ix.add(X.data_ptr(), F32, N)
    ix.query(Y.data_ptr(), F32, Q, 1, oi.data_ptr(), od.data_ptr(), flags=FLAG_NO_CERTIFY)
    torch.cuda.synchronize()
    del ix
