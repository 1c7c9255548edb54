"""Schedule sweep (development aid): time + DRAM-relevant knobs, first pass only."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200.dci import DeviceKNN, F32, FLAG_NO_CERTIFY
from tools.perf_probe import calibrate
N, Q, d = 240000, 24000, 3072
dev = torch.device("cuda:0"); torch.manual_seed(0)
X = torch.randn(N, d, device=dev); Y = torch.randn(Q, d, device=dev)
oi = torch.empty(Q, 1, device=dev, dtype=torch.int32); od = torch.empty(Q, 1, device=dev, dtype=torch.float64)
def run(cg, budget, sync, reps=4, opt=4):
    os.environ["B200KNN_CTA_GROUP"] = str(cg); os.environ["B200KNN_A_BUDGET_MB"] = str(budget); os.environ["B200KNN_SYNC_TILES"] = str(sync)
    os.environ["B200KNN_OPT"] = str(opt)
    ix = DeviceKNN(d, 0); ix.set_stream(torch.cuda.current_stream().cuda_stream); ix.set_profiling(True); ix.add(X.data_ptr(), F32, N)
    ix.query(Y.data_ptr(), F32, Q, 1, oi.data_ptr(), od.data_ptr(), flags=FLAG_NO_CERTIFY); ix.reset_stats()
    for _ in range(reps): ix.query(Y.data_ptr(), F32, Q, 1, oi.data_ptr(), od.data_ptr(), flags=FLAG_NO_CERTIFY)
    s = ix.stats(); ms = s["ms_distance"] / s["distance_launches"]
    print("cg=%d budget=%3d sync=%2d opt=%d: dist %.2f ms  %.1f TF/s" % (cg, budget, sync, opt, ms, 2.0 * N * Q * d / ms / 1e9)); sys.stdout.flush()
    del ix
if len(sys.argv) > 1:      # single config for ncu
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), reps=1)
else:
    calibrate()
    for rep in range(3):
        run(2, 64, 16, opt=4)
        run(2, 64, 16, opt=0)
        run(1, 60, 16)
    calibrate()
