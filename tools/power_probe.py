"""Clock / power under sustained load for each kernel flavour and for cuBLAS (development aid)."""
import os, sys, threading, time
import torch
import pynvml
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200.dci import DeviceKNN, F32, FLAG_NO_CERTIFY

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)

class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True); self.stop = False; self.rows = []
    def run(self):
        while not self.stop:
            try:
                self.rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                  pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
            except Exception as e:
                self.rows.append((0, 0.0, -1))
            time.sleep(0.01)

def run_sampled(name, fn, secs=2.0, flops_per_call=0.0):
    fn(); torch.cuda.synchronize()
    s = Sampler(); s.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    n = 0; t0 = time.time(); e0.record()
    while time.time() - t0 < secs:
        fn(); n += 1
        if n % 4 == 0: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize(); s.stop = True; s.join()
    ms = e0.elapsed_time(e1) / n
    rows = s.rows[len(s.rows) // 3:]
    clk = sorted(r[0] for r in rows)[len(rows) // 2]; pw = sorted(r[1] for r in rows)[len(rows) // 2]
    reasons = 0
    for r in rows: reasons |= max(r[2], 0)
    print("%-28s %8.2f ms/call %8.1f TF/s | SM clock median %4d MHz, power median %6.1f W, reasons 0x%x, samples %d" % (
        name, ms, flops_per_call / ms / 1e9, clk, pw, reasons, len(rows))); sys.stdout.flush()

def main():
    dev = torch.device("cuda:0")
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=torch.bfloat16); b = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
    run_sampled("cuBLAS bf16 8192^3", lambda: a @ b, flops_per_call=2.0 * n**3)
    N, Q, d = 240000, 24000, 3072
    X = torch.randn(N, d, device=dev); Y = torch.randn(Q, d, device=dev)
    oi = torch.empty(Q, 1, device=dev, dtype=torch.int32); od = torch.empty(Q, 1, device=dev, dtype=torch.float64)
    for opt in (0, 1):
        for cg in (1, 2):
            os.environ["B200KNN_CTA_GROUP"] = str(cg); os.environ["B200KNN_OPT"] = str(opt)
            ix = DeviceKNN(d, 0); ix.set_stream(torch.cuda.current_stream().cuda_stream); ix.add(X.data_ptr(), F32, N)
            run_sampled("knn query cg=%d opt=%d" % (cg, opt), lambda: ix.query(Y.data_ptr(), F32, Q, 1, oi.data_ptr(), od.data_ptr(), flags=FLAG_NO_CERTIFY),
                        flops_per_call=2.0 * N * Q * d)
            del ix
    run_sampled("cuBLAS bf16 8192^3 (again)", lambda: a @ b, flops_per_call=2.0 * n**3)

main()
