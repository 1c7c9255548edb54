import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.perf_probe import probe
for dt in (torch.float32, torch.float64):
    probe(50000, 50000, 2048, 4, dtype=dt)
    probe(50000, 32768, 2048, 4, dtype=dt)
    probe(50000, 32768, 2048, 1, dtype=dt)
    probe(300000, 30000, 3072, 1, dtype=dt)
