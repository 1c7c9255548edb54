import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.perf_probe import probe
probe(150000, 30000, 3072, 1, dtype=torch.float64, reps=8)
probe(150000, 30000, 3072, 1, dtype=torch.float64, reps=8)
