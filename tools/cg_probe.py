import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.perf_probe import probe, calibrate
calibrate()
for budget in (40, 64):
    os.environ["B200KNN_A_BUDGET_MB"] = str(budget)
    for opt in (4, 0):
        os.environ["B200KNN_OPT"] = str(opt)
        print("A_BUDGET_MB", budget, "OPT", opt)
        for cg in (1, 2):
            probe(240000, 24000, 3072, 1, cg=cg, reps=3)
calibrate()
os.environ["B200KNN_OPT"] = "4"; os.environ["B200KNN_A_BUDGET_MB"] = "40"
probe(300000, 30000, 3072, 1, cg=1)
probe(300000, 30000, 3072, 1, cg=2)
probe(50000, 50000, 2048, 4, cg=2)
probe(50000, 50000, 2048, 4, cg=1)
probe(300000, 24, 3072, 1, reps=10)
