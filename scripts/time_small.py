"""Small-problem MH walks (K2 variant 4 vs a launch per step): us per MH step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from probayes_b200.engine import get_engine
eng = get_engine(0); rng = np.random.default_rng(0)
lims = np.array([[-6., 6.], [-6., 6.], [0.001, 10.]]); ex = np.zeros((3, 2), int); lg = np.zeros(3, int)
for N, C, T in [(60, 1, 4000), (1000, 64, 2000), (8192, 64, 1000), (1000, 4096, 500)]:
    x = eng.to_device(rng.normal(0, 1, N)); y = eng.to_device(rng.normal(0, 1, N) * .5 - 1)
    sd = 0.5 / np.sqrt(N)
    for var in (0, 1):
        ms = []
        for _ in range(4):
            st = eng.to_device(np.tile(np.array([[-1.], [1.5], [.5]]), (1, C)))
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(eng.stream)
            eng.mh_normreg(st, y, x, T, lims, ex, lg, [2.4 * sd] * 3, seed=1, variant=var, record=True)
            ev1.record(eng.stream); torch.cuda.synchronize()
            ms.append(ev0.elapsed_time(ev1))
        print("N=%5d C=%5d T=%5d variant %d: %.2f us per MH step" % (N, C, T, var, 1e3 * np.median(ms[1:]) / T))
