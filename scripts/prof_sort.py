"""K6 profiling target: argsort of 2^24 normal keys, cumulative probability and
expectation sums of 2^24 log-pscale cells, run twice (first round = warm-up)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probayes_b200.engine import get_engine
eng = get_engine(0)
n = 1 << 24
k = eng.to_device(np.random.default_rng(1).standard_normal(n))
lp = eng.to_device(np.log(np.random.default_rng(2).random(n)) - 100.0)
vals = eng.to_device(np.random.default_rng(3).random((2, n)))
for _ in range(2):
    order, ks = eng.argsort(k, want_keys=True)
    eng.cumprob(lp, True)
    eng.expectation_sums(lp, True, None, vals)
torch.cuda.synchronize()
print("ok", int(order[0]))
