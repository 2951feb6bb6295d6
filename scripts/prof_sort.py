"""One argsort of 2^24 normal keys (profiling target)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probayes_b200.engine import get_engine
eng = get_engine(0)
n = 1 << 24
k = eng.to_device(np.random.default_rng(1).standard_normal(n))
for _ in range(2):
    order, ks = eng.argsort(k, want_keys=True)
p = eng.to_device(np.random.default_rng(2).random(n))
for _ in range(2):
    eng.cumprob(p, False)
torch.cuda.synchronize()
print("ok", int(order[0]))
