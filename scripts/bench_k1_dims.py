"""K1 throughput for D = 1..8 (4096 chains x 10^4 steps, identity-ish covariance)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probayes_b200.engine import get_engine
eng = get_engine(0)
C, T = 4096, 10000
rng = np.random.default_rng(0)
for D in (1, 2, 3, 4, 5, 6, 7, 8):
    A = rng.standard_normal((D, D))
    cov = A @ A.T / D + np.eye(D)
    mean = np.zeros(D)
    bufs = {"x": eng.empty(T, D, C), "prob": eng.empty(T, C)}
    ms = []
    for it in range(5):
        st = eng.to_device(np.zeros((D, C)))
        eng.mh_mvn(st, mean, cov, T, seed=it, accept="log", prop_scale=2.4 / np.sqrt(D), out=bufs)
        ms.append(eng.last_kernel_ms())
    m = float(np.median(ms[1:]))
    print("D=%d  %.3f ms  %.3e chain-steps/s  out %.2f TB/s" % (D, m, C * T / m * 1e3,
                                                              (D + 1) * 8 * C * T / m * 1e3 / 1e12))
