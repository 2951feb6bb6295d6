"""Small-size exercise of every kernel family (ragged sizes, every variant) -- a quick
crash / launch-error check; compute-sanitizer is not available on the pool."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probayes_b200.engine import get_engine
from probayes_b200.cond_cov import CondCov
eng = get_engine(0)
rng = np.random.default_rng(0)
# K1: warp-specialised (D = 2, 3, 5, 8; ragged chain count, T not a multiple of the batch) + per-thread
for D, C, T in ((2, 70, 45), (3, 33, 29), (5, 40, 23), (8, 35, 17)):
    A = rng.standard_normal((D, D)); cov = A @ A.T / D + np.eye(D)
    st = eng.to_device(np.zeros((D, C)))
    eng.mh_mvn(st, np.zeros(D), cov, T, seed=1, accept="log", thin=2)
    eng.mh_mvn(st, np.ones(D), cov, T, seed=2, accept="reference", per_step=True)
# K2: tiles / stream / suffstat, with extras
N = 5000
x = eng.to_device(rng.normal(0, 1, N)); y = eng.to_device(rng.normal(0, 1, N))
lims = np.array([[-6., 6.], [-6., 6.], [0.001, 10.]]); ex = np.array([[0, 0], [0, 0], [1, 0]]); lg = np.zeros(3, int)
for C, var in ((300, 1), (5, 2), (77, 3)):
    st = eng.to_device(np.tile(np.array([[0.], [0.], [1.]]), (1, C)))
    eng.mh_normreg(st, y, x, 7, lims, ex, lg, [0.05] * 3, seed=3, variant=var, per_step=True, prop_bound=True)
    eng.normreg_logjoint(st, y, x, lims, ex, lg, variant=var)
# K3 / K4
M, S = 37, 53
mu = eng.to_device(np.linspace(-1, 1, M)); sg = eng.to_device(np.linspace(.5, 2, S))
lj = eng.grid_norm_logjoint(y, mu, sg, eng.zeros(M), eng.zeros(S))
eng.grid_norm_logjoint(y, mu, sg, eng.zeros(M), eng.zeros(S), suffstat=True)
eng.grid_conditionalise(lj)
# K5
d = 11
A = rng.standard_normal((d, d)); cov = A @ A.T / d + np.eye(d)
cc = CondCov(np.zeros(d), cov, np.tile([-10., 10.], (d, 1)))
st = eng.to_device(np.zeros((d, 45)))
eng.gibbs_mvn(st, cc, 3 * d + 4, thin=1, seed=5, want_prob=True)
eng.mvn_logpdf(eng.to_device(rng.standard_normal((64, 100))), np.zeros(64), np.eye(64))
# K6
for n in (1, 5000, 70001):
    k = eng.to_device(rng.standard_normal(n))
    o, ks = eng.argsort(k, want_keys=True)
    eng.gather(k, o)
    eng.cumprob(eng.to_device(rng.random(n)), False)
    eng.expectation_sums(eng.to_device(rng.random(n)), False, None, k.reshape(1, -1))
a = eng.to_device(rng.random((13, 29)))
eng.take_axis(a, torch.randperm(13, device=eng.device).to(torch.int32), 0)
eng.take_axis(a, torch.randperm(29, device=eng.device).to(torch.int32), 1)
eng.pd_binary('mul', a, False, eng.to_device(rng.random((1, 29))), True, True)
eng.pd_binary('div', a, False, eng.to_device(rng.random((13, 1))), False, False)
eng.box_sample(np.array([[0., 1.], [1., 2.]]), np.array([0, 1]), 333, seed=4)
eng.sync()
print("sanitize run done")
