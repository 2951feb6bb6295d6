"""Runs each secondary kernel a few times at its BASELINE config size so that ncu can
capture one launch of each (see scripts/gpu_prof.sh)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from probayes_b200.engine import get_engine
from probayes_b200.cond_cov import CondCov
eng = get_engine(0)
rng = np.random.default_rng(0)
lims = np.array([[-6., 6.], [-6., 6.], [0.001, 10.]]); ex = np.zeros((3, 2), int); lg = np.zeros(3, int)
which = sys.argv[1:] or ["tiles", "stream", "grid", "gibbs", "k1"]
if "tiles" in which:
    N, C = 1_000_000, 16384
    x = eng.to_device(rng.normal(0, 1, N)); y = eng.to_device(rng.normal(0, 1, N))
    th = eng.to_device(np.stack([rng.normal(-1, .001, C), rng.normal(1.5, .001, C), rng.uniform(.49, .51, C)]))
    for _ in range(3): eng.normreg_logjoint(th, y, x, lims, ex, lg, variant=1)
if "stream" in which:
    Ns = 1 << 27
    xs = torch.randn(Ns, dtype=torch.float64, device=eng.device); ys = torch.randn(Ns, dtype=torch.float64, device=eng.device)
    th = eng.to_device(np.stack([np.full(4, -1.), np.full(4, 1.5), np.full(4, .5)]))
    for _ in range(3): eng.normreg_logjoint(th, ys, xs, lims, ex, lg, variant=2)
    del xs, ys
if "grid" in which:
    Ng, M, S = 20000, 4096, 4096
    data = eng.to_device(rng.normal(50., 10., Ng))
    mu = eng.to_device(np.linspace(40, 60, M + 2)[1:-1]); sg = eng.to_device(np.exp(np.linspace(np.log(5), np.log(20), S + 2)[1:-1]))
    lpm = eng.to_device(np.zeros(M)); lps = eng.to_device(np.zeros(S))
    lj = eng.empty(M, S)
    for _ in range(2):
        eng.grid_norm_logjoint(data, mu, sg, lpm, lps, out=lj)
        eng.grid_conditionalise(lj)
if "gibbs" in which:
    d, Cg = 64, 65536
    A = rng.standard_normal((d, d)); cov = A @ A.T / d + np.eye(d); mean = rng.standard_normal(d)
    cc = CondCov(mean, cov, np.tile([-10., 10.], (d, 1)))
    st = eng.to_device(np.tile(mean[:, None], (1, Cg)))
    for _ in range(2): eng.gibbs_mvn(st, cc, 2 * d, thin=d, seed=5, want_prob=True)
if "k1" in which:
    st = eng.to_device(np.tile(np.array([[0.], [1.]]), (1, 4096)))
    for _ in range(2): eng.mh_mvn(st, [0., 0.], [[2., 1.2], [1.2, 2.]], 10000, seed=3, accept="log")
eng.sync()
print("ok")
