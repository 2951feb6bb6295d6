"""One C5-sized Gibbs launch for ncu (d = 64, 65536 chains, 8 sweeps, every sweep recorded with its density (fused))."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from probayes_b200.engine import get_engine
from probayes_b200.cond_cov import CondCov
eng = get_engine(0); rng = np.random.default_rng(0)
d, Cg = 64, 65536
A = rng.standard_normal((d, d)); cov = A @ A.T / d + np.eye(d); mean = rng.standard_normal(d)
cc = CondCov(mean, cov, np.tile([-10., 10.], (d, 1)))
st = eng.to_device(np.tile(mean[:, None], (1, Cg)))
for _ in range(3): eng.gibbs_mvn(st, cc, 8 * d, thin=d, seed=5, want_prob=True)
eng.sync(); print("ok")
