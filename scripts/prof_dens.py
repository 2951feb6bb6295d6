"""One C5-sized batched mvn density (d = 64, 65536 points) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from probayes_b200.engine import get_engine
eng = get_engine(0); rng = np.random.default_rng(0)
d, Cg = 64, 65536
A = rng.standard_normal((d, d)); cov = A @ A.T / d + np.eye(d); mean = rng.standard_normal(d)
xs = eng.to_device(rng.standard_normal((d, Cg)))
for _ in range(3): eng.mvn_logpdf(xs, mean, cov)
eng.sync(); print("ok")
