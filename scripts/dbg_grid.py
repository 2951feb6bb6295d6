import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from probayes_b200.engine import get_engine
from oracle import np_oracle as o
eng=get_engine(0)
dev=eng.to_device
rng = np.random.default_rng(3)
N, M, S = 3000, 96, 130
data = rng.normal(50., 10., N)
mu = o.uniform_grid(40, 60, M, True, True)
sg = np.exp(o.uniform_grid(np.log(5), np.log(20), S, True, True))
lpm, lps = np.full(M,-np.log(20.)), np.full(S,-np.log(np.log(4.)))
xd, sd, lpsd = dev(data), dev(sg), dev(lps)
ljw = eng.grid_norm_logjoint(xd, dev(mu), sd, dev(lpm), lpsd)
whole = eng.grid_conditionalise(ljw)
eng.sync()
ref = o.grid_norm_logjoint(data, mu, sg, lpm, lps)
print('lj err', np.abs(ljw.cpu().numpy()-ref).max())
pr = o.grid_conditionalise(ref)
pw = whole['post'].cpu().numpy()
print('whole post row0', pw[0,:4], pr[0,:4], 'mask eq', np.array_equal(pw==o.NEARLY_NEGATIVE_INF, pr==o.NEARLY_NEGATIVE_INF))
print('gmax', float(whole['gmax']), ref.max(), 'gsum', float(whole['gsum']), np.exp(ref-ref.max()).sum())
slabs = [eng.grid_norm_logjoint(xd, dev(mu[a:b]), sd, dev(lpm[a:b]), lpsd) for a, b in [(0, 40), (40, 41), (41, 96)]]
eng.sync()
print('slab lj err', np.abs(torch.cat(slabs).cpu().numpy()-ref).max())
gm=[eng.grid_max(s) for s in slabs]; eng.sync(); print([float(g) for g in gm])
gmax = torch.stack(gm).max(dim=0).values
gs=[eng.grid_sumexp(s, gmax) for s in slabs]; eng.sync(); print([float(g) for g in gs])
gsum = torch.stack(gs).sum(dim=0)
parts = [eng.grid_posterior(s, gmax, gsum) for s in slabs]
eng.sync()
post = torch.cat([p[0] for p in parts]).cpu().numpy()
print('slab post row0', post[0,:4], 'mask eq ref', np.array_equal(post==o.NEARLY_NEGATIVE_INF, pr==o.NEARLY_NEGATIVE_INF))
pw2 = whole['post'].cpu().numpy()
print('whole post again row0', pw2[0,:4], np.array_equal(pw,pw2))
