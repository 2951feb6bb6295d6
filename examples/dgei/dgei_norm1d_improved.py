"""Discrete grid exact inference of a normal mean / stdv -- the reference's
examples/dgei/dgei_norm1d_improved.py with the import switched.  Pass a grid size
(e.g. 4096) and observation count (e.g. 100000) for BASELINE config C4."""
import sys
import numpy as np
import scipy.stats
import probayes_b200 as pb

rand_size = int(sys.argv[2]) if len(sys.argv) > 2 else 60
grid = int(sys.argv[1]) if len(sys.argv) > 1 else None
rand_mean = 50.
rand_stdv = 10.
mu_lims = (40, 60)
sigma_lims = (5, 20.)
resolution = {'mu': {grid or 128}, 'sigma': {grid or 192}}

np.random.seed(0)
data = np.random.normal(loc=rand_mean, scale=rand_stdv, size=rand_size)

mu = pb.RV('mu', vtype=float, vset=mu_lims)
sigma = pb.RV('sigma', vtype=float, vset=sigma_lims)
x = pb.RV('x', vtype=float, vset={-np.inf, np.inf})
sigma.set_ufun((np.log, np.exp))
paras = pb.RF(mu, sigma)
stats = pb.RF(x)
model = pb.SD(stats, paras)
model.set_prob(scipy.stats.norm.logpdf, order={'x': 0, 'mu': 'loc', 'sigma': 'scale'},
               pscale='log')
posterior = model({x: data, **resolution}, iid=True, joint=True).conditionalise('x')
post_mean = posterior.marginal('mu')
post_stdv = posterior.marginal('sigma')
post_mean_medn = post_mean.quantile()
post_stdv_medn = post_stdv.quantile()
print("posterior:", posterior.name, posterior.shape)
print("median mu = {:.3f}, median sigma = {:.3f}".format(post_mean_medn['mu'],
                                                       post_stdv_medn['sigma']))
if (grid or 0) <= 512:
    post_expt = posterior.expectation()
    post_expt.pop('x')
    print("expectation:", dict(post_expt))
    print("sum of posterior mass:", posterior.rescaled().prob.sum())
