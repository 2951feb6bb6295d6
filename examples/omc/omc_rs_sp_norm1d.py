"""Ordinary Monte Carlo random sampling of a 1-D gaussian model -- the reference's
examples/omc/omc_rs_sp_norm1d.py with the import switched (plotting dropped).
Optional arguments: number of samples (default 5000), number of observations (60)."""
import sys
import numpy as np
import scipy.stats
import probayes_b200 as pb

n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
rand_size = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rand_mean = 50.
rand_stdv = 10.
mu_lims = (40, 60)
sigma_lims = (5, 20.)

np.random.seed(0)
data = np.random.normal(loc=rand_mean, scale=rand_stdv, size=rand_size)

mu = pb.RV('mu', vtype=float, vset=mu_lims)
sigma = pb.RV('sigma', vtype=float, vset=sigma_lims)
x = pb.RV('x', vtype=float, vset=[-np.inf, np.inf])
sigma.set_ufun((np.log, np.exp))
paras = pb.RF(mu, sigma)
stats = pb.RF(x)
process = pb.SP(stats, paras)
process.set_prob(scipy.stats.norm.logpdf,
                 order={'x': 0, 'mu': 'loc', 'sigma': 'scale'}, pscale='log')

sampler = process.sampler({'mu': {0}, 'sigma': {0}, 'x': data},
                          iid=True, joint=True, stop=n_samples)
samples = process.walk(sampler)          # or: [sample for sample in sampler]
summary = process(samples)

# the reference rescales to linear probabilities here; beyond a few hundred observations
# those underflow, so normalise in log space first (conditionalise on the data)
inference = summary.conditionalise('x').rescaled()
mu_sort = inference.sorted('mu')
sigma_sort = inference.sorted('sigma')
hat_mu = mu_sort.quantile(0.5)['mu']
hat_sigma = sigma_sort.quantile(0.5)['sigma']
expt = inference.expectation(['mu', 'sigma'])
print("summary:", summary.name, summary.shape)
print("median mu = {:.3f}, median sigma = {:.3f}".format(hat_mu, hat_sigma))
print("E[mu] = {:.3f}, E[sigma] = {:.3f}".format(expt['mu'], expt['sigma']))
