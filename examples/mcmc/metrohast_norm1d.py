"""Metropolis-Hastings for the posterior of a normal mean and stdv -- the reference's
examples/mcmc/metrohast_norm1d.py, unchanged apart from the import and the seed."""
import numpy as np
import scipy.stats
import probayes_b200 as pb

rand_size = 60
rand_mean = 50.
rand_stdv = 10.
n_steps = 5000
step_size = (0.005,)
mu_lims = (40, 60)
sigma_lims = (5, 20.)

np.random.seed(0)
x_obs = np.random.normal(loc=rand_mean, scale=rand_stdv, size=rand_size)

mu = pb.RV('mu', vtype=float, vset=mu_lims, pscale='log')
sigma = pb.RV('sigma', vtype=float, vset=sigma_lims, pscale='log')
x = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
sigma.set_ufun((np.log, np.exp))
paras = pb.RF(mu, sigma)
stats = pb.RF(x)
process = pb.SP(stats, paras)
process.set_prob(scipy.stats.norm.logpdf, order={'x': 0, 'mu': 'loc', 'sigma': 'scale'})
tran = lambda **x: 1.
paras.set_tran((tran, tran))
paras.set_delta(step_size, scale=True)
process.set_tran(paras)
process.set_delta(paras)
process.set_scores('hastings')
process.set_update('metropolis')
init_state = {mu: np.mean(mu_lims), sigma: np.mean(sigma_lims)}
sampler = process.sampler(init_state, {x: x_obs}, stop=n_steps, iid=True, joint=True)
samples = process.walk(sampler)
summary = process(samples)
inference = summary.v.rescaled()
n_accept = summary.u.count(True)
mus, sigmas, post = inference['mu'], inference['sigma'], inference.prob
print("accepted", n_accept, "of", n_steps)
print("median mu = {:.2f}, median sigma = {:.2f}".format(np.median(mus), np.median(sigmas)))
