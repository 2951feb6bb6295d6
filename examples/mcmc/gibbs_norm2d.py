"""Gibbs sampler for a 2-D Gaussian -- the reference's examples/mcmc/gibbs_norm2d.py
with the import switched."""
import numpy as np
import scipy.stats
import probayes_b200 as pb

lims = (-10., 10.)
n_steps = 2000
means = [0.5, -0.5]
covar = [[1.5, -1.0], [-1.0, 2.]]

x = pb.RV('x', vtype=float, vset=lims)
y = pb.RV('y', vtype=float, vset=lims)
process = pb.SP(x & y)
process.set_prob(scipy.stats.multivariate_normal, means, covar)
process.set_tran(scipy.stats.multivariate_normal, means, covar, tsteps=1)
process.set_scores('gibbs')
sampler = process.sampler({'x': 0., 'y': 1.}, stop=n_steps, seed=0)
samples = [sample for sample in sampler]
summary = process(samples)
n_accept = summary.u.count(True)
inference = summary.v.rescaled()
xvals, yvals, post = inference['x'], inference['y'], inference.prob
print("updates:", n_accept)
print("sample mean:", xvals.mean(), yvals.mean())
print("sample covariance:\n", np.cov(np.stack([xvals, yvals])))
