"""Metropolis-Hastings on a 2-D correlated normal -- the reference's
examples/mcmc/mcmc_prob4a.py with ``import probayes_b200 as pb``.  The only change
is the proposal: the reference hands set_delta() a Python lambda that calls
scipy.stats.norm.rvs(); a lambda cannot run in a kernel, so the same N(0, 1)
proposal is given as a frozen scipy distribution.  Add chains=4096 to run the
BASELINE config C2."""
import sys
import numpy as np
import scipy.stats
import probayes_b200 as pb

n_steps = 12288
prop_stdv = np.sqrt(1)
chains = int(sys.argv[1]) if len(sys.argv) > 1 else None


def q(**kwds):
    x, xprime = kwds['x'], kwds["x'"]
    y, yprime = kwds['y'], kwds["y'"]
    return scipy.stats.norm.pdf(yprime, loc=y, scale=prop_stdv) * \
        scipy.stats.norm.pdf(xprime, loc=x, scale=prop_stdv)


x = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
y = pb.RV('y', vtype=float, vset=(-np.inf, np.inf))
process = pb.SP(x & y)
process.set_prob(scipy.stats.multivariate_normal, [0., 0.], [[2.0, 1.2], [1.2, 2.0]])
process.set_tran(q)
process.set_delta(scipy.stats.norm(loc=0., scale=prop_stdv))
process.set_scores('hastings')
process.set_update('metropolis')
sampler = process.sampler({'x': 0., 'y': 1.}, stop=n_steps, chains=chains, seed=0)
samples = process.walk(sampler)
summary = process(samples)
n_accept = summary.u.count(True)
inference = summary.v.rescaled()
xvals, yvals, post = inference['x'], inference['y'], inference.prob
print("accepted {} of {} steps".format(n_accept, n_steps * (chains or 1)))
print("sample covariance:\n", np.cov(np.stack([np.ravel(xvals), np.ravel(yvals)])))
if chains:
    print("R-hat:", process.rhat(summary.v))
