"""BASELINE config C3: Bayesian linear regression posterior by batched MH -- the
model of the reference's examples/mcmc/gibbs_linreg.py (priors 28-32, ``norm_reg``
likelihood 35-36) driven as a Metropolis sampler with many chains.  The user-written
``norm_reg`` is recognised by the catalogue by probing it."""
import sys
import numpy as np
import scipy.stats
import probayes_b200 as pb

rand_size = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
n_steps = 1000
rng = np.random.default_rng(2024)
x_obs = rng.normal(0, 1, size=rand_size)
y_obs = rng.normal(1.5 * x_obs - 1., 0.5)

x = pb.RV('x', vtype=float, vset=[-3, 3])
y = pb.RV('y', vtype=float, vset=[-np.inf, np.inf])
beta_0 = pb.RV('beta_0', vtype=float, vset=[-6., 6.], pscale='log')
beta_1 = pb.RV('beta_1', vtype=float, vset=[-6., 6.], pscale='log')
y_sigma = pb.RV('y_sigma', vtype=float, vset=[(0.001), 10.], pscale='log')


def norm_reg(x, y, beta_0, beta_1, y_sigma):
    return scipy.stats.norm.logpdf(y, loc=beta_0 + beta_1 * x, scale=y_sigma)


stats = x & y
paras = beta_0 & beta_1 & y_sigma
process = pb.SP(stats, paras)
process.set_prob(norm_reg, pscale='log')
paras.set_tran(lambda **k: 0.)
paras.set_delta([2.4 * 0.5 / np.sqrt(rand_size)])
process.set_tran(paras)
process.set_delta(paras)
process.set_scores('metropolis')
lr = scipy.stats.linregress(x_obs, y_obs)
init_state = {'beta_0': lr.intercept, 'beta_1': lr.slope, 'y_sigma': 0.5}
sampler = process.sampler(init_state, {'x,y': [x_obs, y_obs]}, stop=n_steps, iid=True,
                          joint=True, chains=chains, thin=10, seed=1)
summary = process(process.walk(sampler))
print("acceptance rate {:.3f}".format(summary.u.rate()))
for k in ('beta_0', 'beta_1', 'y_sigma'):
    print("{}: mean {:.5f}  sd {:.5f}".format(k, summary.v[k][:, 20:].mean(),
                                              summary.v[k][:, 20:].std()))
print("R-hat:", process.rhat(summary.v))
