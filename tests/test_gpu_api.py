"""The reference's example workloads through the drop-in public API
(import probayes_b200 as pb) on the GPU, against the golden fixtures generated
from the live reference with the same injected streams.  The set-up code of each
test is the example script's, unchanged except for the delta injection."""
import numpy as np
import pytest
import scipy.stats
from conftest import load_golden, relerr
from gpu_util import engine
import probayes_b200 as pb

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("name", ["mh_mvn_c1", "mh_mvn_log"])
def test_mcmc_prob4a(name):
    """examples/mcmc/mcmc_prob4a.py:38-54."""
    engine()
    g = load_golden(name)
    n_steps = len(g["thresh"])
    x = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
    y = pb.RV('y', vtype=float, vset=(-np.inf, np.inf))
    process = pb.SP(x & y)
    if bool(g["log_pscale"]):
        process.set_prob(scipy.stats.multivariate_normal, [0., 0.], [[2.0, 1.2], [1.2, 2.0]],
                         pscale='log')
    else:
        process.set_prob(scipy.stats.multivariate_normal, [0., 0.], [[2.0, 1.2], [1.2, 2.0]])
    def q(**kwds):                                       # mcmc_prob4a.py:25-30
        x_, xprime = kwds['x'], kwds["x'"]
        y_, yprime = kwds['y'], kwds["y'"]
        return scipy.stats.norm.pdf(yprime, loc=y_, scale=1.) * \
            scipy.stats.norm.pdf(xprime, loc=x_, scale=1.)
    process.set_tran(q)
    process.set_delta(lambda: None)                      # draws are injected below
    process.set_scores('hastings')
    process.set_update('metropolis')
    sampler = process.sampler({'x': 0., 'y': 1.}, stop=n_steps,
                              inj_delta=g["delta"], inj_thresh=g["thresh"])
    samples = [sample for sample in sampler]
    assert len(samples) == n_steps
    summary = process(samples)
    n_accept = summary.u.count(True)
    inference = summary.v.rescaled()
    xvals, yvals, post = inference['x'], inference['y'], inference.prob
    assert n_accept == int(g["u"].sum())
    assert [u is True for u in summary.u] == list(g["u"])
    assert np.abs(xvals - g["x"][:, 0]).max() <= TOL and np.abs(yvals - g["x"][:, 1]).max() <= TOL
    want = np.exp(g["prob"]) if bool(g["log_pscale"]) else g["prob"]
    assert relerr(post, want) <= TOL
    assert summary.v.name == 'x,y' and len(summary.s) == n_steps - 1
    assert relerr(np.array(summary.s), g["s"][1:]) <= TOL
    assert samples[3].v.name.startswith('x=') and samples[3].u in (True, None)
    # proposals (opqr.p) and predecessors (opqr.o): sp.py:244-258, summated sp.py:170-198
    assert summary.p.name == 'x,y' and summary.p.shape == [n_steps]
    assert relerr(np.stack([summary.p['x'], summary.p['y']], 1), g["xprop"]) <= TOL
    assert relerr(summary.p.prob, g["pprop"]) <= TOL
    assert summary.o.shape == [n_steps - 1] and samples[0].o is None
    assert np.array_equal(summary.o['x'], summary.v['x'][:-1])
    assert np.array_equal(summary.o.prob, summary.v.prob[:-1])
    assert samples[5].o['x'] == samples[4].v['x'] and samples[5].p['x'] == summary.p['x'][5]
    walk = process.walk(process.sampler({'x': 0., 'y': 1.}, stop=n_steps,
                                        inj_delta=g["delta"], inj_thresh=g["thresh"]))
    s2 = process(walk)                       # the Walk path builds the same PDs from arrays
    assert np.array_equal(s2.p['y'], summary.p['y']) and np.array_equal(s2.o.prob, summary.o.prob)
    # the proposal-density PD (opqr.q, summated sp.py:170-198): x',y'|x,y with the proposals,
    # their predecessors (initial state first) and the user's q evaluated on them
    for sq in (summary.q, s2.q):
        assert sq.name == str(g["q_name"]) and list(sq.keys()) == list(g["q_keys"])
        assert relerr(sq.prob, g["q_prob"]) <= TOL
        assert relerr(np.stack([sq["x'"], sq["y'"]], 1), g["q_prop"]) <= TOL
        assert np.abs(np.stack([sq['x'], sq['y']], 1) - g["q_pred"]).max() <= TOL
    assert samples[3].q.name.startswith("x'=") and "|x=" in samples[3].q.name
    assert abs(samples[3].q.prob - g["q_prob"][3]) <= TOL * g["q_prob"][3]
    assert summary.r is None


@pytest.mark.parametrize("name,scores,delta", [
    ("mh_norm1d_spherical", 'hastings', (0.005,)), ("mh_norm1d_hastings", 'hastings', [0.005]),
    ("mh_norm1d_metropolis", 'metropolis', [0.005]),
    ("mh_norm1d_bound_open", 'metropolis', [0.6]), ("mh_norm1d_bound_mixed", 'metropolis', [0.6])])
def test_metrohast_norm1d(name, scores, delta):
    """examples/mcmc/metrohast_norm1d.py:23-45 (tuple step = spherical proposal,
    (tran, tran) pair = the e-exponent hastings score)."""
    engine()
    g = load_golden(name)
    n_steps = len(g["thresh"])
    bound = 'bound' in name
    mixed = name.endswith('mixed')       # mu closed (clips), sigma open below / closed above
    mu = pb.RV('mu', vtype=float, vset=[40, 60] if mixed else (40, 60), pscale='log')
    sigma = pb.RV('sigma', vtype=float, vset=[(5,), 20.] if mixed else (5, 20.), pscale='log')
    x = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
    sigma.set_ufun((np.log, np.exp))
    paras = pb.RF(mu, sigma)
    stats = pb.RF(x)
    process = pb.SP(stats, paras)
    process.set_prob(scipy.stats.norm.logpdf, order={'x': 0, 'mu': 'loc', 'sigma': 'scale'})
    tran = lambda **x: 1.
    paras.set_tran((tran, tran) if scores == 'hastings' else tran)
    if bound:                            # variable.py:700-739
        paras.set_delta(delta, scale=True, bound=True)
    else:
        paras.set_delta(delta, scale=True)
    process.set_tran(paras)
    process.set_delta(paras)
    process.set_scores(scores)
    if scores == 'hastings':
        process.set_update('metropolis')
    init_state = {mu: float(g["init"][0]), sigma: float(g["init"][1])}
    sampler = process.sampler(init_state, {x: g["x_obs"]}, stop=n_steps, iid=True, joint=True,
                              inj_delta=g["delta"], inj_thresh=g["thresh"])
    samples = process.walk(sampler)
    summary = process(samples)
    inference = summary.v.rescaled()
    assert summary.u.count(True) == int(g["u"].sum())
    assert [u is True for u in summary.u] == list(g["u"])
    assert relerr(summary.v['mu'], g["x"][:, 0]) <= TOL
    assert relerr(summary.v['sigma'], g["x"][:, 1]) <= TOL
    assert relerr(summary.v.prob, g["prob"]) <= TOL
    assert summary.v.name == 'mu,sigma,x={{{}}}'.format(len(g["x_obs"]) * n_steps)
    assert inference.pscale == 1.
    assert relerr(np.stack([summary.p['mu'], summary.p['sigma']], 1), g["xprop"]) <= TOL
    assert relerr(summary.p.prob, g["pprop"]) <= TOL
    assert summary.p.name == summary.v.name
    assert summary.o.name == 'mu,sigma,x={{{}}}'.format(len(g["x_obs"]) * (n_steps - 1))
    if "r_name" in g.files:              # a (q, r) pair: both proposal-density PDs, keys swapped
        for pd_, tag in ((summary.q, "q"), (summary.r, "r")):
            assert pd_.name == str(g[tag + "_name"]) and list(pd_.keys()) == list(g[tag + "_keys"])
            assert np.array_equal(pd_.prob, g[tag + "_prob"])
            assert relerr(np.stack([pd_["mu'"], pd_["sigma'"]], 1), g[tag + "_prop"]) <= TOL
            assert relerr(np.stack([pd_['mu'], pd_['sigma']], 1), g[tag + "_pred"]) <= TOL
        assert samples[2].r.name.startswith("mu=") and "|mu'=" in samples[2].r.name
    elif scores == 'metropolis':
        assert summary.r is None and summary.q is not None
    # process(samples, conditionalise=True): normalised over the samples (sp.py:194-196)
    cond = process(samples, conditionalise=True).v
    assert cond.name == 'mu,sigma|x={{{}}}'.format(len(g["x_obs"]) * n_steps)
    assert abs(np.exp(cond.prob).sum() - 1.0) <= 1e-12
    lin = np.exp(g["prob"] - g["prob"].max())
    assert relerr(np.exp(cond.prob), lin / lin.sum()) <= 1e-10


def test_linreg_mh_with_user_likelihood():
    """Config C3's model at reference-feasible size (gibbs_linreg.py:28-36 priors and
    ``norm_reg`` likelihood, driven as MH per SURVEY appendix B.5)."""
    engine()
    g = load_golden("mh_linreg")
    n_steps = len(g["thresh"])
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
    paras.set_delta([0.02])
    process.set_tran(paras)
    process.set_delta(paras)
    process.set_scores('metropolis')
    init = g["init"]
    sampler = process.sampler({'beta_0': init[0], 'beta_1': init[1], 'y_sigma': init[2]},
                              {'x,y': [g["x_obs"], g["y_obs"]]}, stop=n_steps, iid=True,
                              joint=True, inj_delta=g["delta"], inj_thresh=g["thresh"])
    summary = process(process.walk(sampler))
    assert [u is True for u in summary.u] == list(g["u"])
    for j, k in enumerate(['beta_0', 'beta_1', 'y_sigma']):
        assert relerr(summary.v[k], g["x"][:, j]) <= TOL
    assert relerr(summary.v.prob, g["prob"]) <= TOL
    # native RNG, many chains, log rule: same model object, new kwargs only
    sampler = process.sampler({'beta_0': -1., 'beta_1': 1.5, 'y_sigma': 0.5},
                              {'x,y': [g["x_obs"], g["y_obs"]]}, stop=400, iid=True, joint=True,
                              chains=256, thin=4, seed=3)
    summary = process(process.walk(sampler))
    assert summary.v['beta_0'].shape == (256, 100) and summary.v.prob.shape == (256, 100)
    assert 0.3 < summary.u.rate() < 0.95
    lr = scipy.stats.linregress(g["x_obs"], g["y_obs"])
    assert abs(summary.v['beta_1'][:, 20:].mean() - lr.slope) < 0.03
    assert all(abs(v - 1) < 0.2 for v in process.rhat(summary.v).values())


@pytest.mark.parametrize("name", ["dgei_small", "dgei_peaked"])
def test_dgei_norm1d_improved(name):
    """examples/dgei/dgei_norm1d_improved.py:20-46."""
    engine()
    g = load_golden(name)
    data = g["data"]
    resolution = {'mu': {len(g["mu"])}, 'sigma': {len(g["sigma"])}}
    mu = pb.RV('mu', vtype=float, vset=(40, 60))
    sigma = pb.RV('sigma', vtype=float, vset=(5, 20.))
    x = pb.RV('x', vtype=float, vset={-np.inf, np.inf})
    sigma.set_ufun((np.log, np.exp))
    paras = pb.RF(mu, sigma)
    stats = pb.RF(x)
    model = pb.SD(stats, paras)
    model.set_prob(scipy.stats.norm.logpdf, order={'x': 0, 'mu': 'loc', 'sigma': 'scale'},
                   pscale='log')
    joint = model({x: data, **resolution}, iid=True, joint=True)
    posterior = joint.conditionalise('x')
    post_expt = posterior.expectation()
    post_expt.pop('x')
    post_mean = posterior.marginal('mu')
    post_stdv = posterior.marginal('sigma')
    post_mean_medn = post_mean.quantile()
    post_stdv_medn = post_stdv.quantile()
    post_prob = posterior.rescaled().prob
    assert joint.name == str(g["joint_name"]) and posterior.name == str(g["post_name"])
    assert post_mean.name == str(g["marg_mu_name"])
    assert joint.prob_device is not None and posterior.prob_device is not None
    assert np.array_equal(np.ravel(posterior['mu']), g["mu"])
    assert np.array_equal(np.ravel(posterior['sigma']), g["sigma"])
    assert relerr(joint.prob, g["joint"]) <= TOL
    atol = TOL * np.abs(g["joint"]).max()          # see tests/test_gpu_grid.py
    clamp = g["posterior"] == pb.NEARLY_NEGATIVE_INF
    assert np.array_equal(posterior.prob == pb.NEARLY_NEGATIVE_INF, clamp)
    assert np.abs(posterior.prob[~clamp] - g["posterior"][~clamp]).max() <= atol
    assert np.abs(post_mean.prob - g["marg_mu"]).max() <= atol
    assert np.abs(post_stdv.prob - g["marg_sigma"]).max() <= atol
    assert np.abs(post_prob - g["post_linear"]).max() <= atol * g["post_linear"].max()
    assert abs(post_expt['mu'] - g["expt_mu"]) <= 1e-9
    assert abs(post_expt['sigma'] - g["expt_sigma"]) <= 1e-9
    assert abs(post_mean_medn['mu'] - g["med_mu"]) <= 1e-9
    assert abs(post_stdv_medn['sigma'] - g["med_sigma"]) <= 1e-9


def test_gibbs_norm2d():
    """examples/mcmc/gibbs_norm2d.py:10-26."""
    engine()
    g = load_golden("gibbs2d")
    lims = (-10., 10.)
    n_steps = len(g["runif"])
    means = [0.5, -0.5]
    covar = [[1.5, -1.0], [-1.0, 2.]]
    x = pb.RV('x', vtype=float, vset=lims)
    y = pb.RV('y', vtype=float, vset=lims)
    process = pb.SP(x & y)
    process.set_prob(scipy.stats.multivariate_normal, means, covar)
    process.set_tran(scipy.stats.multivariate_normal, means, covar, tsteps=1)
    process.set_scores('gibbs')
    sampler = process.sampler({'x': 0., 'y': 1.}, stop=n_steps, inj_thresh=g["runif"])
    samples = [sample for sample in sampler]
    summary = process(samples)
    n_accept = summary.u.count(True)
    inference = summary.v.rescaled()
    xvals, yvals, post = inference['x'], inference['y'], inference.prob
    assert n_accept == n_steps == int(g["n_true"])
    assert np.abs(xvals - g["x"][:, 0]).max() <= TOL
    assert np.abs(yvals - g["x"][:, 1]).max() <= TOL
    assert relerr(post, g["prob"]) <= TOL
    assert process._cond_cov is not None and relerr(process._cond_cov.stdv, g["stdv"]) <= TOL


def test_three_variable_mvn_through_api():
    """SP(x & y & z) with a non-exchangeable mvn: the d > 2 value re-ordering quirk
    (prob.py:349-358) is reproduced by the drop-in (golden from the live reference)."""
    engine()
    g = load_golden("mh_mvn_3d")
    T = len(g["thresh"])
    x = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
    y = pb.RV('y', vtype=float, vset=(-np.inf, np.inf))
    z = pb.RV('z', vtype=float, vset=(-np.inf, np.inf))
    process = pb.SP(x & y & z)
    process.set_prob(scipy.stats.multivariate_normal, g["mean"], g["cov"])
    process.set_tran(lambda **k: 1.)
    process.set_delta(lambda: None)
    process.set_scores('hastings')
    process.set_update('metropolis')
    init = g["init"]
    sampler = process.sampler({'x': init[0], 'y': init[1], 'z': init[2]}, stop=T,
                              inj_delta=g["delta"], inj_thresh=g["thresh"])
    summary = process(process.walk(sampler))
    assert [u is True for u in summary.u] == list(g["u"])
    for j, k in enumerate('xyz'):
        assert np.abs(summary.v[k] - g["x"][:, j]).max() <= TOL
    assert relerr(summary.v.prob, g["prob"]) <= TOL
    # Gibbs on three variables
    g = load_golden("gibbs3d")
    lims = tuple(g["lims"][0])
    rvs = [pb.RV(k, vtype=float, vset=lims) for k in 'xyz']
    process = pb.SP(rvs[0] & rvs[1] & rvs[2])
    process.set_prob(scipy.stats.multivariate_normal, g["mean"], g["cov"])
    process.set_tran(scipy.stats.multivariate_normal, g["mean"], g["cov"], tsteps=1)
    process.set_scores('gibbs')
    sampler = process.sampler({'x': 0., 'y': 1., 'z': -1.}, stop=len(g["runif"]),
                              inj_thresh=g["runif"])
    summary = process(process.walk(sampler))
    for j, k in enumerate('xyz'):
        assert np.abs(summary.v[k] - g["x"][:, j]).max() <= TOL
    assert relerr(summary.v.prob, g["prob"]) <= TOL


def test_batched_c2_through_api_and_host_stream():
    """Config C2 shape (reduced length) through the public API: native RNG,
    chains=4096, both the device-resident and the host-streaming paths."""
    engine()
    x = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
    y = pb.RV('y', vtype=float, vset=(-np.inf, np.inf))
    process = pb.SP(x & y)
    process.set_prob(scipy.stats.multivariate_normal, [0., 0.], [[2.0, 1.2], [1.2, 2.0]])
    process.set_tran(lambda **kwds: 1.)
    process.set_delta(scipy.stats.norm(0., 1.))
    process.set_scores('hastings')
    process.set_update('metropolis')
    a = process(process.walk(process.sampler({'x': 0., 'y': 1.}, stop=600, chains=4096,
                                             seed=9, thin=2)))
    b = process(process.walk(process.sampler({'x': 0., 'y': 1.}, stop=600, chains=4096,
                                             seed=9, thin=2, host_stream=True)))
    assert a.v['x'].shape == (4096, 300)
    assert np.array_equal(a.v['x'], b.v['x']) and np.array_equal(a.v.prob, b.v.prob)
    assert a.u.count(True) == b.u.count(True) and 0.55 < a.u.rate() < 0.68
    flat = np.stack([a.v['x'][:, 100:].ravel(), a.v['y'][:, 100:].ravel()])
    assert np.abs(np.cov(flat) - [[2.0, 1.2], [1.2, 2.0]]).max() < 0.06
    # resume: a second walk on the same sampler continues the chain
    s = process.sampler({'x': 0., 'y': 1.}, stop=600, chains=64, seed=9)
    w1 = process.walk(s, stop=300)
    w2 = process.walk(s, stop=300)
    full = process.walk(process.sampler({'x': 0., 'y': 1.}, stop=600, chains=64, seed=9))
    assert np.array_equal(np.concatenate([w1.arrays['x'], w2.arrays['x']]), full.arrays['x'])


def _prob4a_process():
    x = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
    y = pb.RV('y', vtype=float, vset=(-np.inf, np.inf))
    process = pb.SP(x & y)
    process.set_prob(scipy.stats.multivariate_normal, [0., 0.], [[2.0, 1.2], [1.2, 2.0]])
    process.set_tran(lambda **kwds: 1.)
    process.set_delta(scipy.stats.norm(0., 1.))
    process.set_scores('hastings')
    process.set_update('metropolis')
    return process


@pytest.mark.parametrize("host_stream", [False, True])
def test_sampler_chain0_shards_the_public_api(host_stream):
    """Two 'ranks' of a sharded run through SP.sampler(chain0=) draw different chains
    and their concatenation is the single-process run (Philox is keyed on the global
    chain id): the public-API counterpart of Engine.mh_mvn(chain0=)."""
    engine()
    process = _prob4a_process()

    def run(chains, chain0):
        smp = process.sampler({'x': 0., 'y': 1.}, stop=64, chains=chains, seed=77,
                              chain0=chain0, host_stream=host_stream)
        return process(process.walk(smp))

    full, lo, hi = run(64, 0), run(32, 0), run(32, 32)
    assert not np.array_equal(lo.v['x'], hi.v['x'])
    assert np.array_equal(np.concatenate([lo.v['x'], hi.v['x']]), full.v['x'])
    assert np.array_equal(np.concatenate([lo.v.prob, hi.v.prob]), full.v.prob)
    assert lo.u.count(True) + hi.u.count(True) == full.u.count(True)


def test_mvn_target_with_bounded_delta_through_api():
    """SP(x & y) on bounded variables with set_delta([d], bound=True): x closed (clips),
    y open (bounces) -- golden from the live reference (variable.py:700-739)."""
    engine()
    g = load_golden("mh_mvn_bound")
    T = len(g["thresh"])
    x = pb.RV('x', vtype=float, vset=[-1.5, 1.5])
    y = pb.RV('y', vtype=float, vset=(-1.5, 1.5))
    process = pb.SP(x & y)
    process.set_prob(scipy.stats.multivariate_normal, list(g["mean"]), g["cov"].tolist())
    process.set_tran(lambda **kwds: 1.)
    process.set_delta([float(g["step"])], bound=True)
    process.set_scores('hastings')
    process.set_update('metropolis')
    sampler = process.sampler({'x': g["init"][0], 'y': g["init"][1]}, stop=T,
                              inj_delta=g["delta"], inj_thresh=g["thresh"])
    summary = process(process.walk(sampler))
    assert [u is True for u in summary.u] == list(g["u"])
    assert np.abs(summary.v['x'] - g["x"][:, 0]).max() <= TOL
    assert np.abs(summary.v['y'] - g["x"][:, 1]).max() <= TOL
    assert relerr(summary.v.prob, g["prob"]) <= TOL
    assert relerr(np.stack([summary.p['x'], summary.p['y']], 1), g["xprop"]) <= TOL
    # native RNG, many chains: every sample inside the box
    s2 = process(process.walk(process.sampler({'x': 0.5, 'y': -1.0}, stop=300, chains=512,
                                              seed=5)))
    assert np.abs(s2.v['x']).max() <= 1.5 and np.abs(s2.v['y']).max() < 1.5
    assert 0.5 < s2.u.rate() < 0.98


@pytest.mark.parametrize("name,tsteps", [("gibbs3d_sweep", 3), ("gibbs3d_all", None)])
def test_gibbs_several_coordinates_per_step(name, tsteps):
    """tsteps > 1 / no tsteps (probayes/rf.py:446-452): three coordinate updates -- a whole
    sweep -- per step, three cdf uniforms per step; golden from the live reference."""
    engine()
    g = load_golden(name)
    T = len(g["prob"])
    lims = tuple(g["lims"][0])
    rvs = [pb.RV(k, vtype=float, vset=lims) for k in 'xyz']
    process = pb.SP(rvs[0] & rvs[1] & rvs[2])
    process.set_prob(scipy.stats.multivariate_normal, g["mean"], g["cov"])
    if tsteps is None:
        process.set_tran(scipy.stats.multivariate_normal, g["mean"], g["cov"])
    else:
        process.set_tran(scipy.stats.multivariate_normal, g["mean"], g["cov"], tsteps=tsteps)
    process.set_scores('gibbs')
    sampler = process.sampler({'x': 0., 'y': 1., 'z': -1.}, stop=T, inj_thresh=g["runif"])
    summary = process(process.walk(sampler))
    for j, k in enumerate('xyz'):
        assert np.abs(summary.v[k] - g["x"][:, j]).max() <= TOL
    assert relerr(summary.v.prob, g["prob"]) <= TOL
    assert summary.u.count(True) == T
    # native RNG, batched, thinned: shapes and moments
    s2 = process(process.walk(process.sampler({'x': 0., 'y': 1., 'z': -1.}, stop=400,
                                              chains=256, thin=2, seed=8)))
    assert s2.v['x'].shape == (256, 200)
    flat = np.stack([s2.v[k][:, 20:].ravel() for k in 'xyz'])
    assert np.abs(flat.mean(axis=1) - g["mean"]).max() < 0.05
    assert np.abs(np.cov(flat) - g["cov"]).max() < 0.08


@pytest.mark.parametrize("explicit", [False, True])
def test_omc_rejection_sp_circle(explicit):
    """examples/omc/omc_rejection_sp_circle.py:10-41 -- ordinary Monte Carlo with rejection
    sampling (set_prop, custom scores / thresh / update) -- against the live-reference
    fixture with the same injected uniforms.  The script's Python functions are recognised
    by probing (with a warning); the explicit catalogue specs give the same run."""
    import warnings
    engine()
    g = load_golden("omc_rejection_circle")
    radius = float(g["radius"])
    steps = len(g["u"])

    def inside(x, y):
        return np.array(x**2 + y**2 <= radius**2, dtype=float)

    def norm2d(x, y, loc=0., scale=radius):
        return scipy.stats.norm.pdf(x, loc=loc, scale=scale) * \
            scipy.stats.norm.pdf(y, loc=loc, scale=scale)

    xy_range = [-radius, radius]
    x = pb.RV("x", xy_range)
    y = pb.RV("y", xy_range)
    process = pb.SP(x & y)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        if explicit:
            process.set_prob(pb.catalogue.BallIndicator(radius))
            process.set_prop(pb.catalogue.NormalProduct(0., radius))
        else:
            process.set_prob(inside)
            process.set_prop(norm2d)
        process.set_scores(lambda opqr: opqr.p.prob)
        coef_max = float(norm2d(radius, 1.))
        process.set_thresh(np.random.uniform, low=0., high=coef_max)
        process.set_update(lambda stu: stu.s >= stu.t)
        sampler = process.sampler({0}, stop=steps, inj_unif=g["runif"])
        samples = [sample for sample in sampler]
    assert explicit or sum("BY PROBING" in str(w.message) for w in caught) == 2
    summary = process(samples)
    assert len(samples) == steps
    xy_vals = np.array([(sample.p['x'], sample.p['y']) for sample in samples])
    p_prop = np.array([sample.q.prob for sample in samples])
    accept = np.array([sample.u for sample in samples])
    assert np.array_equal(xy_vals, g["xy"])                  # box draws: bit-exact
    assert np.array_equal(accept, g["u"])
    assert relerr(p_prop, g["q"]) <= TOL
    assert np.array_equal(np.array([s.p.prob for s in samples]), g["p"])
    assert relerr(np.array([s.t for s in samples]), g["t"]) <= TOL
    names = [str(n) for n in g["names"]]
    assert samples[0].p.name == names[0] and samples[0].q.name == names[1]
    assert summary.p.name == names[2] and summary.q.name == names[3]
    assert summary.p.size == int(g["kept_size"])
    assert np.array_equal(summary.p['x'], g["kept_x"]) and np.array_equal(summary.p['y'], g["kept_y"])
    assert np.array_equal(summary.p.prob, g["kept_prob"])
    expectation = summary.p.size / steps
    assert abs(4. * radius**2 * expectation - np.pi * radius**2) < 0.25
    # the Walk path (arrays -> summary without per-step objects) and native RNG at scale
    big = process(process.walk(process.sampler({0}, stop=2_000_000, seed=11)))
    area = 4. * radius**2 * big.p.size / 2_000_000
    assert abs(area - np.pi * radius**2) < 6e-3               # ~ 5 sigma of the MC error
    assert (big.p['x']**2 + big.p['y']**2 <= radius**2).all()


def test_host_buffers_on_the_gibbs_and_likelihood_paths():
    """host_buffers=dict without host_stream: the final device-to-host copy of samples and
    densities goes through pinned buffers kept in the dict; same values as the plain path,
    the buffers are reused by the next walk of the same shape."""
    engine()
    d = 16
    rng = np.random.default_rng(21)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + np.eye(d)
    mean = rng.standard_normal(d)
    rvs = [pb.RV('x%d' % i, vtype=float, vset=(-10., 10.)) for i in range(d)]
    import functools
    process = pb.SP(functools.reduce(lambda a, b: a & b, rvs))
    process.set_prob(scipy.stats.multivariate_normal, mean, cov)
    process.set_tran(scipy.stats.multivariate_normal, mean, cov, tsteps=1)
    process.set_scores('gibbs')
    init = {'x%d' % i: float(mean[i]) for i in range(d)}
    kw = dict(stop=5 * d, chains=200, thin=d, seed=3)
    plain = process(process.walk(process.sampler(init, **kw)))
    bufs = {}
    a = process(process.walk(process.sampler(init, host_buffers=bufs, **kw)))
    assert len(bufs) == 2 and all(t.is_pinned() for t in bufs.values())
    for k in init:
        assert np.array_equal(a.v[k], plain.v[k])
    assert np.array_equal(a.v.prob, plain.v.prob)
    ptrs = sorted(t.data_ptr() for t in bufs.values())
    keep = {k: a.v[k].copy() for k in init}
    b = process(process.walk(process.sampler(init, host_buffers=bufs, stop=5 * d, chains=200,
                                             thin=d, seed=4)))
    assert sorted(t.data_ptr() for t in bufs.values()) == ptrs          # reused, not regrown
    assert not np.array_equal(b.v['x0'], keep['x0'])                      # a different walk


def test_device_backed_pds_serialise(monkeypatch):
    """Row f4: the DGEI results (device-backed PDs) through serialise / the HDF5 writer
    (h5py stand-in, tests/fake_h5py.py): names, keys and the stored structure as the live
    reference's for the same model (tests/golden/pd_serialise.npz)."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import fake_h5py
    monkeypatch.setitem(sys.modules, 'h5py', fake_h5py)
    engine()
    g = load_golden("pd_serialise")
    mu = pb.RV('mu', vtype=float, vset=(40, 60))
    sigma = pb.RV('sigma', vtype=float, vset=(5, 20.))
    x = pb.RV('x', vtype=float, vset={-np.inf, np.inf})
    sigma.set_ufun((np.log, np.exp))
    model = pb.SD(pb.RF(x), pb.RF(mu, sigma))
    model.set_prob(scipy.stats.norm.logpdf, order={'x': 0, 'mu': 'loc', 'sigma': 'scale'},
                   pscale='log')
    joint = model({x: g["data"], 'mu': {len(g["mu"])}, 'sigma': {len(g["sigma"])}}, iid=True,
                  joint=True)
    posterior = joint.conditionalise('x')
    post_mu = posterior.marginal('mu').rescaled()
    assert joint.prob_device is not None
    for tag, pd in (("joint", joint), ("posterior", posterior), ("post_mu", post_mu)):
        (name, d), = pb.serialise(pd).items()
        assert name == str(g[tag + "_name"])
        assert list(d.keys()) == json.loads(str(g[tag + "_keys"]))
        assert {k: v for k, v in d['attrs'].items()} == json.loads(str(g[tag + "_dims"]))
        pb.write_dist("gpu_" + tag, pd)
        want = json.loads(str(g[tag + "_file"]))
        got = fake_h5py.dump("gpu_" + tag)
        assert list(got.keys()) == [k for k in want if ' ' not in k]
        for key, (kind, shape, vals, attrs) in got[name].items():
            wk, ws, wv, wa = want[name][key]
            assert kind == wk and list(shape) == ws and attrs == wa
            if vals is not None and key != 'pscale':
                tol = 1e-12 * max(1.0, np.abs(g["joint_prob"]).max())
                assert np.abs(np.asarray(vals, dtype=float) - np.asarray(wv, dtype=float)).max() <= tol
        back, = pb.read_dist("gpu_" + tag)
        assert back.name == str(g[tag + "_back_name"])
