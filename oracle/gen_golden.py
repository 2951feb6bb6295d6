"""Generates tests/golden/*.npz by driving the LIVE reference through its public
API on seeded inputs with injected proposal / threshold streams.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the development container:

    python -m oracle.gen_golden            # writes tests/golden/*.npz

The reference's own tests hold no vectors for this path,
so these fixtures -- outputs of the reference itself -- are what pins the oracle
and, through it, the CUDA path.  Sizes are kept small (a few hundred steps, a
few thousand grid cells) so the fixtures stay a few hundred KB.
"""
import os
import sys
import numpy as np
import scipy.stats

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden")


def _summary_arrays(sp, samples, keys):
    summary = sp(samples)
    x = np.stack([np.asarray(summary.v[k], dtype=float) for k in keys], axis=-1)
    s = np.array([np.nan] + [float(v) for v in summary.s])  # s is None on step 1
    t = np.array([float(v) for v in summary.t])
    u = np.array([v is True for v in summary.u])
    xp = np.stack([np.asarray(summary.p[k], dtype=float) for k in keys], axis=-1)
    out = dict(x=x, prob=np.asarray(summary.v.prob, dtype=float), s=s, t=t, u=u,
               xprop=xp, pprop=np.asarray(summary.p.prob, dtype=float))
    if summary.q is not None:
        # the proposal-density PDs (sp.py:170-198): "x',y'|x,y" with the proposals, their
        # predecessors (the initial state first) and q's value per step
        q = summary.q
        out.update(q_name=np.array(q.name), q_keys=np.array(list(q.keys())),
                   q_prob=np.asarray(q.prob, dtype=float),
                   q_pred=np.stack([np.asarray(q[k], dtype=float) for k in keys], axis=-1),
                   q_prop=np.stack([np.asarray(q[k + "'"], dtype=float) for k in keys], axis=-1),
                   q_step3_name=np.array(samples[3].q.name))
    if summary.r is not None:               # (q, r) pairs: the reverse density, keys swapped
        r = summary.r
        out.update(r_name=np.array(r.name), r_keys=np.array(list(r.keys())),
                   r_prob=np.asarray(r.prob, dtype=float),
                   r_pred=np.stack([np.asarray(r[k], dtype=float) for k in keys], axis=-1),
                   r_prop=np.stack([np.asarray(r[k + "'"], dtype=float) for k in keys], axis=-1))
    return out


# ---------------------------------------------------------------------------
def mh_mvn(seed, T, init, log_pscale=False, cov_tran=None):
    """examples/mcmc/mcmc_prob4a.py:38-51 with injected streams.
    cov_tran: use ``set_tran(covariance ndarray)`` (probayes/rf.py:209-220)
    instead of the explicit symmetric q, so the injected delta is coloured by
    the Cholesky factor."""
    pb = ref_shim.load()
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((T, 2))
    U = rng.random(T)
    zi, ui = iter(Z), iter(U)
    mean, cov = [0., 0.], [[2.0, 1.2], [1.2, 2.0]]

    def q(**kwds):
        x, xprime = kwds['x'], kwds["x'"]
        y, yprime = kwds['y'], kwds["y'"]
        return scipy.stats.norm.pdf(yprime, loc=y, scale=1.) * \
            scipy.stats.norm.pdf(xprime, loc=x, scale=1.)

    x = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
    y = pb.RV('y', vtype=float, vset=(-np.inf, np.inf))
    sp = pb.SP(x & y)
    if log_pscale:
        sp.set_prob(scipy.stats.multivariate_normal, mean, cov, pscale='log')
    else:
        sp.set_prob(scipy.stats.multivariate_normal, mean, cov)
    if cov_tran is None:
        sp.set_tran(q)
    else:
        sp.set_tran(np.array(cov_tran))
    sp.set_delta(lambda: (lambda z: sp.Delta(x=z[0], y=z[1]))(next(zi)))
    sp.set_scores('hastings')
    sp.set_update('metropolis')
    sp.set_thresh(lambda: next(ui))        # after set_scores (sp.py:64-65)
    sampler = sp.sampler({'x': init[0], 'y': init[1]}, stop=T)
    samples = [s for s in sampler]
    out = _summary_arrays(sp, samples, ['x', 'y'])
    out.update(delta=Z, thresh=U, init=np.array(init, dtype=float),
               mean=np.array(mean), cov=np.array(cov),
               log_pscale=np.array(log_pscale),
               cov_tran=np.array(cov_tran if cov_tran is not None else np.eye(2)),
               has_cov_tran=np.array(cov_tran is not None))
    return out


def mh_mvn_bound(seed, T):
    """mcmc_prob4a's 2-D normal target on BOUNDED variables with set_delta([d], bound=True)
    (probayes/variable.py:700-739): x closed (proposals clip at the limits), y open
    (proposals beyond a limit bounce back to the current value).  Uniforms injected."""
    pb = ref_shim.load()
    rng = np.random.default_rng(seed)
    mean, cov = [0., 0.], [[2.0, 1.2], [1.2, 2.0]]
    x = pb.RV('x', vtype=float, vset=[-1.5, 1.5])
    y = pb.RV('y', vtype=float, vset=(-1.5, 1.5))
    sp = pb.SP(x & y)
    sp.set_prob(scipy.stats.multivariate_normal, mean, cov)
    sp.set_tran(lambda **kwds: 1.)
    step = 0.9
    sp.set_delta([step], bound=True)
    sp.set_scores('hastings')
    sp.set_update('metropolis')
    R = rng.random((T, 3))               # per step: d_x, d_y, threshold
    init = (0.5, -1.0)
    with ref_shim.injected_uniform(R.ravel()):
        sampler = sp.sampler({'x': init[0], 'y': init[1]}, stop=T)
        samples = [s for s in sampler]
    out = _summary_arrays(sp, samples, ['x', 'y'])
    out.update(delta=-step + (2 * step) * R[:, :2], thresh=R[:, 2], init=np.array(init),
               mean=np.array(mean), cov=np.array(cov), step=np.array(step),
               lims=np.array([[-1.5, 1.5], [-1.5, 1.5]]),
               ex=np.array([[False, False], [True, True]]))
    return out


# ---------------------------------------------------------------------------
def mh_mvn3d(seed, T):
    """Three variables, non-exchangeable mean / covariance: pins the value
    re-ordering of probayes/prob.py:349-358 for d > 2 ([v1, v0, v2])."""
    pb = ref_shim.load()
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((T, 3)) * 0.7
    U = rng.random(T)
    zi, ui = iter(Z), iter(U)
    mean = [0.3, -1.0, 2.0]
    A = rng.standard_normal((3, 3))
    cov = A @ A.T / 3 + np.diag([0.5, 1.0, 2.0])
    x = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
    y = pb.RV('y', vtype=float, vset=(-np.inf, np.inf))
    z = pb.RV('z', vtype=float, vset=(-np.inf, np.inf))
    sp = pb.SP(x & y & z)
    sp.set_prob(scipy.stats.multivariate_normal, mean, cov)
    sp.set_tran(lambda **k: 1.)
    sp.set_delta(lambda: (lambda d: sp.Delta(x=d[0], y=d[1], z=d[2]))(next(zi)))
    sp.set_scores('hastings')
    sp.set_update('metropolis')
    sp.set_thresh(lambda: next(ui))
    init = (0.5, -0.5, 1.5)
    sampler = sp.sampler({'x': init[0], 'y': init[1], 'z': init[2]}, stop=T)
    samples = [s for s in sampler]
    out = _summary_arrays(sp, samples, ['x', 'y', 'z'])
    out.update(delta=Z, thresh=U, init=np.array(init), mean=np.array(mean), cov=cov,
               log_pscale=np.array(False))
    return out


def gibbs3d(seed, T, tsteps=1):
    """Gibbs on a 3-D mvn (tsteps=1): conditional draws in natural order, recorded
    density on the permuted point.  tsteps=3 / None: a whole sweep per step
    (probayes/rf.py:446-452), three uniforms per step."""
    pb = ref_shim.load()
    rng = np.random.default_rng(seed)
    lims = (-8., 8.)
    means = [0.5, -0.5, 1.0]
    A = rng.standard_normal((3, 3))
    covar = A @ A.T / 3 + np.diag([0.6, 1.0, 1.5])
    x = pb.RV('x', vtype=float, vset=lims)
    y = pb.RV('y', vtype=float, vset=lims)
    z = pb.RV('z', vtype=float, vset=lims)
    sp = pb.SP(x & y & z)
    sp.set_prob(scipy.stats.multivariate_normal, means, covar)
    if tsteps is None:
        sp.set_tran(scipy.stats.multivariate_normal, means, covar)
    else:
        sp.set_tran(scipy.stats.multivariate_normal, means, covar, tsteps=tsteps)
    sp.set_scores('gibbs')
    R = rng.random(T * (3 if tsteps in (None, 3) else tsteps))
    with ref_shim.injected_uniform(R):
        sampler = sp.sampler({'x': 0., 'y': 1., 'z': -1.}, stop=T)
        samples = [s for s in sampler]
    summary = sp(samples)
    xs = np.stack([np.asarray(summary.v[k], float) for k in 'xyz'], axis=-1)
    return dict(x=xs, prob=np.asarray(summary.v.prob, float), runif=R,
                init=np.array([0., 1., -1.]), mean=np.array(means), cov=covar,
                lims=np.array([lims] * 3))


# ---------------------------------------------------------------------------
def mh_norm1d(seed, T, N, scores, spherical=False, bound=None):
    """examples/mcmc/metrohast_norm1d.py:23-42 ((mu, sigma) posterior, log
    pscale, sigma with (np.log, np.exp) ufun, iid+joint), uniforms injected by
    patching np.random.uniform.  scores='hastings' keeps the script's
    asymmetric (tran, tran) pair -> the e-exponent score of SURVEY A.4;
    scores='metropolis' is the plain ratio."""
    pb = ref_shim.load()
    rng = np.random.default_rng(seed)
    x_obs = rng.normal(50., 10., size=N)
    mu_lims, sigma_lims = (40, 60), (5, 20.)
    step = 0.005
    ex = [[True, True], [True, True]]
    init_vals = (50., 12.5)
    if bound == 'open':             # bound=True, both ends open: bounce back
        step, init_vals = 0.6, (58.5, 18.)
    elif bound == 'mixed':          # mu closed (clip), sigma open below / closed above
        step, init_vals = 0.6, (41., 6.)
        mu_lims, sigma_lims = [40, 60], [(5,), 20.]
        ex = [[False, False], [True, False]]
    mu = pb.RV('mu', vtype=float, vset=mu_lims, pscale='log')
    sigma = pb.RV('sigma', vtype=float, vset=sigma_lims, pscale='log')
    x = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
    sigma.set_ufun((np.log, np.exp))
    paras = pb.RF(mu, sigma)
    stats = pb.RF(x)
    sp = pb.SP(stats, paras)
    sp.set_prob(scipy.stats.norm.logpdf,
                order={'x': 0, 'mu': 'loc', 'sigma': 'scale'})
    tran = lambda **x: 1.
    if scores == 'hastings':
        paras.set_tran((tran, tran))
    else:
        paras.set_tran(tran)
    if bound:
        paras.set_delta([step], scale=True, bound=True)
    else:
        paras.set_delta((step,) if spherical else [step], scale=True)
    sp.set_tran(paras)
    sp.set_delta(paras)
    sp.set_scores(scores)
    if scores == 'hastings':
        sp.set_update('metropolis')
    R = rng.random((T, 3))               # per step: d_mu, d_sigma, threshold
    init = {mu: init_vals[0], sigma: init_vals[1]}
    with ref_shim.injected_uniform(R.ravel()):
        sampler = sp.sampler(init, {x: x_obs}, stop=T, iid=True, joint=True)
        samples = [s for s in sampler]
    out = _summary_arrays(sp, samples, ['mu', 'sigma'])
    lengths = np.array([20., np.log(20.) - np.log(5.)])
    dmax = step * lengths                # scale=True: probayes/field.py:299-303
    delta = -dmax + (dmax - (-dmax)) * R[:, :2]
    radius = 0.0
    if spherical:                        # probayes/field.py:509-531
        radius = step * np.sqrt(np.sum(lengths ** 2))
        cube = -radius + (2 * radius) * R[:, :2]
        delta = (cube * radius) / np.sqrt(np.sum(cube ** 2, axis=1, keepdims=True)) * lengths
    out.update(radius=np.array(radius), lengths=lengths, spherical=np.array(spherical))
    out.update(x_obs=x_obs, delta=delta, thresh=R[:, 2], runif=R,
               init=np.array(init_vals), dmax=dmax,
               lims=np.array([[40., 60.], [5., 20.]]),
               ex=np.array(ex), bound=np.array(bool(bound)),
               log_ufun=np.array([False, True]),
               coef=np.array(np.e if scores == 'hastings' else 1.0))
    return out


# ---------------------------------------------------------------------------
def mh_linreg(seed, T, N, dstep=0.02):
    """Config C3's model at reference-feasible size: priors of
    examples/mcmc/gibbs_linreg.py:28-32, likelihood 35-36, driven as an MH
    sampler per SURVEY appendix B.5 ('metropolis' scores, [delta] proposal)."""
    pb = ref_shim.load()
    rng = np.random.default_rng(seed)
    x_obs = rng.normal(0, 1, size=N)
    y_obs = rng.normal(1.5 * x_obs - 1.0, 0.5)
    x = pb.RV('x', vtype=float, vset=[-3, 3])
    y = pb.RV('y', vtype=float, vset=[-np.inf, np.inf])
    beta_0 = pb.RV('beta_0', vtype=float, vset=[-6., 6.], pscale='log')
    beta_1 = pb.RV('beta_1', vtype=float, vset=[-6., 6.], pscale='log')
    y_sigma = pb.RV('y_sigma', vtype=float, vset=[(0.001), 10.], pscale='log')

    def norm_reg(x, y, beta_0, beta_1, y_sigma):
        return scipy.stats.norm.logpdf(y, loc=beta_0 + beta_1 * x, scale=y_sigma)

    stats = x & y
    paras = beta_0 & beta_1 & y_sigma
    sp = pb.SP(stats, paras)
    sp.set_prob(norm_reg, pscale='log')
    paras.set_tran(lambda **k: 0.)
    paras.set_delta([dstep])
    sp.set_tran(paras)
    sp.set_delta(paras)
    sp.set_scores('metropolis')
    R = rng.random((T, 4))
    init = np.array([-0.9, 1.4, 0.6])
    with ref_shim.injected_uniform(R.ravel()):
        sampler = sp.sampler({'beta_0': init[0], 'beta_1': init[1],
                              'y_sigma': init[2]},
                             {'x,y': [x_obs, y_obs]}, stop=T, iid=True,
                             joint=True)
        samples = [s for s in sampler]
    out = _summary_arrays(sp, samples, ['beta_0', 'beta_1', 'y_sigma'])
    delta = -dstep + (2 * dstep) * R[:, :3]
    out.update(x_obs=x_obs, y_obs=y_obs, delta=delta, thresh=R[:, 3], runif=R,
               init=init, dmax=np.full(3, dstep),
               lims=np.array([[-6., 6.], [-6., 6.], [0.001, 10.]]),
               ex=np.array([[False, False], [False, False], [False, False]]),
               log_ufun=np.array([False, False, False]), coef=np.array(1.0))
    return out


# ---------------------------------------------------------------------------
def dgei(seed, N, M, S):
    """examples/dgei/dgei_norm1d_improved.py:20-46 with (N, M, S) arguments."""
    pb = ref_shim.load()
    rng = np.random.default_rng(seed)
    data = rng.normal(50., 10., size=N)
    mu = pb.RV('mu', vtype=float, vset=(40, 60))
    sigma = pb.RV('sigma', vtype=float, vset=(5, 20.))
    x = pb.RV('x', vtype=float, vset={-np.inf, np.inf})
    sigma.set_ufun((np.log, np.exp))
    paras = pb.RF(mu, sigma)
    stats = pb.RF(x)
    model = pb.SD(stats, paras)
    model.set_prob(scipy.stats.norm.logpdf,
                   order={'x': 0, 'mu': 'loc', 'sigma': 'scale'}, pscale='log')
    joint = model({x: data, 'mu': {M}, 'sigma': {S}}, iid=True, joint=True)
    posterior = joint.conditionalise('x')
    post_mu = posterior.marginal('mu')
    post_sigma = posterior.marginal('sigma')
    expt = posterior.expectation()
    q_mu = post_mu.quantile()
    q_sigma = post_sigma.quantile()
    q3_mu = post_mu.quantile([0.025, 0.5, 0.975])
    return dict(data=data, mu=np.ravel(joint['mu']), sigma=np.ravel(joint['sigma']),
                joint=np.asarray(joint.prob), posterior=np.asarray(posterior.prob),
                marg_mu=np.asarray(post_mu.prob), marg_sigma=np.asarray(post_sigma.prob),
                post_linear=np.asarray(posterior.rescaled().prob),
                expt_mu=np.array(float(expt['mu'])),
                expt_sigma=np.array(float(expt['sigma'])),
                med_mu=np.array(float(q_mu['mu'])),
                med_sigma=np.array(float(q_sigma['sigma'])),
                q3_mu=np.array([float(q['mu']) for q in q3_mu]),
                joint_name=np.array(joint.name), post_name=np.array(posterior.name),
                marg_mu_name=np.array(post_mu.name))


# ---------------------------------------------------------------------------
def pd_serialise(seed, N, M, S):
    """probayes/pd_utils.py:433-553: the dict form (serialise / deserialise) of a DGEI joint
    (two array keys + the set-valued iid key), its posterior and a marginal, and what the
    reference's own write_serialised stores -- driven through tests/fake_h5py.py (h5py is
    not installed; the reference imports whatever module is called h5py).  The stored
    structure is recorded as JSON."""
    import json
    pb = ref_shim.load()
    sys.path.insert(0, os.path.join(os.path.dirname(OUT)))
    import fake_h5py
    import probayes.pd_utils as rpu
    rng = np.random.default_rng(seed)
    data = rng.normal(50., 10., size=N)
    mu = pb.RV('mu', vtype=float, vset=(40, 60))
    sigma = pb.RV('sigma', vtype=float, vset=(5, 20.))
    x = pb.RV('x', vtype=float, vset={-np.inf, np.inf})
    sigma.set_ufun((np.log, np.exp))
    model = pb.SD(pb.RF(x), pb.RF(mu, sigma))
    model.set_prob(scipy.stats.norm.logpdf,
                   order={'x': 0, 'mu': 'loc', 'sigma': 'scale'}, pscale='log')
    joint = model({x: data, 'mu': {M}, 'sigma': {S}}, iid=True, joint=True)
    posterior = joint.conditionalise('x')
    post_mu = posterior.marginal('mu').rescaled()
    old = rpu.h5py
    rpu.h5py = fake_h5py
    try:
        out = {}
        for tag, pd in (("joint", joint), ("posterior", posterior), ("post_mu", post_mu)):
            ser = rpu.serialise(pd)
            (name, d), = ser.items()
            path = "ref_" + tag
            aux = {"aux data": {"obs": data}} if tag == "joint" else {}
            out[tag + "_name"] = np.array(name)
            out[tag + "_keys"] = np.array(json.dumps([k for k in d.keys()]))
            out[tag + "_dims"] = np.array(json.dumps({k: None if v is None else int(v)
                                                      for k, v in pd.dims.items()}))
            # (the reference's writer rewrites None dims to 'None' IN the PD's own dims dict,
            # pd_utils.py:480-483: write a deep copy)
            import copy
            rpu.write_serialised(path, copy.deepcopy(ser), aux)
            back, = rpu.read_dist(path)
            out[tag + "_pscale_is_log"] = np.array(bool(np.iscomplexobj(pd.pscale)))
            out[tag + "_prob"] = np.asarray(pd.prob)
            out[tag + "_file"] = np.array(json.dumps(
                fake_h5py.dump(path), default=lambda o: o.item() if hasattr(o, "item") else str(o)))
            out[tag + "_back_name"] = np.array(back.name)
            out[tag + "_back_prob"] = np.asarray(back.prob)
    finally:
        rpu.h5py = old
    out.update(data=data, mu=np.ravel(joint['mu']), sigma=np.ravel(joint['sigma']))
    return out


# ---------------------------------------------------------------------------
def omc_rs_norm1d(seed, N, T):
    """examples/omc/omc_rs_sp_norm1d.py:17-52 with the prior draws injected:
    random sampling of (mu, sigma) through SP.sampler without a proposal, the
    summary PD, and its post-processing (rescaled / sorted / quantile /
    expectation: probayes/pd.py:373-499)."""
    pb = ref_shim.load()
    rng = np.random.default_rng(seed)
    data = rng.normal(50., 10., size=N)
    mu = pb.RV('mu', vtype=float, vset=(40, 60))
    sigma = pb.RV('sigma', vtype=float, vset=(5, 20.))
    x = pb.RV('x', vtype=float, vset=[-np.inf, np.inf])
    sigma.set_ufun((np.log, np.exp))
    process = pb.SP(pb.RF(x), pb.RF(mu, sigma))
    process.set_prob(scipy.stats.norm.logpdf,
                     order={'x': 0, 'mu': 'loc', 'sigma': 'scale'}, pscale='log')
    R = rng.random((T, 2))
    with ref_shim.injected_uniform(np.ravel(R)):
        sampler = process.sampler({'mu': {0}, 'sigma': {0}, 'x': data},
                                  iid=True, joint=True, stop=T)
        samples = [s for s in sampler]
    summary = process(samples)
    inference = summary.rescaled()
    mu_sort = inference.sorted('mu')
    sigma_sort = inference.sorted('sigma')
    qs = [0.025, 0.25, 0.5, 0.75, 0.975]
    expt = inference.expectation()
    expt2 = inference.expectation(['mu', 'sigma'], exponent=2)
    return dict(data=data, runif=R, mu=np.asarray(summary['mu'], float),
                sigma=np.asarray(summary['sigma'], float),
                logp=np.asarray(summary.prob, float), lin=np.asarray(inference.prob, float),
                name=np.array(summary.name), first_name=np.array(samples[0].name),
                mu_sorted=np.asarray(mu_sort['mu'], float),
                mu_sorted_sigma=np.asarray(mu_sort['sigma'], float),
                mu_sorted_prob=np.asarray(mu_sort.prob, float),
                sigma_sorted=np.asarray(sigma_sort['sigma'], float),
                sigma_sorted_prob=np.asarray(sigma_sort.prob, float),
                qs=np.array(qs),
                q_mu=np.array([float(q['mu']) for q in mu_sort.quantile(qs)]),
                q_sigma=np.array([float(q['sigma']) for q in sigma_sort.quantile(qs)]),
                med_mu=np.array(float(mu_sort.quantile(0.5)['mu'])),
                expt=np.array([float(expt['mu']), float(expt['sigma'])]),
                expt2=np.array([float(expt2['mu']), float(expt2['sigma'])]))


# ---------------------------------------------------------------------------
def pd_algebra(seed, N, M, S):
    """The joint-product / division algebra around the DGEI model
    (probayes/pd.py:564-615, pd_utils.py:85-328, pscales.py:160-236): RV and RF prior
    PDs, prior * likelihood, joint / evidence, joint / p(mu, x)."""
    pb = ref_shim.load()
    rng = np.random.default_rng(seed)
    data = rng.normal(50., 10., size=N)
    mu = pb.RV('mu', vtype=float, vset=(40, 60))
    sigma = pb.RV('sigma', vtype=float, vset=(5, 20.))
    x = pb.RV('x', vtype=float, vset={-np.inf, np.inf})
    sigma.set_ufun((np.log, np.exp))
    paras = pb.RF(mu, sigma)
    model = pb.SD(pb.RF(x), paras)
    model.set_prob(scipy.stats.norm.logpdf,
                   order={'x': 0, 'mu': 'loc', 'sigma': 'scale'}, pscale='log')
    grid = {'mu': {M}, 'sigma': {S}}
    joint = model({x: data, **grid}, iid=True, joint=True)
    like = model({x: data, **grid}, iid=True, joint=False)
    prior = paras(dict(grid))
    pmu, psg = mu({M}), sigma({S})
    pp = pmu * psg
    j2 = prior * like
    ev = joint.marginal('x')
    post = joint / ev
    mm = joint.marginal(['mu', 'x'])
    pc = joint / mm
    return dict(data=data, mu=np.ravel(joint['mu']), sigma=np.ravel(joint['sigma']),
                pmu=np.ravel(pmu.prob), psg=np.ravel(psg.prob), prior=np.asarray(prior.prob),
                pmu_psg=np.asarray(pp.prob), like=np.asarray(like.prob),
                joint=np.asarray(joint.prob), prior_like=np.asarray(j2.prob),
                evidence=np.array(float(ev.prob)), post=np.asarray(post.prob),
                marg_mu_x=np.asarray(mm.prob), cond_sigma=np.asarray(pc.prob),
                names=np.array([pmu.name, prior.name, pp.name, like.name, j2.name, ev.name,
                                post.name, mm.name, pc.name]))


def omc_rejection_circle(seed, T, radius=1.):
    """examples/omc/omc_rejection_sp_circle.py:26-39: ordinary Monte Carlo with rejection
    sampling (set_prop + custom scores / thresh / update): per step two box-uniform draws
    (x, y), the proposal density norm2d there (q), the target indicator (p), one threshold
    uniform in [0, coef_max); kept iff p >= t.  All uniforms injected, in call order."""
    pb = ref_shim.load()
    rng = np.random.default_rng(seed)

    def inside(x, y):
        return np.array(x**2 + y**2 <= radius**2, dtype=float)

    def norm2d(x, y, loc=0., scale=radius):
        return scipy.stats.norm.pdf(x, loc=loc, scale=scale) * \
            scipy.stats.norm.pdf(y, loc=loc, scale=scale)

    xy_range = [-radius, radius]
    x = pb.RV("x", xy_range)
    y = pb.RV("y", xy_range)
    process = pb.SP(x & y)
    process.set_prob(inside)
    process.set_prop(norm2d)
    process.set_scores(lambda opqr: opqr.p.prob)
    coef_max = float(norm2d(radius, 1.))
    R = rng.random((T, 3))               # per step: x draw, y draw, threshold
    with ref_shim.injected_uniform(R.ravel()):
        process.set_thresh(np.random.uniform, low=0., high=coef_max)   # captures the patch
        process.set_update(lambda stu: stu.s >= stu.t)
        sampler = process.sampler({0}, stop=T)
        samples = [sample for sample in sampler]
    summary = process(samples)
    xy = np.array([(float(s.p['x']), float(s.p['y'])) for s in samples])
    kept = summary.p
    return dict(runif=R, xy=xy, p=np.array([float(s.p.prob) for s in samples]),
                q=np.array([float(s.q.prob) for s in samples]),
                s=np.array([float(s.s) for s in samples]),
                t=np.array([float(s.t) for s in samples]),
                u=np.array([bool(s.u) for s in samples]),
                kept_x=np.asarray(kept['x'], float), kept_y=np.asarray(kept['y'], float),
                kept_prob=np.asarray(kept.prob, float), kept_size=np.array(kept.size),
                coef_max=np.array(coef_max), radius=np.array(radius),
                names=np.array([samples[0].p.name, samples[0].q.name, kept.name,
                                summary.q.name]),
                q_keys=np.array(list(samples[0].q.keys())),
                first_v_is_p=np.array(samples[0].v is samples[0].p),
                v_none=np.array([s.v is None for s in samples]))


def pd_cond_array(seed, N, M, S):
    """PD.conditionalise on ARRAY-valued keys (probayes/pd.py:214-295): p(mu, sigma, x)
    -> p(mu, x | sigma), p(sigma, x | mu) (axis move), p(mu | sigma, x) (normalise first),
    in log and linear pscale, with the marginal taken from a conditional."""
    pb = ref_shim.load()
    rng = np.random.default_rng(seed)
    data = rng.normal(50., 10., size=N)
    mu = pb.RV('mu', vtype=float, vset=(40, 60))
    sigma = pb.RV('sigma', vtype=float, vset=(5, 20.))
    x = pb.RV('x', vtype=float, vset={-np.inf, np.inf})
    sigma.set_ufun((np.log, np.exp))
    model = pb.SD(pb.RF(x), pb.RF(mu, sigma))
    model.set_prob(scipy.stats.norm.logpdf,
                   order={'x': 0, 'mu': 'loc', 'sigma': 'scale'}, pscale='log')
    joint = model({x: data, 'mu': {M}, 'sigma': {S}}, iid=True, joint=True)
    c_sig = joint.conditionalise('sigma')
    c_mu = joint.conditionalise('mu')
    c_sx = joint.conditionalise(['sigma', 'x'])
    lin = joint.conditionalise('x').rescaled()
    l_sig = lin.conditionalise('sigma')
    l_mu = lin.conditionalise('mu')
    dims = lambda pd: np.array([-1 if pd.dims[k] is None else pd.dims[k] for k in ('mu', 'sigma')])
    return dict(data=data, mu=np.ravel(joint['mu']), sigma=np.ravel(joint['sigma']),
                joint=np.asarray(joint.prob),
                c_sig=np.asarray(c_sig.prob), c_mu=np.asarray(c_mu.prob),
                c_sx=np.asarray(c_sx.prob), lin=np.asarray(lin.prob),
                l_sig=np.asarray(l_sig.prob), l_mu=np.asarray(l_mu.prob),
                dims=np.stack([dims(c_sig), dims(c_mu), dims(c_sx), dims(l_sig), dims(l_mu)]),
                shapes=np.array([np.shape(c_mu['mu']), np.shape(c_mu['sigma'])]),
                names=np.array([c_sig.name, c_mu.name, c_sx.name, l_sig.name, l_mu.name]))


# ---------------------------------------------------------------------------
def gibbs2d(seed, T):
    """examples/mcmc/gibbs_norm2d.py:15-22 with the cdf uniforms injected."""
    pb = ref_shim.load()
    rng = np.random.default_rng(seed)
    lims = (-10., 10.)
    means = [0.5, -0.5]
    covar = [[1.5, -1.0], [-1.0, 2.]]
    x = pb.RV('x', vtype=float, vset=lims)
    y = pb.RV('y', vtype=float, vset=lims)
    sp = pb.SP(x & y)
    sp.set_prob(scipy.stats.multivariate_normal, means, covar)
    sp.set_tran(scipy.stats.multivariate_normal, means, covar, tsteps=1)
    sp.set_scores('gibbs')
    R = rng.random(T)
    with ref_shim.injected_uniform(R):
        sampler = sp.sampler({'x': 0., 'y': 1.}, stop=T)
        samples = [s for s in sampler]
    summary = sp(samples)
    xs = np.stack([np.asarray(summary.v['x'], float),
                   np.asarray(summary.v['y'], float)], axis=-1)
    cc = sp._cond_cov
    return dict(x=xs, prob=np.asarray(summary.v.prob, float), runif=R,
                init=np.array([0., 1.]), mean=np.array(means), cov=np.array(covar),
                lims=np.array([lims, lims]), stdv=np.asarray(cc._stdv),
                cdfs=np.asarray(cc._cdfs),
                coef=np.stack([np.ravel(c) for c in cc._coef]),
                n_true=np.array(summary.u.count(True)))


def condcov_bare(seed, d, T):
    """Bare probayes/cond_cov.py CondCov at dimension d: construction constants
    and T cyclic coordinate draws from injected uniforms."""
    ref_shim.load()
    from probayes.cond_cov import CondCov
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + np.eye(d)
    mean = rng.standard_normal(d)
    lims = np.tile(np.array([-10., 10.]), (d, 1))
    cc = CondCov(mean, cov, lims)
    R = rng.random(T)
    x = mean.copy()
    X = np.empty((T, d))
    with ref_shim.injected_uniform(R):
        for k in range(T):
            i = k % d
            args = [float(v) for v in x]
            args[i] = {0}
            x[i] = float(cc.interp(*args))
            X[k] = x
    coef = np.zeros((d, d))
    for i in range(d):
        idx = [j for j in range(d) if j != i]
        coef[i, idx] = np.ravel(cc._coef[i])
    return dict(x=X, runif=R, init=mean.copy(), mean=mean, cov=cov, lims=lims,
                stdv=np.asarray(cc._stdv), cdfs=np.asarray(cc._cdfs), coef=coef)


# ---------------------------------------------------------------------------
def pscales_table():
    """Edge-case table for the clamped log/exp/div of probayes/pscales.py."""
    ref_shim.load()
    from probayes import pscales as ps
    p = np.array([0.0, 1e-320, 2.2250738585072014e-308, 2.3e-308, 1e-300, 0.5,
                  1.0, 7.0, 1e300])
    l = np.array([-1.7976931348623158e308, -1e4, -745.2, -744.0, -708.4, -1.0,
                  0.0, 709.7, 709.79, 1e4])
    num = np.array([0.0, 1e-310, 1e-300, 0.3, 2.0, 1e308])
    den = np.array([0.0, 1e-310, 1e-300, 0.6, 1.0, 1e-308])
    lnum = np.array([-800., -750., -700., -1.0, 0.0, 3.0])
    lden = np.array([-800., -745., -710., -2.0, 0.0, -3.0])
    return dict(p=p, log_prob=ps.log_prob(p), l=l, exp_logp=ps.exp_logp(l),
                num=num, den=den, div_lin=ps.div_prob(num, den),
                lnum=lnum, lden=lden,
                div_log_to_lin=ps.div_prob(lnum, lden, 0j, 0j, pscale=1.),
                resc_lin_to_log=ps.rescale(p, 1., 0j),
                resc_log_to_lin=ps.rescale(l, 0j, 1.))


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = {
        "mh_mvn_c1": lambda: mh_mvn(11, 512, (0., 1.)),
        "mh_mvn_c1_b": lambda: mh_mvn(12, 256, (-2.5, 3.0)),
        "mh_mvn_log": lambda: mh_mvn(13, 256, (0., 1.), log_pscale=True),
        "mh_mvn_3d": lambda: mh_mvn3d(15, 256),
        "mh_mvn_bound": lambda: mh_mvn_bound(16, 400),
        "gibbs3d": lambda: gibbs3d(54, 300),
        "gibbs3d_sweep": lambda: gibbs3d(55, 120, tsteps=3),
        "gibbs3d_all": lambda: gibbs3d(56, 120, tsteps=None),
        "mh_norm1d_hastings": lambda: mh_norm1d(21, 300, 60, 'hastings'),
        "mh_norm1d_metropolis": lambda: mh_norm1d(22, 300, 60, 'metropolis'),
        "mh_norm1d_underflow": lambda: mh_norm1d(23, 60, 1000, 'metropolis'),
        "mh_norm1d_spherical": lambda: mh_norm1d(24, 300, 60, 'hastings', spherical=True),
        "mh_norm1d_bound_open": lambda: mh_norm1d(25, 300, 60, 'metropolis', bound='open'),
        "mh_norm1d_bound_mixed": lambda: mh_norm1d(26, 300, 60, 'metropolis', bound='mixed'),
        "mh_linreg": lambda: mh_linreg(31, 300, 100),
        "dgei_small": lambda: dgei(41, 60, 48, 40),
        "dgei_peaked": lambda: dgei(42, 2000, 40, 36),
        "gibbs2d": lambda: gibbs2d(51, 400),
        "pd_algebra": lambda: pd_algebra(71, 40, 9, 7),
        "pd_cond_array": lambda: pd_cond_array(72, 40, 9, 7),
        "omc_rs_norm1d": lambda: omc_rs_norm1d(61, 60, 400),
        "omc_rejection_circle": lambda: omc_rejection_circle(62, 500),
        "condcov_d8": lambda: condcov_bare(52, 8, 160),
        "condcov_d64": lambda: condcov_bare(53, 64, 256),
        "pscales": pscales_table,
        "pd_serialise": lambda: pd_serialise(73, 30, 7, 5),
    }
    only = sys.argv[1:]
    for name, fn in cases.items():
        if only and name not in only:
            continue
        out = fn()
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
