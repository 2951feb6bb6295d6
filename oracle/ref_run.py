"""Ships and times the REAL reference (pure Python) -- TEST INFRASTRUCTURE, see
oracle/__init__.py.  Only bench.py's ``cpu_baseline`` / ``--impl reference`` legs and
``__graft_entry__.build()`` use this module.

``make_ref()``   development container only: copies the reference's ``probayes``
                 package from /root/reference into the git-ignored ``oracle/_ref/``
                 (next to a stub ``h5py``) so that it travels to
                 the GPU box with the repo snapshot like a built ``.so`` does.  Nothing
                 under ``oracle/_ref/`` is ever committed.
``c1_rate()``    config C1 of BASELINE.json -- examples/mcmc/mcmc_prob4a.py:32-51
                 unchanged except for the pylab import: one chain, the 2-D correlated
                 normal target, ``[sample for sample in sampler]`` (probayes/sp.py:
                 221-258 per step) and ``process(samples)``; returns chain-steps/s.
``c1_rate_all_cores()``  ``n`` independent single-chain samplers in ``n`` processes
                 (the reference has no parallelism of its own), summed steps/s.
``c3_rate()``    config C3's model at the largest size the reference evaluates sensibly
                 through its API (N = 10^5): the MH sampler of
                 examples/mcmc/gibbs_linreg.py's model, one log-likelihood evaluation of
                 N terms per step; returns evals/s.
``c4_rate()``    config C4's model (examples/dgei/dgei_norm1d_improved.py:20-46) at
                 N x M x S = 1000 x 256 x 256 (the [N, M, S] temporary is what limits the
                 reference): joint + conditionalise + both marginals; returns cell
                 evaluations/s and terms/s.
``c5_rate()``    config C5's per-step kernel, ``CondCov.interp`` at d = 64, one chain;
                 returns coordinate updates/s.
"""
import os
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SRC_ROOT = os.environ.get("PROBAYES_REFERENCE", "/root/reference")


def make_ref(force=False):
    """Copies /root/reference/probayes -> oracle/_ref/probayes (+ h5py stub).  Returns
    the directory, or None when the reference checkout is absent (GPU box)."""
    src = os.path.join(SRC_ROOT, "probayes")
    dst = os.path.join(REF_DIR, "probayes")
    if not os.path.isdir(src):
        return REF_DIR if os.path.isdir(dst) else None
    if force or not os.path.isdir(dst):
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        os.makedirs(REF_DIR, exist_ok=True)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    try:
        import h5py  # noqa: F401
    except ImportError:
        os.makedirs(os.path.join(REF_DIR, "h5py"), exist_ok=True)
        with open(os.path.join(REF_DIR, "h5py", "__init__.py"), "w") as f:
            f.write("# stub: the image has no h5py; probayes/pd_utils.py:8 imports it at module\n"
                    "# top and uses it only in the HDF5 serialisers (off the hot path)\n")
    return REF_DIR


def available():
    return os.path.isdir(os.path.join(REF_DIR, "probayes"))


def load():
    """The reference ``probayes`` module imported from oracle/_ref."""
    if not available():
        raise RuntimeError("oracle/_ref/probayes is absent (run oracle.ref_run.make_ref() in "
                           "the development container)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import probayes as pb
    if not os.path.abspath(pb.__file__).startswith(REF_DIR):
        raise RuntimeError("a different 'probayes' is already imported: " + pb.__file__)
    import probayes.prob as _pp
    if "fit" in _pp.SCIPY_DIST_METHODS:          # scipy >= 1.1x frozen mvn has no .fit
        _pp.SCIPY_DIST_METHODS.remove("fit")
    return pb


def c1_rate(n_steps=12288, seed=0):
    """Returns (chain_steps_per_s, seconds, n_accept) of the reference's own sampler
    on config C1 (one core)."""
    import numpy as np
    import scipy.stats
    pb = load()
    np.random.seed(seed)
    prop_stdv = np.sqrt(1)

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
    process.set_delta(lambda: process.Delta(x=scipy.stats.norm.rvs(loc=0., scale=prop_stdv),
                                            y=scipy.stats.norm.rvs(loc=0., scale=prop_stdv)))
    process.set_scores('hastings')
    process.set_update('metropolis')
    t0 = time.perf_counter()
    sampler = process.sampler({'x': 0., 'y': 1.}, stop=n_steps)
    samples = [sample for sample in sampler]
    summary = process(samples)
    n_accept = summary.u.count(True)
    dt = time.perf_counter() - t0
    return n_steps / dt, dt, n_accept


def _worker(args):
    n_steps, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    return c1_rate(n_steps, seed)


def c1_rate_all_cores(n_steps=4096, procs=None):
    """``procs`` independent single-chain reference samplers, one process each.
    Returns (summed chain_steps_per_s, procs, wall seconds)."""
    import multiprocessing as mp
    procs = procs or len(os.sched_getaffinity(0))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        pool.map(_worker, [(64, 1000 + i) for i in range(procs)])       # import + warm-up
        t0 = time.perf_counter()
        pool.map(_worker, [(n_steps, i) for i in range(procs)])
        wall = time.perf_counter() - t0
    return procs * n_steps / wall, procs, wall


def c3_rate(n_obs=100000, n_steps=40, seed=2024):
    """(log-likelihood evals/s, seconds): one eval = N terms (one MH step of one chain)."""
    import numpy as np
    import scipy.stats
    pb = load()
    rng = np.random.default_rng(seed)
    x_obs = rng.normal(0, 1, size=n_obs)
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
    paras.set_delta([0.002])
    sp.set_tran(paras)
    sp.set_delta(paras)
    sp.set_scores('metropolis')
    np.random.seed(seed)
    t0 = time.perf_counter()
    sampler = sp.sampler({'beta_0': -1., 'beta_1': 1.5, 'y_sigma': 0.5},
                         {'x,y': [x_obs, y_obs]}, stop=n_steps, iid=True, joint=True)
    samples = [s for s in sampler]
    dt = time.perf_counter() - t0
    return len(samples) / dt, dt


def c4_rate(n_obs=1000, m=256, s=256, seed=7):
    """(cell evaluations/s, terms/s, seconds) of joint + conditionalise + both marginals."""
    import numpy as np
    import scipy.stats
    pb = load()
    rng = np.random.default_rng(seed)
    data = rng.normal(50., 10., size=n_obs)
    mu = pb.RV('mu', vtype=float, vset=(40, 60))
    sigma = pb.RV('sigma', vtype=float, vset=(5, 20.))
    x = pb.RV('x', vtype=float, vset={-np.inf, np.inf})
    sigma.set_ufun((np.log, np.exp))
    model = pb.SD(pb.RF(x), pb.RF(mu, sigma))
    model.set_prob(scipy.stats.norm.logpdf,
                   order={'x': 0, 'mu': 'loc', 'sigma': 'scale'}, pscale='log')
    t0 = time.perf_counter()
    joint = model({x: data, 'mu': {m}, 'sigma': {s}}, iid=True, joint=True)
    posterior = joint.conditionalise('x')
    posterior.marginal('mu')
    posterior.marginal('sigma')
    dt = time.perf_counter() - t0
    return m * s / dt, n_obs * m * s / dt, dt


def c5_rate(d=64, n_steps=4096, seed=0):
    """(coordinate updates/s, seconds) of the reference's conditional-normal Gibbs update at
    dimension d, one chain: ``CondCov.interp`` (probayes/cond_cov.py:42-65) called as
    ``RF.eval_tfun`` calls it, cycling the coordinate.  (Through ``SP`` the reference gives
    every RV its own broadcast axis and numpy stops at 32 dimensions, so d = 64 cannot run
    through its sampler at all; this is its per-step kernel without the SP overhead, i.e. an
    upper bound on what its sampler could do.)"""
    import numpy as np
    load()
    from probayes.cond_cov import CondCov
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + np.eye(d)
    mean = rng.standard_normal(d)
    cc = CondCov(mean, cov, np.tile(np.array([-10., 10.]), (d, 1)))
    np.random.seed(seed)
    x = mean.copy()
    t0 = time.perf_counter()
    for k in range(n_steps):
        i = k % d
        args = [float(v) for v in x]
        args[i] = {0}
        x[i] = float(cc.interp(*args))
    dt = time.perf_counter() - t0
    return n_steps / dt, dt


if __name__ == "__main__":
    print(make_ref(force="--force" in sys.argv))
    if "--time" in sys.argv:
        print(c1_rate(2048))
        print(c1_rate_all_cores(1024))
        print("c3", c3_rate())
        print("c4", c4_rate())
        print("c5", c5_rate())
