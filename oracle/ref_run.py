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


if __name__ == "__main__":
    print(make_ref(force="--force" in sys.argv))
    if "--time" in sys.argv:
        print(c1_rate(2048))
        print(c1_rate_all_cores(1024))
