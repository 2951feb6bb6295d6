"""K5 parity (GPU, through the C ABI): CondCov Gibbs updates and the batched mvn
density against the golden fixtures (live reference) and the numpy oracle."""
import numpy as np
import pytest
from conftest import load_golden, relerr
from gpu_util import engine, dev, tcd_to_tdc, host
from oracle import np_oracle as o
from oracle import philox

pytestmark = pytest.mark.gpu
TOL = 1e-12
# Gibbs trajectories: device normcdfinv against scipy's ndtri (Cephes); both are accurate to
# ~1 ulp, and every later coordinate inherits the difference through the conditional means
GTOL = 1e-12


def test_condcov_constants_match_reference():
    from probayes_b200.cond_cov import CondCov
    for name in ("condcov_d8", "condcov_d64", "gibbs2d"):
        g = load_golden(name)
        cc = CondCov(g["mean"], g["cov"], g["lims"])
        assert relerr(cc.stdv, g["stdv"]) <= TOL
        assert np.abs(cc.cdfs - g["cdfs"]).max() <= TOL
        coef = g["coef"] if name != "gibbs2d" else \
            np.array([[0, g["coef"][0, 0]], [g["coef"][1, 0], 0]])
        assert np.abs(cc.coef_matrix() - coef).max() <= TOL


@pytest.mark.parametrize("name", ["gibbs2d", "gibbs3d"])
def test_gibbs2d_golden(name):
    """examples/mcmc/gibbs_norm2d.py (and a 3-D variant) through the reference with
    injected uniforms: trajectory and the recorded (permuted-order) mvn pdf."""
    from probayes_b200.cond_cov import CondCov
    eng = engine()
    g = load_golden(name)
    cc = CondCov(g["mean"], g["cov"], g["lims"])
    T = len(g["runif"])
    state = dev(eng, g["init"][:, None])
    out = eng.gibbs_mvn(state, cc, T, inj_runif=dev(eng, g["runif"][:, None]))
    eng.sync()
    assert np.abs(host(out["x"])[:, :, 0] - g["x"]).max() <= GTOL
    assert relerr(host(out["prob"])[:, 0], g["prob"]) <= GTOL


@pytest.mark.parametrize("name", ["condcov_d8", "condcov_d64"])
def test_condcov_golden_trajectory(name):
    from probayes_b200.cond_cov import CondCov
    eng = engine()
    g = load_golden(name)
    cc = CondCov(g["mean"], g["cov"], g["lims"])
    T = len(g["runif"])
    state = dev(eng, g["init"][:, None])
    out = eng.gibbs_mvn(state, cc, T, inj_runif=dev(eng, g["runif"][:, None]),
                        want_prob=False)
    eng.sync()
    assert np.abs(host(out["x"])[:, :, 0] - g["x"]).max() <= GTOL


@pytest.mark.parametrize("d,C,T,thin", [(2, 100, 41, 1), (3, 33, 50, 3), (8, 257, 64, 8),
                                        (20, 64, 70, 5), (64, 130, 150, 64),
                                        (128, 40, 256, 128), (100, 33, 210, 7),
                                        # whole sweeps -> the tensor-core kernel, incl. d off
                                        # the 8-coordinate block (zero padding) and ragged C
                                        (5, 9, 10, 5), (13, 50, 39, 13), (33, 20, 66, 33),
                                        (64, 70, 128, 64), (100, 11, 200, 100)])
def test_oracle_injected_and_resume(d, C, T, thin):
    from probayes_b200.cond_cov import CondCov
    eng = engine()
    rng = np.random.default_rng(d * 7 + C)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + np.eye(d)
    mean = rng.standard_normal(d)
    lims = np.tile([-10., 10.], (d, 1))
    cc = CondCov(mean, cov, lims)
    init = rng.standard_normal((C, d))
    R = rng.random((T, C))
    ref = o.gibbs_mvn_walk(init, R, mean, cov, lims, log_pscale=True)
    state = dev(eng, init.T)
    out = eng.gibbs_mvn(state, cc, T, thin=thin, log_pscale=True, inj_runif=dev(eng, R),
                        stats=True)
    eng.sync()
    sel = slice(thin - 1, None, thin)
    X = tcd_to_tdc(ref["x"])[sel]
    assert np.abs(host(out["x"]) - X).max() <= GTOL * max(1.0, np.abs(X).max())
    assert relerr(host(out["prob"]), ref["prob"][sel]) <= GTOL
    assert np.abs(host(state) - ref["x"][-1].T).max() <= GTOL * max(1.0, np.abs(X).max())
    assert np.abs(host(out["stat_sum"]) - X.sum(axis=0)).max() <= 1e-11 * max(1.0, np.abs(X).max())
    # resume mid-sweep: two calls == one call
    h = T // 2
    st2 = dev(eng, init.T)
    eng.gibbs_mvn(st2, cc, h, inj_runif=dev(eng, R[:h]), record=False)
    eng.gibbs_mvn(st2, cc, T - h, step0=h, inj_runif=dev(eng, R[h:]), record=False)
    eng.sync()
    assert np.abs(host(st2) - host(state)).max() <= 1e-13


def test_philox_replay():
    from probayes_b200.cond_cov import CondCov
    eng = engine()
    d, C, T, seed = 8, 96, 80, 2468
    rng = np.random.default_rng(1)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + np.eye(d)
    mean = rng.standard_normal(d)
    lims = np.tile([-10., 10.], (d, 1))
    cc = CondCov(mean, cov, lims)
    init = np.tile(mean, (C, 1))
    t = np.arange(T, dtype=np.uint64)[:, None]
    c = np.arange(C, dtype=np.uint64)[None, :]
    R = philox.gibbs_uniforms(seed, t, c)
    ref = o.gibbs_mvn_walk(init, R, mean, cov, lims)
    state = dev(eng, init.T)
    out = eng.gibbs_mvn(state, cc, T, seed=seed)
    eng.sync()
    assert np.abs(host(out["x"]) - tcd_to_tdc(ref["x"])).max() <= GTOL


@pytest.mark.parametrize("d,n", [(2, 1000), (5, 777), (64, 4096), (64, 1003), (64, 7),
                                 (128, 515), (100, 64)])
def test_mvn_logpdf_batched(d, n):
    """d = 64 goes through the FP64 tensor-core (DMMA) kernel."""
    eng = engine()
    rng = np.random.default_rng(d + n)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + np.eye(d)
    mean = rng.standard_normal(d)
    x = rng.standard_normal((n, d)) * 1.5 + mean
    U, lpd = o.mvn_whiten(cov)
    for reorder in (True, False):
        order = o.mvn_value_order(d) if reorder else list(range(d))
        want = o.mvn_logpdf(x[:, order], mean, U, lpd)
        got = host(eng.mvn_logpdf(dev(eng, x.T), mean, cov, log_pscale=True, reorder=reorder))
        assert relerr(got, want) <= TOL
    got = host(eng.mvn_logpdf(dev(eng, x.T), mean, cov, log_pscale=False))
    want = np.exp(o.mvn_logpdf(x[:, o.mvn_value_order(d)], mean, U, lpd))
    assert relerr(got, want) <= 1e-11


def test_gibbs_d64_moments():
    """Config C5 shape at reduced size: d = 64, 4096 chains, native RNG; pooled
    mean / covariance of the draws match N(mean, cov) within Monte Carlo error."""
    from probayes_b200.cond_cov import CondCov
    eng = engine()
    d, C = 64, 4096
    rng = np.random.default_rng(0)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + np.eye(d)
    mean = rng.standard_normal(d)
    cc = CondCov(mean, cov, np.tile([-10., 10.], (d, 1)))
    state = dev(eng, np.tile(mean[:, None], (1, C)))
    eng.gibbs_mvn(state, cc, 40 * d, seed=11, record=False)              # burn-in sweeps
    out = eng.gibbs_mvn(state, cc, 20 * d, seed=11, step0=40 * d, thin=d, want_prob=False)
    eng.sync()
    X = host(out["x"]).transpose(1, 0, 2).reshape(d, -1)
    assert np.abs(X.mean(axis=1) - mean).max() < 0.03
    assert np.abs(np.cov(X) - cov).max() < 0.06


def test_table_ndtri_device():
    """pbx_ndtri on the device: the table path against scipy's ndtri and, bit for bit, against
    the library's host mirror; the out-of-table inputs through normcdfinv (limits kept)."""
    import ctypes as C
    import torch
    from scipy.special import ndtri
    from probayes_b200 import _lib
    eng = engine()
    rng = np.random.default_rng(3)
    u = np.concatenate([rng.random(1 << 20), 2.0 ** -rng.uniform(1, 63, 1 << 18),
                        1 - 2.0 ** -rng.uniform(1, 52, 1 << 16),
                        [0.5, 2.0 ** -53, 1 - 2.0 ** -53]])
    got = host(eng.ndtri(dev(eng, u)))
    ref = ndtri(u)
    excess = np.abs(got - ref) - (3e-15 * np.abs(ref) + 2e-16)
    assert excess.max() <= 0, excess.max()
    hm = np.empty_like(u)
    assert _lib.load().pbx_ndtri_host(u.ctypes.data_as(C.c_void_p), C.c_int64(len(u)),
                                      hm.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(got, hm)
    edge = np.array([0.0, 1.0, 2.0 ** -70, 1e-300])
    ge = host(eng.ndtri(dev(eng, edge)))
    assert ge[0] == -np.inf and ge[1] == np.inf
    assert abs(ge[2] / ndtri(edge[2]) - 1) <= 1e-14 and abs(ge[3] / ndtri(edge[3]) - 1) <= 1e-14
