"""K2 parity (GPU, through the C ABI): streaming normal log-likelihood MH against
the golden fixtures (live reference) and the numpy oracle.  Accept decisions
identical; parameters / log-joints within 1e-12 relative (north_star, fp64)."""
import numpy as np
import pytest
from conftest import load_golden, relerr
from gpu_util import engine, dev, tcd_to_tdc, host
from oracle import np_oracle as o
from oracle import philox

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _run_golden(eng, g, has_slope, accept, variant):
    T = len(g["thresh"])
    y = dev(eng, g["y_obs"] if has_slope else g["x_obs"])
    x = dev(eng, g["x_obs"]) if has_slope else None
    state = dev(eng, g["init"][:, None])
    out = eng.mh_normreg(state, y, x, T, g["lims"], g["ex"], g["log_ufun"], g["dmax"],
                         accept=accept, accept_coef=float(g["coef"]), variant=variant,
                         prop_bound=bool(g["bound"]) if "bound" in g.files else False,
                         inj_delta=dev(eng, tcd_to_tdc(g["delta"][:, None, :])),
                         inj_thresh=dev(eng, g["thresh"][:, None]), per_step=True)
    eng.sync()
    return out


@pytest.mark.parametrize("variant", [1, 2, 4])
@pytest.mark.parametrize("name,has_slope", [
    ("mh_norm1d_hastings", False), ("mh_norm1d_metropolis", False),
    ("mh_norm1d_underflow", False), ("mh_linreg", True),
    ("mh_norm1d_bound_open", False), ("mh_norm1d_bound_mixed", False)])
def test_golden_injected_reference_rule(name, has_slope, variant):
    eng = engine()
    g = load_golden(name)
    out = _run_golden(eng, g, has_slope, "reference", variant)
    assert np.array_equal(host(out["accept"])[:, 0].astype(bool), g["u"])
    assert relerr(host(out["x"])[:, :, 0], g["x"]) <= TOL
    assert relerr(host(out["prob"])[:, 0], g["prob"]) <= TOL
    s = host(out["score"])[:, 0]
    assert np.isnan(s[0])
    assert np.nanmax(np.abs(s[1:] - g["s"][1:])) <= 1e-9   # s = exp(~ -400 .. 0): abs tol


def test_underflow_golden_log_rule_accepts():
    """At N=1000 the reference degenerates (only step 1 accepted); the log-space
    rule reproduces the oracle's log-rule decisions instead."""
    eng = engine()
    g = load_golden("mh_norm1d_underflow")
    out = _run_golden(eng, g, False, "log", 0)
    ref = o.mh_normreg_walk(g["init"][None], g["delta"][:, None, :], g["thresh"][:, None],
                            None, g["x_obs"], g["lims"], g["ex"], g["log_ufun"],
                            has_slope=False, accept="log")
    assert np.array_equal(host(out["accept"]).astype(bool), ref["u"])
    assert ref["u"].sum() > 10
    assert relerr(host(out["x"]), tcd_to_tdc(ref["x"])) <= TOL


@pytest.mark.parametrize("P,C,N,T,variant", [
    (3, 300, 5000, 40, 1), (3, 7, 4097, 60, 2), (2, 1100, 2048, 30, 1), (2, 5, 12345, 50, 2),
    (3, 2200, 777, 25, 1), (3, 4300, 4096 * 3 + 1, 12, 1), (2, 1, 3, 20, 2), (3, 9, 1, 20, 1),
    # variant 4: the whole walk in one launch, a warp per chain, observations in shared memory
    (3, 37, 1000, 60, 4), (2, 5, 8192, 30, 4), (3, 130, 8192, 12, 4), (2, 3, 1, 20, 4),
    (3, 1, 33, 25, 4), (2, 257, 64, 40, 4)])
def test_oracle_injected(P, C, N, T, variant):
    """Ragged / tiny / multi-tile observation counts, both kernels, KC = 1, 2, 4."""
    eng = engine()
    rng = np.random.default_rng(1000 * P + C + N)
    has_slope = P == 3
    x_obs = rng.normal(0, 1, N)
    y_obs = rng.normal(1.5 * x_obs - 1.0, 0.5) if has_slope else rng.normal(50., 10., N)
    if has_slope:
        lims = np.array([[-6., 6.], [-6., 6.], [0.001, 10.]])
        ex = np.array([[0, 0], [0, 0], [1, 0]])
        lg = np.array([0, 0, 0])
        init = np.stack([rng.normal(-1, .05, C), rng.normal(1.5, .05, C),
                         rng.uniform(.4, .7, C)], axis=1)
        dmax = np.array([0.02, 0.02, 0.02])
    else:
        lims = np.array([[40., 60.], [5., 20.]])
        ex = np.array([[1, 1], [1, 1]])
        lg = np.array([0, 1])
        init = np.stack([rng.uniform(45, 55, C), rng.uniform(8, 12, C)], axis=1)
        dmax = np.array([0.3, 0.05])
    delta = (rng.random((T, C, P)) * 2 - 1) * dmax
    thresh = rng.random((T, C))
    ref = o.mh_normreg_walk(init, delta, thresh, x_obs if has_slope else None, y_obs, lims,
                            ex.astype(bool), lg.astype(bool), has_slope=has_slope,
                            accept="log")
    state = dev(eng, init.T)
    out = eng.mh_normreg(state, dev(eng, y_obs), dev(eng, x_obs) if has_slope else None, T,
                         lims, ex, lg, dmax, accept="log", variant=variant,
                         inj_delta=dev(eng, tcd_to_tdc(delta)), inj_thresh=dev(eng, thresh),
                         per_step=True)
    eng.sync()
    assert np.array_equal(host(out["accept"]).astype(bool), ref["u"])
    assert relerr(host(out["x"]), tcd_to_tdc(ref["x"])) <= TOL
    assert relerr(host(out["prob"]), ref["prob"]) <= TOL
    assert np.array_equal(host(out["accept_count"]), ref["u"].sum(axis=0))
    assert relerr(host(state), ref["x"][-1].T) <= TOL


def test_out_of_box_proposals_are_rejected():
    """Proposals leaving the prior box get NEARLY_NEGATIVE_INF and are rejected
    (rv_utils.py:30-38); the walk stays inside."""
    eng = engine()
    rng = np.random.default_rng(5)
    N, C, T = 500, 64, 80
    y_obs = rng.normal(50., 10., N)
    lims = np.array([[49.5, 50.5], [9., 11.]])
    ex = np.array([[1, 1], [0, 0]])
    lg = np.array([0, 1])
    init = np.tile([50., 10.], (C, 1))
    dmax = np.array([1.0, 0.2])
    delta = (rng.random((T, C, 2)) * 2 - 1) * dmax
    thresh = rng.random((T, C))
    ref = o.mh_normreg_walk(init, delta, thresh, None, y_obs, lims, ex.astype(bool),
                            lg.astype(bool), has_slope=False, accept="log")
    out = eng.mh_normreg(dev(eng, init.T), dev(eng, y_obs), None, T, lims, ex, lg, dmax,
                         inj_delta=dev(eng, tcd_to_tdc(delta)), inj_thresh=dev(eng, thresh),
                         per_step=True)
    eng.sync()
    X = host(out["x"])
    assert np.array_equal(host(out["accept"]).astype(bool), ref["u"])
    assert relerr(X, tcd_to_tdc(ref["x"])) <= TOL
    # a chain that is inside the box never moves out of it (step 1 accepts
    # unconditionally, so start from the chains that stayed inside on step 1)
    inside = (X[:, 0] > 49.5) & (X[:, 0] < 50.5) & (X[:, 1] >= 9.) & (X[:, 1] <= 11.)
    ok0 = inside[0]
    assert ok0.sum() > 10 and inside[:, ok0].all()
    assert (~ref["u"]).sum() > 100


@pytest.mark.parametrize("variant,C", [(1, 200), (2, 6), (4, 70)])
def test_philox_replay_and_resume(variant, C):
    eng = engine()
    rng = np.random.default_rng(8)
    N, T, seed = 3000, 60, 991
    x_obs = rng.normal(0, 1, N)
    y_obs = rng.normal(1.5 * x_obs - 1.0, 0.5)
    lims = np.array([[-6., 6.], [-6., 6.], [0.001, 10.]])
    ex = np.zeros((3, 2), int); lg = np.zeros(3, int)
    dmax = np.array([0.02, 0.03, 0.01])
    init = np.tile([-1., 1.5, 0.5], (C, 1))
    R = philox.uniforms(seed, T, C, 3)
    delta = -dmax + 2 * dmax * R
    U = philox.thresholds(seed, T, C)
    ref = o.mh_normreg_walk(init, delta, U, x_obs, y_obs, lims, ex.astype(bool),
                            lg.astype(bool), has_slope=True, accept="log")
    yd, xd = dev(eng, y_obs), dev(eng, x_obs)
    st = dev(eng, init.T)
    out = eng.mh_normreg(st, yd, xd, T, lims, ex, lg, dmax, seed=seed, variant=variant)
    eng.sync()
    assert np.array_equal(host(out["accept_count"]), ref["u"].sum(axis=0))
    assert relerr(host(out["x"]), tcd_to_tdc(ref["x"])) <= 1e-11
    assert relerr(host(out["prob"]), ref["prob"]) <= 1e-11
    # resume in two halves + chain sharding reproduce the same stream
    st2 = dev(eng, init.T)
    a = eng.mh_normreg(st2, yd, xd, T // 2, lims, ex, lg, dmax, seed=seed, variant=variant)
    b = eng.mh_normreg(st2, yd, xd, T // 2, lims, ex, lg, dmax, seed=seed, variant=variant,
                       step0=T // 2, state_lp=a["state_lp"])
    eng.sync()
    assert np.array_equal(np.concatenate([host(a["x"]), host(b["x"])]), host(out["x"]))
    if C > 8:
        h = C // 2
        lo = eng.mh_normreg(dev(eng, init[:h].T), yd, xd, T, lims, ex, lg, dmax, seed=seed,
                            variant=variant)
        hi = eng.mh_normreg(dev(eng, init[h:].T), yd, xd, T, lims, ex, lg, dmax, seed=seed,
                            variant=variant, chain0=h)
        eng.sync()
        assert np.array_equal(np.concatenate([host(lo["x"]), host(hi["x"])], axis=2),
                              host(out["x"]))


def test_logjoint_matches_oracle_large_n():
    """The array density evaluation alone, N = 10^6 (BASELINE config C3 size): each
    log-joint within 1e-12 relative of the C restatement (sequential sum)."""
    from oracle.c import liboracle as lo
    eng = engine()
    rng = np.random.default_rng(2024)
    N = 1_000_000
    x_obs = rng.normal(0, 1, N)
    y_obs = rng.normal(-1 + 1.5 * x_obs, 0.5)
    lims = np.array([[-6., 6.], [-6., 6.], [0.001, 10.]])
    ex = np.array([[0, 0], [0, 0], [1, 0]]); lg = np.zeros(3, int)
    yd, xd = dev(eng, y_obs), dev(eng, x_obs)
    for C, variant in [(8, 2), (640, 1), (5000, 1)]:
        theta = np.stack([rng.normal(-1, .01, C), rng.normal(1.5, .01, C),
                          rng.uniform(.45, .55, C)], axis=1)
        got = host(eng.normreg_logjoint(dev(eng, theta.T), yd, xd, lims, ex, lg,
                                        variant=variant))
        idx = rng.choice(C, size=min(C, 24), replace=False)
        want = lo.normreg_logjoint(theta[idx], x_obs, y_obs, lims, ex, lg)
        assert relerr(got[idx], want) <= TOL


def test_linreg_posterior_moments_philox():
    """Native RNG + log rule at N = 20000: posterior mean/sd of (b0, b1, sigma)
    match the analytic large-N values within Monte Carlo error."""
    eng = engine()
    rng = np.random.default_rng(77)
    N, C = 20000, 512
    x_obs = rng.normal(0, 1, N)
    y_obs = rng.normal(-1 + 1.5 * x_obs, 0.5)
    lims = np.array([[-6., 6.], [-6., 6.], [0.001, 10.]])
    ex = np.array([[0, 0], [0, 0], [1, 0]]); lg = np.zeros(3, int)
    A = np.stack([np.ones(N), x_obs], axis=1)
    beta, res, *_ = np.linalg.lstsq(A, y_obs, rcond=None)
    sig = np.sqrt(res[0] / N)
    sd = sig / np.sqrt(N)
    dmax = np.array([2.5 * sd, 2.5 * sd, 2.0 * sd])
    st = dev(eng, np.tile([beta[0], beta[1], sig], (C, 1)).T)
    yd, xd = dev(eng, y_obs), dev(eng, x_obs)
    burn = eng.mh_normreg(st, yd, xd, 300, lims, ex, lg, dmax, seed=5, record=False)
    out = eng.mh_normreg(st, yd, xd, 1500, lims, ex, lg, dmax, seed=5, step0=300,
                         state_lp=burn["state_lp"], thin=5)
    eng.sync()
    X = host(out["x"])                                   # [R, 3, C]
    m = X.mean(axis=(0, 2)); s = X.std(axis=(0, 2))
    assert abs(m[0] - beta[0]) < 0.3 * sd and abs(m[1] - beta[1]) < 0.3 * sd
    assert abs(m[2] - sig) < 0.3 * sd
    assert 0.8 * sd < s[0] < 1.2 * sd and 0.8 * sd < s[1] < 1.2 * sd
    assert 0.8 * sd / np.sqrt(2) < s[2] < 1.2 * sd / np.sqrt(2)
    rate = host(out["accept_count"]).sum() / (C * 1500)
    assert 0.15 < rate < 0.6


def test_bound_native_rng_stays_inside_and_matches_oracle():
    """set_delta([d], bound=True) with the Philox stream, many chains: closed ends clip,
    open ends bounce (variable.py:700-727) -- replayed on the oracle with the same draws."""
    eng = engine()
    rng = np.random.default_rng(4)
    C, T, N, seed = 257, 60, 500, 99
    y = rng.normal(50., 10., N)
    lims = np.array([[40., 60.], [5., 20.]])
    ex = np.array([[0, 0], [1, 0]])
    lg = np.array([0, 1])
    scale = np.array([6.0, 0.5])
    init = np.stack([rng.uniform(41, 59, C), rng.uniform(6, 19, C)])
    out = eng.mh_normreg(dev(eng, init), dev(eng, y), None, T, lims, ex, lg, scale, seed=seed,
                         accept="log", prop="uniform", prop_bound=True, per_step=True)
    eng.sync()
    R = philox.uniforms(seed, T, C, 2)
    delta = -scale + 2.0 * scale * R
    U = philox.thresholds(seed, T, C)
    ref = o.mh_normreg_walk(init.T, delta, U, None, y, lims, ex, lg, has_slope=False,
                            accept="log", bound=True)
    assert np.array_equal(host(out["accept"]).astype(bool), ref["u"])
    x = host(out["x"])
    assert relerr(x, tcd_to_tdc(ref["x"])) <= 1e-11
    assert x[:, 0].min() >= 40. and x[:, 0].max() <= 60. and x[:, 1].min() > 5. and x[:, 1].max() <= 20.
