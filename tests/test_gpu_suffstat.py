"""The opt-in sufficient-statistics variants of K2 / K3 (C ABI variant 3,
pbx_grid_norm_logjoint_ss): same values as the term-by-term kernels and the oracle
to fp64 round-off, same accept decisions, for well- and badly-conditioned data."""
import numpy as np
import pytest
from conftest import load_golden, relerr
from gpu_util import engine, dev, tcd_to_tdc, host
from oracle import np_oracle as o
from oracle import philox

pytestmark = pytest.mark.gpu
TOL = 1e-12
LIMS3 = np.array([[-6., 6.], [-6., 6.], [0.001, 10.]])
EX3 = np.array([[0, 0], [0, 0], [1, 0]])
LG3 = np.zeros(3, int)


def _data(kind, N, rng):
    x = rng.normal(0., 1., N)
    if kind == "c3":                     # BASELINE config C3
        y = -1. + 1.5 * x + rng.normal(0., .5, N)
    elif kind == "offset":               # huge common offset: naive moments would cancel
        x = x + 1.0e6
        y = 3.0e5 + 1.5 * (x - 1.0e6) + rng.normal(0., .5, N)
    else:                                # near-perfect fit: RSS << Syy
        y = -1. + 1.5 * x + rng.normal(0., 1e-5, N)
    return x, y


@pytest.mark.parametrize("kind", ["c3", "offset", "tight"])
@pytest.mark.parametrize("N", [1, 7, 4097, 200001])
def test_regression_logjoint_matches_oracle(kind, N):
    eng = engine()
    rng = np.random.default_rng(N + len(kind))
    x, y = _data(kind, N, rng)
    C = 300
    b0 = (3.0e5 - 1.5e6 if kind == "offset" else -1.) + rng.normal(0, .01, C)
    lims = LIMS3.copy()
    lims[0] = [-2e6, 2e6]
    theta = np.stack([b0, 1.5 + rng.normal(0, .01, C), rng.uniform(.3, .7, C)])
    got = host(eng.normreg_logjoint(dev(eng, theta), dev(eng, y), dev(eng, x), lims, EX3, LG3,
                                    variant=3))
    want = o.normreg_logjoint(theta.T, x, y, lims, EX3, LG3, has_slope=True)
    # the oracle's own term-by-term sum carries |y|/sigma * eps per term for the offset data
    tol = 1e-9 if kind == "offset" else TOL
    assert relerr(got, want) <= tol
    if kind != "offset":
        stream = host(eng.normreg_logjoint(dev(eng, theta), dev(eng, y), dev(eng, x), lims, EX3,
                                           LG3, variant=1))
        assert relerr(got, stream) <= TOL


def test_offset_data_against_extended_precision():
    """Where the term-by-term fp64 sum itself loses digits (|x| ~ 1e6), the centred
    statistics stay at round-off of a long-double evaluation."""
    eng = engine()
    rng = np.random.default_rng(3)
    x, y = _data("offset", 50001, rng)
    theta = np.array([[3.0e5 - 1.5e6 + 0.01], [1.5003], [0.52]])
    lims = LIMS3.copy()
    lims[0] = [-2e6, 2e6]
    got = float(host(eng.normreg_logjoint(dev(eng, theta), dev(eng, y), dev(eng, x), lims, EX3,
                                          LG3, variant=3))[0])
    xl, yl = x.astype(np.longdouble), y.astype(np.longdouble)
    r = yl - np.longdouble(theta[0, 0]) - np.longdouble(theta[1, 0]) * xl
    sg = np.longdouble(theta[2, 0])
    want = float(-np.sum(r * r) / (2 * sg * sg) - len(x) * (np.log(np.sqrt(2 * np.pi)) + np.log(sg)))
    want += float(-np.log(4e6) - np.log(12.) - np.log(10. - 0.001))
    assert abs(got - want) <= 1e-12 * abs(want)


@pytest.mark.parametrize("name", ["mh_norm1d_hastings", "mh_norm1d_metropolis", "mh_linreg",
                                  "mh_norm1d_bound_mixed"])
def test_walk_golden_injected(name):
    """The live-reference fixtures through the one-launch walk: identical decisions."""
    eng = engine()
    g = load_golden(name)
    has_slope = name == "mh_linreg"
    T = len(g["thresh"])
    y = dev(eng, g["y_obs"] if has_slope else g["x_obs"])
    x = dev(eng, g["x_obs"]) if has_slope else None
    out = eng.mh_normreg(dev(eng, g["init"][:, None]), y, x, T, g["lims"], g["ex"], g["log_ufun"],
                         g["dmax"], accept="reference", accept_coef=float(g["coef"]), variant=3,
                         prop_bound=bool(g["bound"]) if "bound" in g.files else False,
                         inj_delta=dev(eng, tcd_to_tdc(g["delta"][:, None, :])),
                         inj_thresh=dev(eng, g["thresh"][:, None]), per_step=True)
    eng.sync()
    assert np.array_equal(host(out["accept"])[:, 0].astype(bool), g["u"])
    assert relerr(host(out["x"])[:, :, 0], g["x"]) <= TOL
    assert relerr(host(out["prob"])[:, 0], g["prob"]) <= TOL
    assert relerr(host(out["xprop"])[:, :, 0], g["xprop"]) <= TOL
    assert relerr(host(out["pprop"])[:, 0], g["pprop"]) <= TOL


def test_walk_native_rng_equals_streaming_walk():
    """Same Philox stream, same decisions and trajectory as the term-by-term K2 walk;
    resumable (two calls == one)."""
    eng = engine()
    rng = np.random.default_rng(12)
    N, C, T, seed = 20000, 700, 120, 5
    x, y = _data("c3", N, rng)
    sd = 0.5 / np.sqrt(N)
    init = np.tile(np.array([[-1.], [1.5], [.5]]), (1, C))
    kw = dict(seed=seed, accept="log", prop="uniform", per_step=True)
    a = eng.mh_normreg(dev(eng, init), dev(eng, y), dev(eng, x), T, LIMS3, EX3, LG3, [4 * sd] * 3,
                       variant=1, **kw)
    b = eng.mh_normreg(dev(eng, init), dev(eng, y), dev(eng, x), T, LIMS3, EX3, LG3, [4 * sd] * 3,
                       variant=3, **kw)
    eng.sync()
    assert np.array_equal(host(a["accept"]), host(b["accept"]))
    assert relerr(host(b["x"]), host(a["x"])) <= TOL
    assert relerr(host(b["prob"]), host(a["prob"])) <= TOL
    assert 0.02 < host(b["accept"]).mean() < 0.9
    st = dev(eng, init)
    h1 = eng.mh_normreg(st, dev(eng, y), dev(eng, x), 50, LIMS3, EX3, LG3, [4 * sd] * 3,
                        variant=3, seed=seed, accept="log", prop="uniform")
    h2 = eng.mh_normreg(st, dev(eng, y), dev(eng, x), T - 50, LIMS3, EX3, LG3, [4 * sd] * 3,
                        variant=3, seed=seed, accept="log", prop="uniform", step0=50,
                        state_lp=h1["state_lp"])
    both = np.concatenate([host(h1["x"]), host(h2["x"])])
    assert np.array_equal(both, host(b["x"]))


@pytest.mark.parametrize("name", ["dgei_small", "dgei_peaked"])
def test_grid_logjoint_golden(name):
    eng = engine()
    g = load_golden(name)
    M, S = len(g["mu"]), len(g["sigma"])
    lpm, lps = np.full(M, -np.log(20.)), np.full(S, -np.log(np.log(20.) - np.log(5.)))
    args = (dev(eng, g["data"]), dev(eng, g["mu"]), dev(eng, g["sigma"]), dev(eng, lpm),
            dev(eng, lps))
    ss = host(eng.grid_norm_logjoint(*args, suffstat=True))
    assert relerr(ss, g["joint"]) <= TOL
    assert relerr(ss, host(eng.grid_norm_logjoint(*args))) <= TOL


def test_grid_logjoint_full_size_agrees_on_a_slab():
    """C4 sizes: 4096^2 cells, N = 1e5; the O(N) kernel on a 64-row slab is the check."""
    eng = engine()
    rng = np.random.default_rng(7)
    N, M, S = 100000, 4096, 4096
    data = dev(eng, rng.normal(50., 10., N))
    mu = np.linspace(40, 60, M + 2)[1:-1]
    sg = np.exp(np.linspace(np.log(5), np.log(20), S + 2)[1:-1])
    lpm, lps = np.full(M, -np.log(20.)), np.full(S, -np.log(np.log(4.)))
    full = eng.grid_norm_logjoint(data, dev(eng, mu), dev(eng, sg), dev(eng, lpm), dev(eng, lps),
                                  suffstat=True)
    rows = slice(1000, 1064)
    slab = eng.grid_norm_logjoint(data, dev(eng, mu[rows]), dev(eng, sg), dev(eng, lpm[rows]),
                                  dev(eng, lps))
    assert relerr(host(full[rows]), host(slab)) <= TOL


def test_api_suffstat_options():
    """sampler(..., suffstat=True) and model(..., suffstat=True) through the public API give
    the same walk / grid as the default term-by-term paths."""
    engine()
    import scipy.stats
    import probayes_b200 as pb
    rng = np.random.default_rng(21)
    N = 5000
    xo = rng.normal(0., 1., N)
    yo = -1. + 1.5 * xo + rng.normal(0., .5, N)

    def norm_reg(x, y, beta_0, beta_1, y_sigma):
        return scipy.stats.norm.logpdf(y, loc=beta_0 + beta_1 * x, scale=y_sigma)

    def run(**extra):
        x = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
        y = pb.RV('y', vtype=float, vset=(-np.inf, np.inf))
        b0 = pb.RV('beta_0', vtype=float, vset=[-6., 6.], pscale='log')
        b1 = pb.RV('beta_1', vtype=float, vset=[-6., 6.], pscale='log')
        ys = pb.RV('y_sigma', vtype=float, vset=[(0.001,), 10.], pscale='log')
        paras = pb.RF(b0, b1, ys)
        sp = pb.SP(pb.RF(x, y), paras)
        sp.set_prob(norm_reg, pscale='log')
        paras.set_tran(lambda **k: 1.)
        paras.set_delta([0.02])
        sp.set_tran(paras)
        sp.set_delta(paras)
        sp.set_scores('metropolis')
        smp = sp.sampler({'beta_0': -1., 'beta_1': 1.5, 'y_sigma': .5}, {'x,y': [xo, yo]},
                         stop=400, iid=True, joint=True, chains=64, seed=9, **extra)
        return sp(sp.walk(smp))
    a, b = run(), run(suffstat=True)
    assert a.u.count(True) == b.u.count(True)
    for k in ('beta_0', 'beta_1', 'y_sigma'):
        assert relerr(b.v[k], a.v[k]) <= TOL
    assert relerr(b.v.prob, a.v.prob) <= TOL

    g = load_golden("dgei_small")
    mu = pb.RV('mu', vtype=float, vset=(40, 60))
    sigma = pb.RV('sigma', vtype=float, vset=(5, 20.))
    xr = pb.RV('x', vtype=float, vset={-np.inf, np.inf})
    sigma.set_ufun((np.log, np.exp))
    model = pb.SD(pb.RF(xr), pb.RF(mu, sigma))
    model.set_prob(scipy.stats.norm.logpdf, order={'x': 0, 'mu': 'loc', 'sigma': 'scale'},
                   pscale='log')
    M, S = len(g["mu"]), len(g["sigma"])
    joint = model({xr: g["data"], 'mu': {M}, 'sigma': {S}}, iid=True, joint=True, suffstat=True)
    assert joint.name == str(g["joint_name"]) and relerr(joint.prob, g["joint"]) <= TOL
    post = joint.conditionalise('x')
    assert np.abs(post.prob - g["posterior"])[g["posterior"] > -1e300].max() <= \
        TOL * np.abs(g["joint"]).max()
