"""Pins the numpy oracle against fixtures generated from the LIVE reference
(oracle/gen_golden.py): identical accept decisions, trajectories and densities
within 1e-12 relative (north_star tolerance for fp64)."""
import numpy as np
import pytest
from conftest import load_golden, relerr
from oracle import np_oracle as o

TOL = 1e-12


@pytest.mark.parametrize("name", ["mh_mvn_c1", "mh_mvn_c1_b", "mh_mvn_log", "mh_mvn_3d"])
def test_mh_mvn(name):
    g = load_golden(name)
    r = o.mh_mvn_walk(g["init"][None], g["delta"][:, None, :], g["thresh"][:, None],
                      g["mean"], g["cov"], log_pscale=bool(g["log_pscale"]))
    assert np.array_equal(r["u"][:, 0], g["u"])
    assert np.abs(r["x"][:, 0] - g["x"]).max() <= TOL
    assert relerr(r["prob"][:, 0], g["prob"]) <= TOL
    assert relerr(r["s"][1:, 0], g["s"][1:]) <= TOL
    assert relerr(r["xprop"][:, 0], g["xprop"]) <= TOL
    assert relerr(r["pprop"][:, 0], g["pprop"]) <= TOL
    assert np.isnan(g["s"][0]) and g["u"][0]          # step 1 accepts with s=None


def test_mh_mvn_bounded_delta():
    """set_delta([d], bound=True) on bounded variables with the mvn target
    (variable.py:700-739): closed limits clip, open limits bounce back."""
    g = load_golden("mh_mvn_bound")
    r = o.mh_mvn_walk(g["init"][None], g["delta"][:, None, :], g["thresh"][:, None],
                      g["mean"], g["cov"], bound=(g["lims"], g["ex"]))
    assert np.array_equal(r["u"][:, 0], g["u"])
    assert np.abs(r["x"][:, 0] - g["x"]).max() <= TOL
    assert relerr(r["prob"][:, 0], g["prob"]) <= TOL
    assert relerr(r["xprop"][:, 0], g["xprop"]) <= TOL
    assert (np.abs(g["xprop"][:, 0]) == 1.5).sum() > 20          # the fixture does clip


@pytest.mark.parametrize("name", ["mh_norm1d_hastings", "mh_norm1d_metropolis",
                                  "mh_norm1d_underflow", "mh_norm1d_bound_open",
                                  "mh_norm1d_bound_mixed"])
def test_mh_norm1d(name):
    g = load_golden(name)
    r = o.mh_normreg_walk(g["init"][None], g["delta"][:, None, :], g["thresh"][:, None],
                          None, g["x_obs"], g["lims"], g["ex"], g["log_ufun"],
                          has_slope=False, coef=float(g["coef"]),
                          bound=bool(g["bound"]) if "bound" in g.files else False)
    assert np.array_equal(r["u"][:, 0], g["u"])
    assert relerr(r["x"][:, 0], g["x"]) <= TOL
    assert relerr(r["prob"][:, 0], g["prob"]) <= TOL
    assert np.nanmax(np.abs(r["s"][1:, 0] - g["s"][1:])) <= TOL


def test_mh_norm1d_underflow_is_degenerate():
    """at N=1000 the reference's linear-space ratio underflows and
    only the first step is accepted; the log-space rule does accept."""
    g = load_golden("mh_norm1d_underflow")
    assert g["u"].sum() == 1
    r = o.mh_normreg_walk(g["init"][None], g["delta"][:, None, :], g["thresh"][:, None],
                          None, g["x_obs"], g["lims"], g["ex"], g["log_ufun"],
                          has_slope=False, accept="log")
    assert r["u"].sum() > 10


def test_mh_linreg():
    g = load_golden("mh_linreg")
    r = o.mh_normreg_walk(g["init"][None], g["delta"][:, None, :], g["thresh"][:, None],
                          g["x_obs"], g["y_obs"], g["lims"], g["ex"], g["log_ufun"],
                          has_slope=True)
    assert np.array_equal(r["u"][:, 0], g["u"])
    assert relerr(r["x"][:, 0], g["x"]) <= TOL
    assert relerr(r["prob"][:, 0], g["prob"]) <= TOL
    # where the linear ratio does not underflow the log rule decides identically
    r2 = o.mh_normreg_walk(g["init"][None], g["delta"][:, None, :], g["thresh"][:, None],
                           g["x_obs"], g["y_obs"], g["lims"], g["ex"], g["log_ufun"],
                           has_slope=True, accept="log")
    assert np.array_equal(r2["u"], r["u"])


@pytest.mark.parametrize("name", ["dgei_small", "dgei_peaked"])
def test_dgei(name):
    g = load_golden(name)
    M, S = len(g["mu"]), len(g["sigma"])
    mu = o.uniform_grid(40, 60, M, True, True)
    sg = np.exp(o.uniform_grid(np.log(5), np.log(20), S, True, True))
    assert np.array_equal(mu, g["mu"]) and np.array_equal(sg, g["sigma"])
    lj = o.grid_norm_logjoint(g["data"], mu, sg, np.full(M, -np.log(20.)),
                              np.full(S, -np.log(np.log(20) - np.log(5))))
    assert relerr(lj, g["joint"]) <= TOL
    post = o.grid_conditionalise(lj)
    assert relerr(post, g["posterior"]) <= TOL
    assert relerr(o.grid_marginal(post, 1), g["marg_mu"]) <= TOL
    assert relerr(o.grid_marginal(post, 0), g["marg_sigma"]) <= TOL
    assert abs(o.grid_expectation(post, mu, 0) - g["expt_mu"]) <= 1e-12 * 60
    assert abs(o.grid_expectation(post, sg, 1) - g["expt_sigma"]) <= 1e-12 * 20
    if name == "dgei_peaked":       # clamped cells: log_prob(<tiny) = -1.797e308
        assert (g["posterior"] == o.NEARLY_NEGATIVE_INF).sum() > 0
        assert np.array_equal(post == o.NEARLY_NEGATIVE_INF,
                              g["posterior"] == o.NEARLY_NEGATIVE_INF)


def test_mvn_value_order_quirk_is_pinned():
    """d = 3: the reference evaluates scipy's mvn on [v1, v0, v2] (prob.py:349-358);
    the natural order gives different accept decisions."""
    g = load_golden("mh_mvn_3d")
    assert o.mvn_value_order(3) == [1, 0, 2] and o.mvn_value_order(2) == [1, 0]
    assert o.mvn_value_order(5) == [3, 2, 1, 0, 4]
    r = o.mh_mvn_walk(g["init"][None], g["delta"][:, None, :], g["thresh"][:, None],
                      g["mean"], g["cov"], reorder=False)
    assert not np.array_equal(r["u"][:, 0], g["u"])


@pytest.mark.parametrize("name", ["gibbs2d", "gibbs3d"])
def test_gibbs_nd(name):
    g = load_golden(name)
    r = o.gibbs_mvn_walk(g["init"][None], g["runif"][:, None], g["mean"], g["cov"],
                         g["lims"])
    assert np.abs(r["x"][:, 0] - g["x"]).max() <= TOL
    assert relerr(r["prob"][:, 0], g["prob"]) <= TOL


def test_gibbs2d():
    g = load_golden("gibbs2d")
    r = o.gibbs_mvn_walk(g["init"][None], g["runif"][:, None], g["mean"], g["cov"],
                         g["lims"])
    assert np.abs(r["x"][:, 0] - g["x"]).max() <= TOL
    assert relerr(r["prob"][:, 0], g["prob"]) <= TOL
    assert int(g["n_true"]) == len(g["runif"])          # gibbs always updates


@pytest.mark.parametrize("name", ["condcov_d8", "condcov_d64"])
def test_condcov(name):
    g = load_golden(name)
    cc = o.CondCovOracle(g["mean"], g["cov"], g["lims"])
    assert relerr(cc.stdv, g["stdv"]) <= TOL
    assert np.abs(cc.coef - g["coef"]).max() <= TOL
    assert np.abs(cc.cdfs - g["cdfs"]).max() <= TOL
    # precision-matrix identities the device path relies on
    P = np.linalg.inv(g["cov"])
    for i in range(cc.n):
        idx = [j for j in range(cc.n) if j != i]
        assert np.allclose(cc.coef[i, idx], -P[i, idx] / P[i, i], rtol=1e-9, atol=1e-12)
        assert np.isclose(cc.stdv[i], P[i, i] ** -0.5, rtol=1e-10)
    x = g["init"][None].copy()
    for k in range(len(g["runif"])):
        x = cc.step(x, k % cc.n, g["runif"][k:k + 1])
        assert np.abs(x[0] - g["x"][k]).max() <= 1e-11


def test_pscales_table():
    g = load_golden("pscales")
    assert np.array_equal(o.log_prob(g["p"]), g["log_prob"])
    assert np.array_equal(o.exp_logp(g["l"]), g["exp_logp"])
    with np.errstate(over="ignore"):
        assert np.array_equal(o.div_prob_linear(g["num"], g["den"]), g["div_lin"])
        assert np.array_equal(
            o.div_prob_linear(o.exp_logp(g["lnum"]), o.exp_logp(g["lden"])),
            g["div_log_to_lin"])
    assert np.array_equal(o.from_linear(g["p"], True), g["resc_lin_to_log"])
    assert np.array_equal(o.to_linear(g["l"], True), g["resc_log_to_lin"])


OMC_LIMS = np.array([[40., 60.], [5., 20.]])
OMC_EX = np.array([[1, 1], [1, 1]])
OMC_LOG = np.array([0, 1])


def test_omc_random_sampling():
    """examples/omc/omc_rs_sp_norm1d.py: prior draws, log-joint, summary name."""
    g = load_golden("omc_rs_norm1d")
    th = o.box_sample(OMC_LIMS, OMC_LOG, g["runif"])
    assert np.array_equal(th[0], g["mu"]) and np.array_equal(th[1], g["sigma"])
    lj = o.normreg_logjoint(th.T, None, g["data"], OMC_LIMS, OMC_EX, OMC_LOG, has_slope=False)
    assert relerr(lj, g["logp"]) <= TOL
    assert relerr(o.exp_logp(lj), g["lin"]) <= 1e-11       # exp amplifies |lj| * eps
    assert str(g["name"]) == "mu,sigma,x={%d}" % (g["data"].size * g["mu"].size)


def test_pd_sorted_quantile_expectation():
    """PD.sorted / quantile / expectation on the OMC summary (pd.py:373-493)."""
    g = load_golden("omc_rs_norm1d")
    order = o.pd_sorted_order(g["mu"])
    assert np.array_equal(g["mu"][order], g["mu_sorted"])
    assert np.array_equal(g["sigma"][order], g["mu_sorted_sigma"])
    assert np.array_equal(g["lin"][order], g["mu_sorted_prob"])
    q = o.pd_quantile_1d(g["mu_sorted"], g["mu_sorted_prob"], False, g["qs"])
    assert relerr(q, g["q_mu"]) <= TOL
    so = o.pd_sorted_order(g["sigma"])
    assert np.array_equal(g["sigma"][so], g["sigma_sorted"])
    q = o.pd_quantile_1d(g["sigma_sorted"], g["lin"][so], False, g["qs"])
    assert relerr(q, g["q_sigma"]) <= TOL
    # the same quantiles straight from the log-pscale summary (rescale folded in)
    q = o.pd_quantile_1d(g["mu_sorted"], g["logp"][order], True, g["qs"])
    assert relerr(q, g["q_mu"]) <= 1e-10
    tot, _, cv = o.pd_expectation(g["lin"], False, col_vals=[g["mu"], g["sigma"],
                                                             g["mu"] ** 2, g["sigma"] ** 2])
    e = np.array(cv) / max(o.NEARLY_POSITIVE_ZERO, tot)
    assert relerr(e[:2], g["expt"]) <= TOL and relerr(e[2:], g["expt2"]) <= TOL


def test_pd_product_and_division_algebra():
    """prod_rule / div_prob with broadcasting on the DGEI pieces (pscales.py:160-236)."""
    g = load_golden("pd_algebra")
    p, lg = o.pd_product(g["pmu"][:, None], False, g["psg"][None, :], False)
    assert not lg and np.array_equal(p, g["prior"]) and np.array_equal(p, g["pmu_psg"])
    j, lg = o.pd_product(g["prior"], False, g["like"], True)
    assert lg and np.array_equal(j, g["joint"]) and np.array_equal(j, g["prior_like"])
    assert np.array_equal(o.pd_divide(g["joint"], True, g["evidence"], True), g["post"])
    assert np.array_equal(o.pd_divide(g["joint"], True, g["marg_mu_x"][:, None], True),
                          g["cond_sigma"])
