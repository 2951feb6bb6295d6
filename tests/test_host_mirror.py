"""Host-side mirror of the reference interface (no GPU needed): pscales, RV domain
semantics, grids, priors, PD algebra on host arrays, catalogue recognition and
error behaviour -- checked against the golden fixtures and, when the live
reference is present (development container), differentially against it."""
import numpy as np
import pytest
import scipy.stats
from conftest import load_golden, relerr
import probayes_b200 as pb
from probayes_b200 import catalogue
from oracle import ref_shim


def test_pscales_table_matches_reference():
    g = load_golden("pscales")
    assert np.array_equal(pb.log_prob(g["p"]), g["log_prob"])
    assert np.array_equal(pb.exp_logp(g["l"]), g["exp_logp"])
    assert np.array_equal(pb.div_prob(g["num"], g["den"]), g["div_lin"])
    assert np.array_equal(pb.div_prob(g["lnum"], g["lden"], 0j, 0j, pscale=1.),
                          g["div_log_to_lin"])
    assert np.array_equal(pb.rescale(g["p"], 1., 0j), g["resc_lin_to_log"])
    assert np.array_equal(pb.rescale(g["l"], 0j, 1.), g["resc_log_to_lin"])


def test_eval_pscale_and_products():
    assert pb.eval_pscale(None) == 1. and pb.eval_pscale(1) == 1.
    assert pb.eval_pscale('log') == 0j and pb.eval_pscale(0) == 0j and pb.eval_pscale('ln') == 0j
    assert pb.eval_pscale(2.5) == 2.5 and pb.eval_pscale(3 + 0j) == 3 + 0j
    with pytest.raises(ValueError):
        pb.eval_pscale('linear')
    assert pb.prod_pscale([1., 1.]) == 1. and pb.prod_pscale([0j, 1.]) == 0j
    p, ps = pb.prod_rule(np.log(0.2), np.log(0.5), pscales=[0j, 0j])
    assert ps == 0j and np.isclose(p, np.log(0.1))
    p, ps = pb.prod_rule(0.2, np.log(0.5), pscales=[1., 0j])
    assert ps == 0j and np.isclose(p, np.log(0.1))
    p, ps = pb.prod_rule(0.2, 0.5, pscales=[1., 1.])
    assert ps == 1. and np.isclose(p, 0.1)


def test_rv_domain_semantics():
    a = pb.RV('a', vtype=float, vset=(40, 60))                 # tuple: both ends open
    assert a.vset == [(40.,), (60.,)] and a.open_ends == (True, True)
    assert not a.inside(40.) and a.inside(50.) and not a.inside(60.)
    b = pb.RV('b', vtype=float, vset=[-6., 6.])                # list: closed
    assert b.inside(-6.) and b.inside(6.) and b.length == 12.
    c = pb.RV('c', vtype=float, vset=[(0.001,), 10.])          # mixed
    assert c.open_ends == (True, False) and not c.inside(0.001) and c.inside(10.)
    d = pb.RV('d', vtype=float, vset=[(0.001), 10.])           # (0.001) is NOT a tuple
    assert d.open_ends == (False, False)
    e = pb.RV('e')
    assert np.isinf(e.length) and not e.isfinite
    with pytest.raises(NotImplementedError):
        pb.RV('k', vtype=int, vset=[0, 1])
    s = pb.RV('s', vtype=float, vset=(5, 20.), pscale='log')
    s.set_ufun((np.log, np.exp))
    assert np.isclose(s.length, np.log(4.)) and s.log_ufun
    assert np.isclose(s.eval_prob(7.), -np.log(np.log(4.)))     # no Jacobian
    assert s.eval_prob(25.) == pb.NEARLY_NEGATIVE_INF
    lin = pb.RV('q', vtype=float, vset=[0., 4.])
    assert np.isclose(lin.eval_prob(1.), 0.25) and lin.eval_prob(5.) == 0.
    with pytest.raises(NotImplementedError):
        s.set_ufun((np.sqrt, np.square))


def test_grids_match_golden():
    g = load_golden("dgei_small")
    mu = pb.RV('mu', vtype=float, vset=(40, 60))
    sg = pb.RV('sigma', vtype=float, vset=(5, 20.))
    sg.set_ufun((np.log, np.exp))
    assert np.array_equal(mu.evaluate({len(g["mu"])}), g["mu"])
    assert np.array_equal(sg.evaluate({len(g["sigma"])}), g["sigma"])
    assert np.array_equal(pb.uniform(0, 1, 4), np.linspace(0, 1, 4))
    assert np.array_equal(pb.uniform(0, 1, 1), [0.5])
    assert np.array_equal(pb.uniform(0, 1, 3, True, False), np.linspace(0, 1, 4)[1:])
    assert np.array_equal(pb.uniform(0, 1, 3, False, True), np.linspace(0, 1, 4)[:-1])


def test_pd_host_algebra_matches_golden():
    """PD.conditionalise / marginal / expectation / quantile / rescaled on a
    host-backed PD reproduce the reference's outputs (dgei_small)."""
    g = load_golden("dgei_small")
    N = len(g["data"])
    joint = pb.PD("mu=[],sigma=[],x={{{}}}".format(N),
                  {'mu': g["mu"], 'sigma': g["sigma"], 'x': {N}},
                  dims={'mu': 0, 'sigma': 1, 'x': None}, prob=g["joint"], pscale='log')
    assert joint.name == str(g["joint_name"])
    post = joint.conditionalise('x')
    assert post.name == str(g["post_name"])
    assert relerr(post.prob, g["posterior"]) <= 1e-12
    pm, psg = post.marginal('mu'), post.marginal('sigma')
    assert pm.name == str(g["marg_mu_name"])
    assert relerr(pm.prob, g["marg_mu"]) <= 1e-12 and relerr(psg.prob, g["marg_sigma"]) <= 1e-12
    ex = post.expectation()
    assert abs(ex['mu'] - g["expt_mu"]) <= 1e-10 and abs(ex['sigma'] - g["expt_sigma"]) <= 1e-10
    assert abs(pm.quantile()['mu'] - g["med_mu"]) <= 1e-10
    assert abs(psg.quantile()['sigma'] - g["med_sigma"]) <= 1e-10
    q3 = pm.quantile([0.025, 0.5, 0.975])
    assert np.abs(np.array([q['mu'] for q in q3]) - g["q3_mu"]).max() <= 1e-10
    assert relerr(post.rescaled().prob, g["post_linear"]) <= 1e-12
    srt = pm.sorted('mu')
    assert np.array_equal(srt['mu'], np.sort(g["mu"]))
    iid = pb.PD("x=[]", {'x': np.arange(3.)}, prob=np.log([.1, .2, .3]), pscale='log').prod('x')
    assert iid.name == "x={3}" and np.isclose(iid.prob, np.log(.006))


def test_pd_host_postprocessing_matches_omc_golden():
    """PD.sorted / quantile (incl. the {size} answer for non-monotonic values) /
    expectation(exponent=) on a host-backed sample-set PD reproduce the reference's
    outputs on the OMC fixture (pd.py:373-493)."""
    g = load_golden("omc_rs_norm1d")
    T, n = len(g["mu"]), len(g["data"]) * len(g["mu"])
    inf = pb.PD(str(g["name"]), {'mu': g["mu"], 'sigma': g["sigma"], 'x': {n}},
                dims={'mu': 0, 'sigma': 0, 'x': None}, prob=g["lin"], pscale=1.)
    assert inf.name == str(g["name"]) and inf.shape == [T]
    ms = inf.sorted('mu')
    assert np.array_equal(ms['mu'], g["mu_sorted"])
    assert np.array_equal(ms['sigma'], g["mu_sorted_sigma"])
    assert np.array_equal(ms.prob, g["mu_sorted_prob"])
    q = ms.quantile(g["qs"].tolist())
    assert relerr([v['mu'] for v in q], g["q_mu"]) <= 1e-12
    assert q[2]['sigma'] == {T} and q[2]['x'] == {n}
    assert abs(ms.quantile(0.5)['mu'] - g["med_mu"]) <= 1e-12 * g["med_mu"]
    ss = inf.sorted('sigma')
    assert np.array_equal(ss.prob, g["sigma_sorted_prob"])
    assert relerr([v['sigma'] for v in ss.quantile(g["qs"].tolist())], g["q_sigma"]) <= 1e-12
    e = inf.expectation()
    assert relerr([e['mu'], e['sigma']], g["expt"]) <= 1e-12 and e['x'] == {n}
    e2 = inf.expectation(['mu', 'sigma'], exponent=2)
    assert relerr([e2['mu'], e2['sigma']], g["expt2"]) <= 1e-12
    # 2-D: mu (axis 0) takes the cell's value, sigma (last axis) is interpolated
    d = load_golden("dgei_small")
    post = pb.PD("mu=[],sigma=[]|x={60}", {'mu': d["mu"], 'sigma': d["sigma"], 'x': {60}},
                 dims={'mu': 0, 'sigma': 1, 'x': None}, prob=d["posterior"], pscale='log')
    q2 = post.quantile(0.5)
    assert q2['mu'] in d["mu"] and d["sigma"].min() <= q2['sigma'] <= d["sigma"].max()


def test_pd_product_division_match_reference():
    """RV / RF prior PDs, PD.__mul__ (pd_utils.product) and PD.__truediv__ on host-backed
    PDs reproduce the reference's names, dims and arrays (pd_algebra fixture)."""
    g = load_golden("pd_algebra")
    names = [str(n) for n in g["names"]]
    M, S, N = len(g["mu"]), len(g["sigma"]), len(g["data"])
    mu = pb.RV('mu', vtype=float, vset=(40, 60))
    sigma = pb.RV('sigma', vtype=float, vset=(5, 20.))
    sigma.set_ufun((np.log, np.exp))
    pmu, psg = mu({M}), sigma({S})
    assert pmu.name == names[0] and np.array_equal(pmu.prob, g["pmu"]) and pmu.pscale == 1.
    assert np.array_equal(pmu['mu'], g["mu"]) and relerr(psg['sigma'], g["sigma"]) <= 1e-15
    prior = pb.RF(mu, sigma)({'mu': {M}, 'sigma': {S}})
    pp = pmu * psg
    for d in (prior, pp):
        assert d.name == names[1] and d.shape == [M, S] and d.pscale == 1.
        assert dict(d.dims) == {'mu': 0, 'sigma': 1} and np.array_equal(d.prob, g["prior"])
    like = pb.PD(names[3], {'mu': g["mu"], 'sigma': g["sigma"], 'x': {N}},
                 dims={'mu': 0, 'sigma': 1, 'x': None}, prob=g["like"], pscale='log')
    joint = prior * like
    assert joint.name == names[4] and joint.pscale == 0j and joint['x'] == {N}
    assert np.array_equal(joint.prob, g["prior_like"])
    assert (like * prior).name == 'x={%d},mu=[],sigma=[]' % N
    ev = joint.marginal('x')
    assert ev.name == names[5] and abs(ev.prob - g["evidence"]) <= 1e-12 * abs(g["evidence"])
    post = joint / ev
    assert post.name == names[6] and relerr(post.prob, g["post"]) <= 1e-12
    mm = joint.marginal(['mu', 'x'])
    assert mm.name == names[7] and relerr(mm.prob, g["marg_mu_x"]) <= 1e-12
    pc = joint / mm
    assert pc.name == names[8] and dict(pc.dims) == {'sigma': 1, 'mu': 0, 'x': None}
    assert relerr(pc.prob, g["cond_sigma"]) <= 1e-12
    # scalars: product of two scalar PDs, mismatching marginals refused
    a = pb.PD('a=1.0', {'a': 1.0}, prob=0.25)
    b = pb.PD('b=2.0|a=1.0', {'b': 2.0, 'a': 1.0}, prob=0.5)
    ab = a * b
    assert ab.name == 'a=1.0,b=2.0' and ab.prob == 0.125
    with pytest.raises(AssertionError, match="Non-unique"):
        pmu * pmu


def test_sp_random_sampling_recognition():
    """The proposal-free sampler with {0} requests is the OMC mode (sp.py:227-234);
    {n != 0} and a configured delta are not."""
    mu = pb.RV('mu', vtype=float, vset=(40, 60))
    sigma = pb.RV('sigma', vtype=float, vset=(5, 20.))
    x = pb.RV('x', vtype=float, vset=[-np.inf, np.inf])
    sigma.set_ufun((np.log, np.exp))
    sp = pb.SP(pb.RF(x), pb.RF(mu, sigma))
    sp.set_prob(scipy.stats.norm.logpdf, order={'x': 0, 'mu': 'loc', 'sigma': 'scale'},
                pscale='log')
    data = np.arange(5.)
    s = sp.sampler({'mu': {0}, 'sigma': {0}, 'x': data}, iid=True, joint=True, stop=10)
    assert s.opts['omc'] and list(s.obs.keys()) == ['x']
    with pytest.raises(NotImplementedError, match="one value per step"):
        sp.sampler({'mu': {3}, 'sigma': {0}, 'x': data}, stop=10)
    s2 = sp.sampler({'mu': 50., 'sigma': 10.}, {'x': data}, stop=10)
    assert not s2.opts['omc']
    sp.set_delta([0.5])
    assert not sp.sampler({'mu': {0}, 'sigma': {0}, 'x': data}, stop=10).opts['omc']


def _mvn_sp():
    x = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
    y = pb.RV('y', vtype=float, vset=(-np.inf, np.inf))
    sp = pb.SP(x & y)
    sp.set_prob(scipy.stats.multivariate_normal, [0., 0.], [[2.0, 1.2], [1.2, 2.0]])
    return sp


def test_catalogue_recognition():
    sp = _mvn_sp()
    spec = catalogue.identify_target(sp, sp.leafs, sp.roots)
    assert spec['kind'] == 'mvn' and spec['names'] == ['x', 'y'] and not spec['log_pscale']
    sp.set_delta(scipy.stats.norm(0., 0.7))
    pr = catalogue.identify_proposal(sp._proposal_rf(), sp.pscale)
    assert pr['kind'] == 'normal' and np.allclose(pr['scale'], 0.7)
    sp.set_delta(lambda: None)
    with pytest.raises(NotImplementedError):
        catalogue.identify_proposal(sp._proposal_rf(), sp.pscale)
    catalogue.identify_proposal(sp._proposal_rf(), sp.pscale, injected=True)
    sp.set_tran(np.array([[0.5, 0.2], [0.2, 0.8]]))
    sp.set_delta(scipy.stats.norm(0., 1.))
    pr = catalogue.identify_proposal(sp._proposal_rf(), sp.pscale)
    assert np.allclose(pr['chol'], np.linalg.cholesky([[0.5, 0.2], [0.2, 0.8]]))
    sp.set_prob(lambda x, y: x + y)
    with pytest.raises(NotImplementedError):
        catalogue.identify_target(sp, sp.leafs, sp.roots)


def test_catalogue_normal_likelihoods():
    mu = pb.RV('mu', vtype=float, vset=(40, 60), pscale='log')
    sigma = pb.RV('sigma', vtype=float, vset=(5, 20.), pscale='log')
    x = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
    sigma.set_ufun((np.log, np.exp))
    paras, stats = pb.RF(mu, sigma), pb.RF(x)
    sp = pb.SP(stats, paras)
    sp.set_prob(scipy.stats.norm.logpdf, order={'x': 0, 'mu': 'loc', 'sigma': 'scale'})
    spec = catalogue.identify_target(sp, sp.leafs, sp.roots)
    assert spec == dict(kind='normreg', has_slope=False, obs_y='x', obs_x=None,
                        params=['mu', 'sigma'], log_pscale=True)
    tran = lambda **k: 1.
    paras.set_tran((tran, tran))
    paras.set_delta((0.005,), scale=True)
    sp.set_tran(paras)
    sp.set_delta(paras)
    pr = catalogue.identify_proposal(sp._proposal_rf(), sp.pscale)
    g = load_golden("mh_norm1d_spherical")
    assert pr['kind'] == 'spherical' and np.isclose(pr['radius'], float(g["radius"]))
    assert np.allclose(pr['scale'], g["lengths"]) and np.isclose(pr['coef'], np.e)
    paras.set_delta([0.005], scale=True)
    pr = catalogue.identify_proposal(sp._proposal_rf(), sp.pscale)
    assert pr['kind'] == 'uniform' and np.allclose(pr['scale'], load_golden("mh_norm1d_hastings")["dmax"])
    sp.set_prob(scipy.stats.norm.pdf, order={'x': 0, 'mu': 'loc', 'sigma': 'scale'})
    with pytest.raises(NotImplementedError):
        catalogue.identify_target(sp, sp.leafs, sp.roots)

    # a user-written regression log-likelihood is recognised by probing it
    xx = pb.RV('x', vtype=float, vset=[-3, 3])
    yy = pb.RV('y', vtype=float, vset=[-np.inf, np.inf])
    b0 = pb.RV('beta_0', vtype=float, vset=[-6., 6.], pscale='log')
    b1 = pb.RV('beta_1', vtype=float, vset=[-6., 6.], pscale='log')
    ys = pb.RV('y_sigma', vtype=float, vset=[(0.001), 10.], pscale='log')

    def norm_reg(x, y, beta_0, beta_1, y_sigma):
        return scipy.stats.norm.logpdf(y, loc=beta_0 + beta_1 * x, scale=y_sigma)

    proc = pb.SP(xx & yy, b0 & b1 & ys)
    proc.set_prob(norm_reg, pscale='log')
    spec = catalogue.identify_target(proc, proc.leafs, proc.roots)
    assert spec['kind'] == 'normreg' and spec['has_slope']
    assert (spec['obs_x'], spec['obs_y']) == ('x', 'y')
    assert spec['params'] == ['beta_0', 'beta_1', 'y_sigma']
    proc.set_prob(lambda x, y, beta_0, beta_1, y_sigma: -np.abs(y - beta_0), pscale='log')
    with pytest.raises(NotImplementedError):
        catalogue.identify_target(proc, proc.leafs, proc.roots)


def test_sp_interface_errors_and_registry():
    sp = _mvn_sp()
    assert pb.MCMC_SAMPLERS == ('metropolis', 'hastings', 'gibbs')
    sp.set_scores('hastings')
    assert sp.scores == sp.thresh == sp.update == 'hastings'
    sp.set_update('metropolis')
    assert sp.update == 'metropolis'
    with pytest.raises(NotImplementedError):
        sp.set_scores(lambda opqr: 1.)
    with pytest.raises(AssertionError):
        sp.set_scores('metropolis', 3)
    with pytest.raises(NotImplementedError):
        sp.sampler(stop=10)
    s = sp.sampler({'x': 0., 'y': 1.}, stop=10)
    assert sp.get_counter(0) == 0 and sp.get_sampler(0) is s
    with pytest.raises(ValueError):
        sp.walk(sp.sampler({'x': 0., 'y': 1.}))            # no stop


def test_condcov_mirror_matches_reference_constants():
    g = load_golden("condcov_d8")
    cc = pb.CondCov(g["mean"], g["cov"], g["lims"])
    assert relerr(cc.stdv, g["stdv"]) <= 1e-12
    assert np.abs(cc.coef_matrix() - g["coef"]).max() <= 1e-12
    assert np.abs(cc.cdfs - g["cdfs"]).max() <= 1e-12
    with pytest.raises(AssertionError):
        pb.CondCov([0., 0.], np.eye(3), [[-1, 1]] * 2)


@pytest.mark.skipif(not ref_shim.available(), reason="live reference not present")
def test_differential_against_live_reference():
    """RV lengths / priors / grids / pscales helpers against the live reference."""
    ref = ref_shim.load()
    from probayes import pscales as rps
    rng = np.random.default_rng(0)
    for vset, ufun in [((40, 60), None), ((5, 20.), (np.log, np.exp)), ([-6., 6.], None),
                       ([(0.001,), 10.], None), ([0.5, (3.,)], (np.log, np.exp))]:
        for pscale in (None, 'log'):
            a = ref.RV('v', vtype=float, vset=vset, pscale=pscale)
            b = pb.RV('v', vtype=float, vset=vset, pscale=pscale)
            if ufun:
                a.set_ufun(ufun)
                b.set_ufun(ufun)
            assert np.isclose(a.length, b.length, rtol=1e-15)
            lo, hi = a.vlims
            pts = np.concatenate([[lo, hi], rng.uniform(lo - 1, hi + 1, 12)])
            assert np.array_equal(np.asarray(a.eval_prob(pts)), np.asarray(b.eval_prob(pts)))
            for n in (1, 2, 7):
                ga = np.ravel(a.evaluate({n})[a.name] if isinstance(a.evaluate({n}), dict)
                              else a.evaluate({n}))
                assert np.allclose(ga, b.evaluate({n}), rtol=1e-15, atol=0)
    vals = rng.uniform(-800, 5, 40)
    assert np.array_equal(rps.rescale(vals, 0j, 1.), pb.rescale(vals, 0j, 1.))
    assert np.array_equal(rps.rescale(np.exp(vals), 1., 0j), pb.rescale(np.exp(vals), 1., 0j))
    assert rps.prod_pscale([0j, 2.0]) == pb.prod_pscale([0j, 2.0])


def test_conditionalise_on_array_keys_host():
    """Host-backed PDs: PD.conditionalise on array-valued keys reproduces the live
    reference bit for bit (pd.py:214-295; pd_cond_array fixture)."""
    import probayes_b200 as pb
    g = load_golden("pd_cond_array")
    M, S = len(g["mu"]), len(g["sigma"])
    vals = {'mu': g["mu"].reshape(M, 1), 'sigma': g["sigma"].reshape(1, S), 'x': {len(g["data"])}}
    joint = pb.PD('mu,sigma,x', vals, dims={'mu': 0, 'sigma': 1, 'x': None}, prob=g["joint"],
                  pscale='log')
    for i, (keys, k) in enumerate([('sigma', "c_sig"), ('mu', "c_mu"), (['sigma', 'x'], "c_sx")]):
        c = joint.conditionalise(keys)
        assert np.array_equal(c.prob, g[k])
        assert [c.dims['mu'], c.dims['sigma']] == list(g["dims"][i])
    lin = joint.conditionalise('x').rescaled()
    assert np.array_equal(lin.prob, g["lin"])
    assert np.array_equal(lin.conditionalise('sigma').prob, g["l_sig"])
    assert np.array_equal(lin.conditionalise('mu').prob, g["l_mu"])


def test_rejection_catalogue_recognition():
    """The rejection-sampling pieces of omc_rejection_sp_circle.py are mapped onto kernel
    modes by probing; anything else is refused (no Python runs in the kernel)."""
    import warnings
    radius = 1.5
    inside = lambda x, y: np.array(x**2 + y**2 <= radius**2, dtype=float)

    def norm2d(x, y, loc=0.25, scale=0.75):
        return scipy.stats.norm.pdf(x, loc=loc, scale=scale) * \
            scipy.stats.norm.pdf(y, loc=loc, scale=scale)
    rvs = [pb.RV("x", [-2., 2.]), pb.RV("y", [-2., 2.])]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        t = catalogue.identify_rejection_target(inside, (), {}, rvs)
        q = catalogue.identify_rejection_prop(norm2d, (), {}, rvs)
    assert len(w) == 2
    assert t['kind'] == 'ball' and t['radius'] == radius and list(t['centre']) == [0., 0.]
    assert q['kind'] == 'normal' and list(q['loc']) == [0.25, 0.25] and list(q['scale']) == [0.75] * 2
    assert catalogue.identify_scores(lambda opqr: opqr.p.prob) == 'p'
    assert catalogue.identify_scores(lambda opqr: opqr.p.prob / opqr.q.prob) == 'p/q'
    assert catalogue.identify_thresh(np.random.uniform, (), dict(low=0., high=0.3)) == \
        ('uniform', 0., 0.3)
    assert catalogue.identify_update(lambda stu: stu.s >= stu.t) == 's>=t'
    with pytest.raises(NotImplementedError):
        catalogue.identify_scores(lambda opqr: opqr.p.prob ** 2)
    with pytest.raises(NotImplementedError):
        catalogue.identify_update(lambda stu: stu.s > stu.t)
    with pytest.raises(NotImplementedError):
        catalogue.identify_rejection_target(lambda x, y: np.array(abs(x) + abs(y) <= 1., float),
                                            (), {}, rvs)
    with pytest.raises(NotImplementedError):
        catalogue.identify_thresh(np.random.normal, (), {})


def test_table_ndtri_host_mirror_against_scipy():
    """The Gibbs kernel's table-driven ndtri (pbx_ndtri.cuh), evaluated by the library's host
    mirror of the same table and arithmetic (no GPU): within 3e-15 of scipy's ndtri relative
    to max(|x|, 1e-3) over the centre, both tails down to 2^-63, and the grid ends of the
    52-bit uniforms (degree-5 polynomials on 64 segments per binade); NaN outside the table (where the device calls normcdfinv)."""
    import ctypes as C
    from scipy.special import ndtri
    from probayes_b200 import build, _lib
    build.build()
    lib = _lib.load()
    rng = np.random.default_rng(0)
    u = np.concatenate([rng.random(200000), 2.0 ** -rng.uniform(1, 63, 100000),
                        1 - 2.0 ** -rng.uniform(1, 52, 50000),
                        [0.5, 0.25, 2.0 ** -53, 1 - 2.0 ** -53, 0.5 + 2.0 ** -53, 0.5 - 2.0 ** -54,
                         2.0 ** -63.9]])
    out = np.empty_like(u)
    rc = lib.pbx_ndtri_host(u.ctypes.data_as(C.c_void_p), C.c_int64(len(u)),
                            out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    ref = ndtri(u)
    # |error| <= 3e-15 |x| + 2e-16: the absolute floor is below the resolution of the argument
    # (one grid step 2^-53 of u near 0.5 moves x by 2.8e-16)
    excess = np.abs(out - ref) - (3e-15 * np.abs(ref) + 2e-16)
    assert excess.max() <= 0, excess.max()      # scipy itself is ~1e-15 here and there
    err = np.abs(out - ref) / np.maximum(np.abs(ref), 1e-3)
    assert np.quantile(err, 0.999) <= 1.5e-15
    assert abs(out[len(u) - 7]) <= 1e-16                            # ndtri(0.5)
    bad = np.array([0.0, 1.0, 2.0 ** -70, -0.1, 1.5, np.nan])
    ob = np.empty_like(bad)
    lib.pbx_ndtri_host(bad.ctypes.data_as(C.c_void_p), C.c_int64(len(bad)),
                       ob.ctypes.data_as(C.c_void_p))
    assert np.isnan(ob).all()


def test_blocked_gibbs_closed_form_equals_the_recursion():
    """The algebra behind the tensor-core Gibbs kernel (pbx_gibbs.cu, gibbs_mma_prep_kernel),
    restated in numpy: with C_b the block's 8 rows of the CondCov coefficient matrix split into
    E (columns outside the block) and the strictly lower / upper parts L, U of its diagonal
    block, the eight sequential coordinate updates of cond_cov.py:42-65 equal
    x' = (M [E | U]) x + M z with M = (I - L)^-1 -- the same sweep, up to rounding."""
    from probayes_b200.cond_cov import CondCov
    rng = np.random.default_rng(12)
    for d in (8, 13, 64):
        A = rng.standard_normal((d, d))
        cov = A @ A.T / d + np.eye(d)
        mean = rng.standard_normal(d)
        cc = CondCov(mean, cov, np.tile([-10., 10.], (d, 1)))
        coef = cc.coef_matrix()                                  # zero diagonal
        assert np.abs(np.diag(coef)).max() == 0.0
        x0 = rng.standard_normal(d)
        z = rng.standard_normal(d)            # stands for ndtri(u) * stdv + (mean_i - coef_i . mean)
        # coordinate by coordinate (Gauss-Seidel order)
        xs = x0.copy()
        for i in range(d):
            xs[i] = z[i] + coef[i] @ xs
        # blocks of 8 in closed form, zero padding to a multiple of 8
        dp = (d + 7) // 8 * 8
        Cp = np.zeros((dp, dp)); Cp[:d, :d] = coef
        xb = np.zeros(dp); xb[:d] = x0
        zp = np.zeros(dp); zp[:d] = z
        for b in range(dp // 8):
            sl = slice(8 * b, 8 * b + 8)
            blk = Cp[sl, sl]
            L, U = np.tril(blk, -1), np.triu(blk, 1)
            M = np.linalg.inv(np.eye(8) - L)
            assert np.allclose(np.triu(M, 1), 0.0) and np.allclose(np.diag(M), 1.0)
            EU = Cp[sl].copy()
            EU[:, sl] = U                                        # lower part and diagonal dropped
            xb[sl] = (M @ EU) @ xb + M @ zp[sl]
        assert np.abs(xb[:d] - xs).max() <= 1e-13 * max(1.0, np.abs(xs).max())
        assert not np.any(xb[d:])                                # padding coordinates stay 0
