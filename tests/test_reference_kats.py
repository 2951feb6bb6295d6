"""The known-answer tests the reference's own suite holds at this path's boundary
(SURVEY.md section 8c), restated against the oracle and the host mirror:

  * tests/test_prob.py:18-28   norm: pdf == exp(logpdf), ppf(cdf(x)) == x (allclose)
  * tests/test_variable.py:71-91  a log-ufun RV's sampled grid and its bounded random
    walk (set_delta([1], bound=True)) stay within the limits
"""
import numpy as np
import pytest
import scipy.stats
import scipy.special
from oracle import np_oracle as o
import probayes_b200 as pb


def test_norm_pdf_is_exp_logpdf_and_cdf_inverts():
    """tests/test_prob.py:12-28 with (loc=1, scale=0.5) on linspace(-1, 3, 1000)."""
    values = np.linspace(-1, 3, 1000)
    prob = o.norm_pdf(values, 1., 0.5)
    logp = o.norm_logpdf(values, 1., 0.5)
    assert np.allclose(prob, np.exp(logp))
    assert np.allclose(logp, scipy.stats.norm.logpdf(values, loc=1., scale=0.5), rtol=1e-14,
                       atol=1e-14)
    cump = scipy.special.ndtr((values - 1.) / 0.5)          # what CondCov's limits use
    invc = scipy.special.ndtri(cump) * 0.5 + 1.
    assert np.allclose(values, invc)


# tests/test_variable.py RAN_LOG_TESTS
RAN_LOG_TESTS = [([(1,), (100,)], {100}), ([1, 100], {-100})]


@pytest.mark.parametrize("ran,val", RAN_LOG_TESTS)
def test_log_ufun_grid_and_bounded_walk_stay_within_limits(ran, val):
    """tests/test_variable.py:71-91: grid / random samples of a log-ufun RV, then a bounded
    walk in ufun space -- here the walk is the oracle's ufun_propose with bound (the
    kernel's arithmetic, tests/test_gpu_mh_normreg.py runs the device side)."""
    y = pb.RV('y', vtype=float, vset=ran)
    y.set_ufun((np.log, np.exp))
    np.random.seed(3)
    vals = y.evaluate(val)
    assert np.max(vals) <= y.vlims[1] and np.min(vals) >= y.vlims[0]
    steps = int(abs(list(val)[0]))
    value = np.array([[np.mean(vals)]])
    lims = np.array([y.vlims])
    ex = np.array([[int(e) for e in y.open_ends]])
    rng = np.random.default_rng(5)
    out = np.empty(steps)
    for i in range(steps):
        delta = np.array([[-1. + 2. * rng.random()]])                # [1] -> U(-1, 1)
        value = o.ufun_propose(value, delta, [True], (lims, ex))
        out[i] = value[0, 0]
    assert np.max(out) <= y.vlims[1] and np.min(out) >= y.vlims[0]
    assert np.ptp(out) > 0
