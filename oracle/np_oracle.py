"""numpy restatement of the probayes hot path -- TEST INFRASTRUCTURE ONLY.

See oracle/__init__.py for who may import this.  Every function cites the
reference file:line it follows (relative to the reference checkout).  Third
party arithmetic the reference delegates to and that is NOT under the reference
tree is restated from its published formula:

  * scipy (reference pins ``scipy>=1.16.3`` in setup.cfg:15; 1.18.1 installed)
      - ``stats.norm.logpdf``   : -z*z/2 - log(sqrt(2 pi)) - log(scale),
                                  z = (x - loc)/scale
      - ``stats.multivariate_normal.logpdf`` : eigh whitening,
                                  -0.5*(d*log(2 pi) + log_pdet + |dev @ U|^2),
                                  U = u * sqrt(1/s);  pdf = exp(logpdf)
      - ``stats.norm.cdf/ppf``  : ndtr / ndtri (taken from scipy.special here;
                                  the oracle is allowed the same leaf library)
  * numpy (``>=2.3.3``): sum / exp / log / linalg.inv / linalg.eigh

All arithmetic is fp64.  Functions are vectorised over a leading *chain* axis;
the reference itself is single-chain, so chain c of a batched run is by
definition the reference run on chain c's own (init, proposal, threshold)
streams.
"""
import numpy as np

# probayes/constants.py:9-32
NEARLY_POSITIVE_ZERO = 2.2250738585072014e-308
NEARLY_POSITIVE_INF = 1.7976931348623158e+308
NEARLY_NEGATIVE_INF = -NEARLY_POSITIVE_INF
LOG_NEARLY_POSITIVE_INF = float(np.log(NEARLY_POSITIVE_INF))
LOG_2PI = float(np.log(2.0 * np.pi))
LOG_SQRT_2PI = float(np.log(np.sqrt(2.0 * np.pi)))


# ----------------------------------------------------------------------------
# pscales (probayes/pscales.py)
# ----------------------------------------------------------------------------
def log_prob(prob):
    """probayes/pscales.py:44-53 -- clamped log."""
    prob = np.asarray(prob, dtype=float)
    out = np.full(prob.shape, NEARLY_NEGATIVE_INF)
    ok = prob >= NEARLY_POSITIVE_ZERO
    out[ok] = np.log(prob[ok])
    return out


def exp_logp(logp):
    """probayes/pscales.py:56-65 -- clamped exp (note: NaN maps to +huge)."""
    logp = np.asarray(logp, dtype=float)
    out = np.full(logp.shape, NEARLY_POSITIVE_INF)
    ok = logp <= LOG_NEARLY_POSITIVE_INF
    out[ok] = np.exp(logp[ok])
    return out


def to_linear(prob, log_pscale):
    """rescale(prob, pscale, 1.) for pscale in {1., 0j}: pscales.py:100-131."""
    return exp_logp(prob) if log_pscale else np.asarray(prob, dtype=float)


def from_linear(prob, log_pscale):
    """rescale(prob, 1., pscale) for pscale in {1., 0j}: pscales.py:100-131."""
    return log_prob(prob) if log_pscale else np.asarray(prob, dtype=float)


def div_prob_linear(num, den):
    """probayes/pscales.py:219-236 with both operands already linear."""
    return np.asarray(num, dtype=float) / np.maximum(NEARLY_POSITIVE_ZERO, den)


# ----------------------------------------------------------------------------
# leaf densities (scipy, restated)
# ----------------------------------------------------------------------------
def norm_logpdf(x, loc, scale):
    """scipy.stats.norm.logpdf: _continuous_distns.py (norm_gen._logpdf) through
    rv_continuous.logpdf: -(z*z)/2 - log(sqrt(2 pi)) - log(scale)."""
    z = (x - loc) / scale
    return -(z * z) / 2.0 - LOG_SQRT_2PI - np.log(scale)


def norm_pdf(x, loc, scale):
    """scipy.stats.norm.pdf: exp(-z*z/2)/sqrt(2 pi)/scale."""
    z = (x - loc) / scale
    return np.exp(-(z * z) / 2.0) / np.sqrt(2.0 * np.pi) / scale


def mvn_whiten(cov):
    """scipy.stats._multivariate._PSD: s,u = eigh(cov); U = u*sqrt(1/s);
    log_pdet = sum(log s).  Returns (U, log_pdet)."""
    cov = np.asarray(cov, dtype=float)
    s, u = np.linalg.eigh(cov)
    return u * np.sqrt(1.0 / s), float(np.sum(np.log(s)))


def mvn_value_order(d):
    """probayes/prob.py:349-358 -- the value list is reversed (and rotated for
    d > 2) before np.meshgrid, so component j of the point scipy sees is
    variable order[j]: d=2 -> [1, 0]; d>2 -> [d-2, ..., 0, d-1]."""
    if d == 1:
        return [0]
    order = list(range(d))[::-1]
    if d > 2:
        order = order[1:] + [order[0]]
    return order


def mvn_logpdf(x, mean, U, log_pdet):
    """scipy multivariate_normal_gen._logpdf: x[..., d]."""
    dev = x - mean
    maha = np.sum(np.square(dev @ U), axis=-1)
    return -0.5 * (mean.shape[-1] * LOG_2PI + log_pdet + maha)


# ----------------------------------------------------------------------------
# vtypes / priors
# ----------------------------------------------------------------------------
def uniform_grid(lo, hi, n, ex_lo=False, ex_hi=False):
    """probayes/vtypes.py:169-204 for n > 0 (deterministic grid)."""
    if not ex_lo and not ex_hi:
        if n == 1:
            return np.linspace(lo, hi, 3)[1:-1]
        return np.linspace(lo, hi, n)
    if ex_lo and ex_hi:
        return np.linspace(lo, hi, n + 2)[1:-1]
    if ex_lo:
        return np.linspace(lo, hi, n + 1)[1:]
    return np.linspace(lo, hi, n + 1)[:-1]


def box_inside(x, lo, hi, ex_lo, ex_hi):
    """probayes/variable.py:352-366 -- open/closed ends of a float vset."""
    a = (x > lo) if ex_lo else (x >= lo)
    b = (x < hi) if ex_hi else (x <= hi)
    return np.logical_and(a, b)


def box_log_prior(x, lo, hi, ex_lo, ex_hi, log_ufun=False):
    """Default prior of a float RV in log pscale: -log(length) inside, where
    the length is measured in ufun-space (no Jacobian for tuple ufuns), and
    NEARLY_NEGATIVE_INF outside.  probayes/rv.py:153-166,305-320;
    rv_utils.py:8-47; variable.py:337-343."""
    if log_ufun:
        length = np.log(hi) - np.log(lo)
    else:
        length = hi - lo
    nlhv = -np.log(length)
    return np.where(box_inside(x, lo, hi, ex_lo, ex_hi), nlhv,
                    NEARLY_NEGATIVE_INF)


# ----------------------------------------------------------------------------
# MH accept rule
# ----------------------------------------------------------------------------
def mh_score_reference(p_succ, p_pred, log_pscale, coef=1.0):
    """metropolis_scores / hastings_scores: probayes/sp_utils.py:19-64 with
    div_prob (pscales.py:219-236).  ``coef`` is the linear proposal density
    that hastings_scores multiplies into the target *in the target's pscale*
    (sp_utils.py:62-64) -- 1.0 for metropolis / symmetric."""
    num = to_linear(np.asarray(p_succ) * coef, log_pscale)
    den = to_linear(np.asarray(p_pred) * coef, log_pscale)
    return np.minimum(1.0, div_prob_linear(num, den))


def mh_walk(init, delta, thresh, target, propose, log_pscale,
            accept="reference", coef=1.0):
    """The MH step loop: probayes/sp.py:221-258, sp_utils.py:19-37,
    sd.py:253-288.

    init   [C, D]      state before step 1
    delta  [T, C, D]   injected proposal draws (what the reference's delta
                       callable / np.random.uniform would have returned)
    thresh [T, C]      injected thresholds (np.random.uniform() per step, drawn
                       on every step including the first: sp.py:248-249)
    target(x[C, D]) -> prob[C] in the target's pscale
    propose(x[C, D], delta[C, D]) -> x'[C, D]
    accept 'reference': s = min(1, lin(p')/max(tiny, lin(p))), accept iff s >= t
           'log'      : accept iff (p' - p) >= log(t)  (log pscale only; the
                        form the device path uses where the reference's linear
                        ratio underflows -- SURVEY section 0.3)
    Step 1 always accepts (s = None: sp_utils.py:24-25, 35).

    Returns dict: x[T,C,D], prob[T,C] (v after the decision), s[T,C] (nan on
    step 1), u[T,C] bool, xprop[T,C,D], pprop[T,C].
    """
    init = np.array(init, dtype=float)
    T, C, D = delta.shape
    x = init.copy()
    p = np.zeros(C)
    X = np.empty((T, C, D)); P = np.empty((T, C)); S = np.full((T, C), np.nan)
    Uacc = np.zeros((T, C), dtype=bool)
    XP = np.empty((T, C, D)); PP = np.empty((T, C))
    for k in range(T):
        xp = propose(x, delta[k])
        pp = target(xp)
        if k == 0:
            acc = np.ones(C, dtype=bool)
        elif accept == "reference":
            s = mh_score_reference(pp, p, log_pscale, coef)
            S[k] = s
            acc = s >= thresh[k]
        elif accept == "log":
            assert log_pscale
            d = coef * (pp - p)
            S[k] = np.minimum(1.0, np.exp(np.minimum(d, 0.0)))
            acc = d >= np.log(thresh[k])
        else:
            raise ValueError(accept)
        x = np.where(acc[:, None], xp, x)
        p = np.where(acc, pp, p)
        X[k], P[k], Uacc[k], XP[k], PP[k] = x, p, acc, xp, pp
    return dict(x=X, prob=P, s=S, u=Uacc, xprop=XP, pprop=PP)


# ----------------------------------------------------------------------------
# config C1/C2: 2-D (d-D) correlated-normal target, additive proposal
# ----------------------------------------------------------------------------
def mh_mvn_walk(init, delta, thresh, mean, cov, tran_chol=None,
                log_pscale=False, reorder=True, accept="reference", bound=None):
    """MH on scipy.stats.multivariate_normal(mean, cov) as set up in
    examples/mcmc/mcmc_prob4a.py:38-49.

    proposal: x' = x + delta, or x' = x + chol @ delta when the transition was
    given as a covariance matrix (probayes/rf.py:209-220,346-348).
    target:   pdf (linear pscale) or logpdf (log pscale) -- prob.py:347-348 --
              evaluated on the permuted point of prob.py:349-358 when
              ``reorder``.
    """
    mean = np.asarray(mean, dtype=float)
    U, log_pdet = mvn_whiten(cov)
    d = mean.shape[0]
    order = mvn_value_order(d) if reorder else list(range(d))

    def target(x):
        lp = mvn_logpdf(x[:, order], mean, U, log_pdet)
        return lp if log_pscale else np.exp(lp)

    def propose(x, dl):
        if tran_chol is not None:
            dl = dl @ np.asarray(tran_chol, dtype=float).T
        if bound is not None:            # set_delta(..., bound=True): variable.py:700-727
            return ufun_propose(x, dl, [False] * d, bound)
        return x + dl

    return mh_walk(init, delta, thresh, target, propose, log_pscale, accept)


# ----------------------------------------------------------------------------
# config C3 (+ metrohast_norm1d): iid normal likelihood with affine mean
# ----------------------------------------------------------------------------
def normreg_logjoint(theta, x_obs, y_obs, lims, ex, log_ufun, has_slope=True):
    """log p(theta, data) in log pscale for
        y_i ~ N(b0 + b1*x_i, sigma)        theta = (b0, b1, sigma)   has_slope
        y_i ~ N(mu, sigma)                 theta = (mu, sigma)       otherwise
    = sum_i norm.logpdf (iid product in log pscale: probayes/rf.py:541-562,
    pd.py:332-370 -> np.sum over the observation axis) + sum of independent
    box priors (joint=True: sd.py:154-161, rf_utils.py:10-22, pscales.py:160-216).

    theta [C, P]; lims [P, 2]; ex [P, 2] bool (open ends); log_ufun [P] bool.
    """
    theta = np.asarray(theta, dtype=float)
    C, P = theta.shape
    out = np.empty(C)
    for c in range(C):
        if has_slope:
            loc = theta[c, 0] + theta[c, 1] * x_obs
            sig = theta[c, 2]
        else:
            loc = theta[c, 0]
            sig = theta[c, 1]
        with np.errstate(invalid="ignore", divide="ignore"):
            out[c] = np.sum(norm_logpdf(y_obs, loc, sig))
    prior = np.zeros(C)
    for j in range(P):
        prior = prior + box_log_prior(theta[:, j], lims[j][0], lims[j][1],
                                      ex[j][0], ex[j][1], log_ufun[j])
    return prior + out


def ufun_propose(theta, delta, log_ufun, bound=None):
    """probayes/variable.py:693-697: x' = x + d, or ufun^-1(ufun(x) + d) with
    the (np.log, np.exp) tuple ufun.  bound = (lims [P, 2], ex [P, 2]) applies
    set_delta(..., bound=True) (variable.py:700-727, the scalar branch the sampler
    takes): closed ends clip, a proposal beyond an open end bounces back."""
    out = theta + delta
    for j, lg in enumerate(log_ufun):
        if lg:
            out[:, j] = np.exp(np.log(theta[:, j]) + delta[:, j])
    if bound is not None:
        lims, ex = bound
        for j in range(theta.shape[1]):
            lo, hi = lims[j]
            olo, ohi = bool(ex[j][0]), bool(ex[j][1])
            v, old = out[:, j], theta[:, j]
            if not olo and not ohi:
                out[:, j] = np.maximum(lo, np.minimum(hi, v))
            elif olo and ohi:
                out[:, j] = np.where((v > lo) & (v < hi), v, old)
            elif olo:
                out[:, j] = np.where(v < lo, old, np.minimum(hi, v))
            else:
                out[:, j] = np.where(v > hi, old, np.maximum(lo, v))
    return out


def mh_normreg_walk(init, delta, thresh, x_obs, y_obs, lims, ex, log_ufun,
                    has_slope=True, accept="reference", coef=1.0, bound=False):
    """MH for the (mu, sigma) / (beta_0, beta_1, y_sigma) posterior in log
    pscale with iid=True, joint=True: examples/mcmc/metrohast_norm1d.py:23-42,
    examples/mcmc/gibbs_linreg.py:28-36 + SURVEY appendix B.5."""
    def target(th):
        return normreg_logjoint(th, x_obs, y_obs, lims, ex, log_ufun, has_slope)

    def propose(th, dl):
        return ufun_propose(th, dl, log_ufun, (lims, ex) if bound else None)

    return mh_walk(init, delta, thresh, target, propose, True, accept, coef)


# ----------------------------------------------------------------------------
# config C4: discrete grid exact inference (normal mean/std)
# ----------------------------------------------------------------------------
def grid_norm_logjoint(x_obs, mu, sigma, logprior_mu, logprior_sigma,
                       chunk=64):
    """log-joint [M, S] = sum_i norm.logpdf(x_i; mu_m, sigma_s) + priors.
    The reference materialises [N, M, S] and np.sum(axis=0)s it
    (probayes/rf.py:565-581, pd.py:368) then adds the priors with prod_rule
    (pd_utils.py:85-328, pscales.py:160-216: prior first, likelihood second).
    Here the mu axis is chunked to bound memory; the per-cell summation order
    over observations (sequential, axis 0) is the reference's."""
    x_obs = np.asarray(x_obs, dtype=float)
    M, S = len(mu), len(sigma)
    out = np.empty((M, S))
    xs = x_obs[:, None, None]
    sg = np.asarray(sigma, dtype=float)[None, None, :]
    for m0 in range(0, M, chunk):
        mm = np.asarray(mu[m0:m0 + chunk], dtype=float)[None, :, None]
        out[m0:m0 + chunk] = np.sum(norm_logpdf(xs, mm, sg), axis=0)
    prior = np.asarray(logprior_mu, dtype=float).reshape(-1, 1) + \
        np.asarray(logprior_sigma, dtype=float).reshape(1, -1)
    return prior + out


def grid_conditionalise(logjoint):
    """PD.conditionalise over the scalar-ised iid key in log pscale:
    probayes/pd.py:285-295 -> prob - max; exp; div_prob(prob, sum); log_prob."""
    p = logjoint - np.max(logjoint)
    p = exp_logp(p)
    p = div_prob_linear(p, np.sum(p))
    return log_prob(p)


def grid_marginal(logpost, axis):
    """PD.marginalise in log pscale: probayes/pd.py:162-164 -> exp (no
    max-shift); sum over the marginalised axes; clamped log."""
    p = exp_logp(logpost)
    return log_prob(np.sum(p, axis=axis))


def grid_expectation(logpost, vals, axis_of_val):
    """PD.expectation: probayes/pd.py:373-405 (all marginal keys): E[val] =
    sum(prob*val) / max(tiny, sum(prob)) over every axis."""
    p = exp_logp(logpost)
    shape = [1] * p.ndim
    shape[axis_of_val] = -1
    v = np.asarray(vals, dtype=float).reshape(shape)
    return float(div_prob_linear(np.sum(p * v), np.sum(p)))


# ----------------------------------------------------------------------------
# config C5: CondCov Gibbs
# ----------------------------------------------------------------------------
class CondCovOracle:
    """probayes/cond_cov.py:22-39: per-coordinate Schur regression
    coefficients, conditional stdv and fixed cdf limits."""

    def __init__(self, mean, cov, lims):
        from scipy.special import ndtr
        self.mean = np.atleast_1d(np.asarray(mean, dtype=float))
        self.cov = np.atleast_2d(np.asarray(cov, dtype=float))
        n = self.n = len(self.mean)
        rl = np.atleast_2d(np.asarray(lims, dtype=float)) - self.mean[:, None]
        self.stdv = np.empty(n)
        self.coef = np.zeros((n, n))     # row i: coef over j != i, 0 at j == i
        for i in range(n):
            idx = [j for j in range(n) if j != i]
            sub = self.cov[np.ix_(idx, idx)]
            ru = self.cov[i, idx]
            co = ru @ np.linalg.inv(sub)
            self.coef[i, idx] = co
            self.stdv[i] = np.sqrt(self.cov[i, i] - float(co @ self.cov[idx, i]))
        # norm.cdf(lim, loc=0, scale=stdv) = ndtr(lim/stdv)
        self.cdfs = ndtr(rl / self.stdv[:, None])

    def step(self, x, i, r):
        """One coordinate update for a batch x[C, n] with uniforms r[C] in
        (0,1): probayes/cond_cov.py:42-65 via rf_utils.py:50-65.
        u = lo + (hi-lo)*r  (np.random.uniform(lo, hi): vtypes.py:186);
        x_i = ndtri(u)*stdv_i + mean_i + coef_i.(x_-i - mean_-i)."""
        from scipy.special import ndtri
        lo, hi = self.cdfs[i]
        u = lo + (hi - lo) * r
        dmu = x - self.mean
        cm = self.mean[i] + dmu @ self.coef[i]      # coef[i, i] == 0
        out = x.copy()
        out[:, i] = ndtri(u) * self.stdv[i] + cm
        return out


def gibbs_mvn_walk(init, runif, mean, cov, lims, log_pscale=False,
                   reorder=True, start=0):
    """Gibbs through SP.next with tsteps=1: probayes/rf.py:446-462 cycles one
    coordinate per step; scores/thresh are nan and every step is kept
    (sp_utils.py:75-84); the mvn target is still evaluated and recorded every
    step (sd.py:286) on the permuted point (prob.py:349-358).

    init [C, d]; runif [T, C].  Returns x[T, C, d], prob[T, C]."""
    cc = CondCovOracle(mean, cov, lims)
    mean = np.asarray(mean, dtype=float)
    U, log_pdet = mvn_whiten(cov)
    d = cc.n
    order = mvn_value_order(d) if reorder else list(range(d))
    x = np.array(init, dtype=float)
    T, C = runif.shape
    X = np.empty((T, C, d)); P = np.empty((T, C))
    for k in range(T):
        x = cc.step(x, (start + k) % d, runif[k])
        lp = mvn_logpdf(x[:, order], mean, U, log_pdet)
        X[k] = x
        P[k] = lp if log_pscale else np.exp(lp)
    return dict(x=X, prob=P)


# ----------------------------------------------------------------------------
# chain summaries (new in the device path; plain definitions for the checker)
# ----------------------------------------------------------------------------
def chain_moments(X):
    """X[T, C, D] -> per-chain mean[C, D], unbiased var[C, D]."""
    return X.mean(axis=0), X.var(axis=0, ddof=1)


def rhat(X):
    """Gelman-Rubin potential scale reduction from X[T, C, D] -> [D]."""
    T = X.shape[0]
    m, v = chain_moments(X)
    W = v.mean(axis=0)
    B = T * m.var(axis=0, ddof=1)
    return np.sqrt(((T - 1) / T * W + B / T) / W)


# ----------------------------------------------------------------------------
# PD post-processing and OMC random sampling (rank 3)
# ----------------------------------------------------------------------------
def box_sample(lims, log_ufun, runif):
    """Variable.evaluate({0}): probayes/variable.py:558-583 with vtypes.uniform
    (vtypes.py:186): one uniform in ufun space per variable, mapped back through
    ufun^-1.  runif [T, P] in draw order (variables in field order per step)."""
    lims = np.asarray(lims, dtype=float)
    runif = np.asarray(runif, dtype=float)
    out = np.empty(runif.shape[::-1])
    for j in range(lims.shape[0]):
        lo, hi = lims[j]
        if log_ufun[j]:
            ulo, uhi = np.log(lo), np.log(hi)
            out[j] = np.exp(ulo + (uhi - ulo) * runif[:, j])
        else:
            out[j] = lo + (hi - lo) * runif[:, j]
    return out


def pd_sorted_order(keys):
    """PD.sorted: probayes/pd.py:473-474 (np.argsort; the stable variant so that
    ties have a defined order)."""
    return np.argsort(np.ravel(keys), kind='stable')


def pd_cumprob(prob, log_pscale):
    """PD.quantile: probayes/pd.py:426-429 -- rescale to linear, cumsum,
    div_prob by the last element."""
    rav = to_linear(np.ravel(np.asarray(prob, dtype=float)), log_pscale)
    cum = np.cumsum(rav)
    return div_prob_linear(cum, cum[-1]), rav


def pd_quantile_index(cum, q):
    """probayes/pd.py:430."""
    return np.maximum(0, np.digitize(np.atleast_1d(np.asarray(q, float)), cum) - 1)


def pd_quantile_1d(vals, prob, log_pscale, q):
    """PD.quantile for a 1-D distribution whose ``vals`` are monotonic
    (probayes/pd.py:433-457): the bracketing cell from the cumulative
    probability, then interpolation in cumulative probability when the two
    cell probabilities are close, else their probability-weighted mean."""
    vals = np.ravel(np.asarray(vals, dtype=float))
    cum, rav = pd_cumprob(prob, log_pscale)
    out = []
    for qq, i in zip(np.atleast_1d(q), pd_quantile_index(cum, q)):
        i = int(min(i, len(vals) - 1))
        if i == len(vals) - 1:
            out.append(vals[i])
        elif abs(rav[i + 1] - rav[i]) < min(qq, 1. - qq):
            out.append(float(np.interp(qq, cum[i:i + 2], vals[i:i + 2])))
        else:
            w = rav[i:i + 2]
            out.append(float(np.sum(w * vals[i:i + 2]) / np.sum(w)))
    return np.array(out)


def pd_expectation(prob, log_pscale, row_vals=None, col_vals=None):
    """PD.expectation over every array axis: probayes/pd.py:387-402 -- returns
    (total, [sum p*row_val], [sum p*col_val]) for a [rows, cols] prob."""
    p = to_linear(np.atleast_2d(np.asarray(prob, dtype=float)), log_pscale)
    rv = [] if row_vals is None else [float(np.sum(p * np.asarray(v, float)[:, None]))
                                      for v in row_vals]
    cv = [] if col_vals is None else [float(np.sum(p * np.asarray(v, float)[None, :]))
                                      for v in col_vals]
    return float(np.sum(p)), rv, cv


# ----------------------------------------------------------------------------
# joint-product / division algebra
# ----------------------------------------------------------------------------
def pd_product(a, a_log, b, b_log):
    """pscales.prod_rule for two factors (probayes/pscales.py:160-216): sum in log
    space as soon as one factor is in log pscale (the linear one through the
    clamped log), plain product otherwise.  Returns (prob, is_log)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if a_log or b_log:
        return (a if a_log else log_prob(a)) + (b if b_log else log_prob(b)), True
    return a * b, False


def pd_divide(a, a_log, b, b_log):
    """pscales.div_prob (probayes/pscales.py:219-236) as PD.__truediv__ calls it
    (pd.py:614): both to linear, num / max(tiny, den), back to the dividend's
    pscale."""
    with np.errstate(over="ignore"):
        q = div_prob_linear(to_linear(np.asarray(a, dtype=float), a_log),
                            to_linear(np.asarray(b, dtype=float), b_log))
    return from_linear(q, a_log)
