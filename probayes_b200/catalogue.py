"""The device catalogue: maps what the user configured through the reference API
(set_prob / set_tran / set_delta / set_scores ...) onto a libpbx kernel, or refuses.

The reference accepts arbitrary Python callables everywhere (probayes/prob.py:
125-189, expression.py:376-452); a callable cannot run on the GPU, so the drop-in
*recognises* the cases the reference's own examples use and raises
``NotImplementedError`` for the rest -- there is no CPU fallback.

Targets
  mvn      scipy.stats.multivariate_normal(mean, cov)        (mcmc_prob4a, gibbs_norm2d)
  norm     scipy.stats.norm.logpdf + order={'x': 0, 'mu': 'loc', 'sigma': 'scale'}
           with iid=True                                      (metrohast_norm1d, dgei_*)
  normreg  a Python callable that IS norm.logpdf(y, loc=b0 + b1*x, scale=s) -- verified
           numerically by probing it against every role assignment (gibbs_linreg's
           ``norm_reg``), or an explicit ``NormalRegression`` spec
Proposals (field.py:220-317, variable.py:600-638, rf.py:209-220)
  [d] uniform per RV, (d,) spherical, frozen scipy.stats.norm, covariance-matrix tran
"""
import inspect
import itertools
import numpy as np
import scipy.stats

from .pscales import iscomplex, rescale


class NormalRegression:
    """Explicit catalogue spec of y ~ N(intercept + slope*x, scale) (log-density)."""

    def __init__(self, x, y, intercept, slope, scale):
        self.x, self.y, self.intercept, self.slope, self.scale = x, y, intercept, slope, scale


def is_scipy_mvn(obj):
    return obj is scipy.stats.multivariate_normal or \
        type(obj).__name__ in ('multivariate_normal_gen', 'multivariate_normal_frozen')


def mvn_args(args, kwds):
    args = list(args)
    mean = kwds.get('mean', args.pop(0) if args else None)
    cov = kwds.get('cov', args.pop(0) if args else None)
    if mean is None or cov is None:
        raise ValueError("multivariate_normal needs (mean, cov)")
    return np.atleast_1d(np.asarray(mean, dtype=float)), np.atleast_2d(np.asarray(cov, dtype=float))


def _norm_method(prob):
    """'logpdf' / 'pdf' if prob is that bound method of scipy.stats.norm."""
    owner = getattr(prob, '__self__', None)
    if owner is scipy.stats.norm or type(owner).__name__ == 'norm_gen':
        name = getattr(prob, '__name__', None)
        if name in ('logpdf', 'pdf'):
            return name
    return None


def identify_target(holder, stats, paras):
    """holder: the object set_prob was called on (SD/SP).  Returns a spec dict."""
    prob, args, kwds = holder._prob, holder._prob_args, holder._prob_kwds
    log_pscale = iscomplex(holder.pscale)
    if prob is None:
        raise NotImplementedError("no probability set: call set_prob() first")
    if is_scipy_mvn(prob):
        if type(prob).__name__ == 'multivariate_normal_frozen':
            mean, cov = np.atleast_1d(prob.mean), np.atleast_2d(prob.cov)
        else:
            mean, cov = mvn_args(args, kwds)
        names = [rv.name for rv in (stats.varlist if paras is None else
                                    stats.varlist + paras.varlist)]
        if len(mean) != len(names):
            raise ValueError("multivariate_normal over {} variables but {} means".format(
                len(names), len(mean)))
        return dict(kind='mvn', mean=mean, cov=cov, names=names, log_pscale=log_pscale)
    method = _norm_method(prob)
    if method is not None:
        order = holder._order
        if not isinstance(order, dict):
            raise NotImplementedError(
                "scipy.stats.norm.{} needs order={{obs: 0, mean: 'loc', sd: 'scale'}}".format(method))
        roles = {v: k for k, v in order.items()}
        try:
            obs, loc, scale = roles[0], roles['loc'], roles['scale']
        except KeyError:
            raise NotImplementedError("unrecognised order mapping {}".format(order))
        if method == 'pdf' or not log_pscale:
            raise NotImplementedError(
                "the iid normal likelihood is in the catalogue in log pscale only "
                "(set_prob(scipy.stats.norm.logpdf, ..., pscale='log'))")
        if paras is None or obs not in stats.keyset or {loc, scale} - paras.keyset:
            raise NotImplementedError("order must map a stats RV to 0 and parameter RVs "
                                      "to 'loc'/'scale'")
        return dict(kind='normreg', has_slope=False, obs_y=obs, obs_x=None,
                    params=[loc, scale], log_pscale=True)
    if isinstance(prob, NormalRegression):
        return dict(kind='normreg', has_slope=True, obs_y=prob.y, obs_x=prob.x,
                    params=[prob.intercept, prob.slope, prob.scale], log_pscale=True)
    if callable(prob):
        spec = _probe_normreg(prob, stats, paras, log_pscale)
        if spec is not None:
            return spec
    raise NotImplementedError(
        "target {!r} is outside the device catalogue (multivariate_normal, "
        "norm.logpdf with order=..., or a normal linear-regression log-likelihood)".format(prob))


def _probe_normreg(func, stats, paras, log_pscale):
    """Is ``func`` (called by keyword with the RV names, as the reference does) the
    log-density of y ~ N(a + b*x, s)?  Checked on random probes to 1e-12."""
    if paras is None or not log_pscale or stats.nvars != 2 or paras.nvars != 3:
        return None
    try:
        sig = inspect.signature(func)
    except (TypeError, ValueError):
        return None
    names = stats.keylist + paras.keylist
    if not set(names).issubset(sig.parameters.keys()):
        return None
    rng = np.random.default_rng(12345)
    obs = {k: rng.normal(0., 1., 7) for k in stats.keylist}
    pars = {k: float(v) for k, v in zip(paras.keylist, rng.uniform(0.5, 1.5, 3))}
    try:
        got = np.asarray(func(**obs, **pars), dtype=float)
    except Exception:
        return None
    if got.shape != (7,):
        return None
    for (xk, yk) in itertools.permutations(stats.keylist, 2):
        for (a, b, s) in itertools.permutations(paras.keylist, 3):
            want = scipy.stats.norm.logpdf(obs[yk], loc=pars[a] + pars[b] * obs[xk],
                                           scale=pars[s])
            if np.allclose(got, want, rtol=1e-12, atol=1e-12):
                return dict(kind='normreg', has_slope=True, obs_y=yk, obs_x=xk,
                            params=[a, b, s], log_pscale=True)
    return None


# ---------------------------------------------------------------------------
def _frozen_norm_scale(obj):
    if type(obj).__name__ == 'rv_continuous_frozen' and getattr(obj.dist, 'name', '') == 'norm':
        loc = obj.kwds.get('loc', obj.args[0] if len(obj.args) > 0 else 0.)
        scale = obj.kwds.get('scale', obj.args[1] if len(obj.args) > 1 else 1.)
        if float(loc) != 0.:
            raise NotImplementedError("normal proposals must be centred (loc=0)")
        return float(scale)
    return None


def identify_proposal(rf, target_pscale, injected=False):
    """rf: the RF whose set_delta / set_tran describe the proposal.
    Returns dict(kind, scale[D], radius, chol, coef)."""
    D = rf.nvars
    delta, kw = rf._delta, rf._delta_kwds
    lengths = rf.lengths
    out = dict(kind='normal', scale=np.ones(D), radius=0.0, chol=None, coef=1.0,
               bound=bool(kw.get('bound')))
    if isinstance(rf._tfun, np.ndarray):
        out['chol'] = np.asarray(rf._tfun, dtype=float)
    fro = _frozen_norm_scale(delta)
    if fro is not None:
        out.update(kind='normal', scale=np.full(D, fro))
    elif isinstance(delta, list):
        assert len(delta) == 1, "List delta requires a single element"
        d = float(delta[0])
        scale = np.full(D, d)
        if kw.get('scale'):
            assert np.all(np.isfinite(lengths)), "Cannot scale by infinite length"
            scale = d * lengths
        if rf._delta_args:
            unscale = rf._delta_args[0]
            assert isinstance(unscale, dict), \
                "Optional positional arguments must comprises a single dict"
            for i, key in enumerate(rf.keylist):
                if key in unscale:
                    v = unscale[key]
                    scale[i] = float(v[0] if isinstance(v, list) else v)
        out.update(kind='uniform', scale=scale)
    elif isinstance(delta, tuple):
        assert len(delta) == 1, "Tuple delta must contain one element"
        d = float(delta[0])
        assert np.all(np.isfinite(lengths)), "Cannot spherise Variable with infinite length"
        if kw.get('scale'):
            rss = float(np.sqrt(np.sum(lengths ** 2)))
            out.update(kind='spherical', radius=d * rss, scale=lengths.copy())
        else:
            out.update(kind='spherical', radius=d, scale=np.ones(D))
    elif delta is None or callable(delta):
        if not injected:
            raise NotImplementedError(
                "a Python-callable (or missing) delta cannot run on the device: use [d], (d,), "
                "a frozen scipy.stats.norm(0, s), or inject the draws (inj_delta=...)")
    else:
        raise NotImplementedError("unrecognised delta specification {!r}".format(delta))
    # asymmetric (q, r) pair: hastings_scores multiplies the linear proposal density
    # into the target *in the target's pscale* (sp_utils.py:52-64, rf.py:531-536)
    tran = rf._tran
    if isinstance(tran, tuple):
        assert len(tran) == 2 and all(callable(t) for t in tran), \
            "tuple transitions must be a (q, r) pair of callables"
        vals = _probe_constant(tran[0], rf)
        if vals is None:
            raise NotImplementedError("state-dependent asymmetric proposal densities are not "
                                      "in the device catalogue")
        out['coef'] = float(rescale(vals, rf.pscale, 1.)) if iscomplex(target_pscale) else 1.0
    return out


def _probe_constant(func, rf):
    rng = np.random.default_rng(99)
    seen = []
    for _ in range(3):
        kwds = {}
        for rv in rf.varlist:
            lo, hi = rv.vlims
            lo = lo if np.isfinite(lo) else -1.
            hi = hi if np.isfinite(hi) else 1.
            a, b = rng.uniform(lo, hi, 2)
            kwds[rv.name] = float(a)
            kwds[rv.name + "'"] = float(b)
        try:
            seen.append(float(func(**kwds)))
        except Exception:
            return None
    return seen[0] if max(seen) == min(seen) else None
