"""Probability scales: a pscale is a positive float (linear coefficient) or a
complex number (log offset; ``'log'`` is ``0j``).  Host-side mirror of the
reference's probayes/pscales.py semantics for scalars and small arrays (priors,
proposal densities, scores); the big-array conversions of the hot path run in
libpbx (``Engine.log_prob_`` / ``Engine.exp_logp_`` and the fused kernels).

Function names, argument meaning and clamping behaviour follow the reference:
eval_pscale (pscales.py:21-41), log_prob (44-53), exp_logp (56-65), logp_offs /
prob_coef (68-85), rescale (100-131), prod_pscale (134-157), prod_rule (160-216),
div_prob (219-236)."""
import numpy as np
from .constants import (NEARLY_POSITIVE_ZERO, NEARLY_POSITIVE_INF, NEARLY_NEGATIVE_INF,
                        LOG_NEARLY_POSITIVE_INF, COMPLEX_ZERO)


def iscomplex(pscale):
    return isinstance(pscale, complex)


def eval_pscale(pscale=None):
    """None/1 -> 1.0; 'log'/'ln'/0 -> 0j; other reals/complex pass through."""
    if pscale is None:
        return 1.
    if iscomplex(pscale):
        return pscale
    if isinstance(pscale, str):
        if pscale in ('log', 'ln'):
            return COMPLEX_ZERO
        raise ValueError("Cannot evaluate pscale={}".format(pscale))
    if isinstance(pscale, (int, float, np.integer, np.floating)):
        if pscale == 0:
            return COMPLEX_ZERO
        return float(pscale)
    raise ValueError("Cannot evaluate pscale={}".format(pscale))


def log_prob(prob):
    """log with values below the smallest normal mapped to -1.797e308."""
    if np.isscalar(prob):
        return float(np.log(prob)) if prob >= NEARLY_POSITIVE_ZERO else NEARLY_NEGATIVE_INF
    prob = np.asarray(prob, dtype=float)
    out = np.full(prob.shape, NEARLY_NEGATIVE_INF)
    ok = prob >= NEARLY_POSITIVE_ZERO
    out[ok] = np.log(prob[ok])
    return out


def exp_logp(logp):
    """exp with arguments above log(1.797e308) (and NaN) mapped to 1.797e308."""
    if np.isscalar(logp):
        return float(np.exp(logp)) if logp <= LOG_NEARLY_POSITIVE_INF else NEARLY_POSITIVE_INF
    logp = np.asarray(logp, dtype=float)
    out = np.full(logp.shape, NEARLY_POSITIVE_INF)
    ok = logp <= LOG_NEARLY_POSITIVE_INF
    out[ok] = np.exp(logp[ok])
    return out


def logp_offs(pscale=None):
    pscale = eval_pscale(pscale)
    if not iscomplex(pscale):
        return float(np.log(pscale))
    if abs(pscale.imag) < NEARLY_POSITIVE_ZERO:
        return float(pscale.real)
    return -float(pscale.real)


def prob_coef(pscale=None):
    pscale = eval_pscale(pscale)
    if not iscomplex(pscale):
        return float(pscale)
    return float(np.exp(logp_offs(pscale)))


def rescale(prob, *args):
    """rescale(prob, to) or rescale(prob, from, to)."""
    if not np.isscalar(prob):
        prob = np.asarray(prob, dtype=float)
    if not args:
        return prob
    src, dst = (None, args[0]) if len(args) == 1 else (args[0], args[1])
    src, dst = eval_pscale(src), eval_pscale(dst)
    s_log, d_log = iscomplex(src), iscomplex(dst)
    if s_log == d_log and src == dst:
        return prob
    if not s_log and not d_log:
        coef = src / dst
        return prob if coef == 1. else coef * prob
    if not s_log:
        prob = log_prob(prob)
    shift = logp_offs(src) - logp_offs(dst)
    if abs(shift) >= NEARLY_POSITIVE_ZERO:
        prob = prob + shift
    return prob if d_log else exp_logp(prob)


def prod_pscale(pscales, use_logp=None):
    if not len(pscales):
        return None
    if use_logp is None:
        use_logp = any(iscomplex(p) for p in pscales)
    acc = 0. if use_logp else 1.
    for p in pscales:
        p = eval_pscale(p)
        if use_logp:
            acc += logp_offs(p)
        else:
            acc *= prob_coef(p)
    if not use_logp:
        return acc
    if abs(acc) < NEARLY_POSITIVE_ZERO:
        return COMPLEX_ZERO
    if acc > 0:
        return complex(np.log(acc), 0.)
    return complex(np.log(-acc), np.pi)


def prod_rule(*args, **kwds):
    """Product of probability arrays given per-argument pscales -> (prob, pscale):
    sums in log space if any pscale is complex, products otherwise."""
    pscales = kwds.get('pscales', [1.] * len(args))
    assert len(pscales) == len(args), \
        "Input pscales length {} incommensurate with number of arguments {}".format(
            len(pscales), len(args))
    use_logp = kwds.get('use_logp', any(iscomplex(p) for p in pscales))
    natural = prod_pscale(pscales, use_logp)
    pscale = kwds.get('pscale', natural)
    terms = []
    for arg, ps in zip(args, pscales):
        if use_logp and not iscomplex(ps):
            arg = log_prob(arg)
        elif not use_logp and iscomplex(ps):
            arg = exp_logp(arg)
        terms.append(arg)
    prob = np.copy(terms[0]) if len(terms) == 1 else None
    if prob is None:
        prob = terms[0] + terms[1] if use_logp else terms[0] * terms[1]
        for t in terms[2:]:
            prob = prob + t if use_logp else prob * t
    if use_logp != iscomplex(pscale):
        prob = rescale(prob, natural, pscale)
    return prob, pscale


def div_prob(dividend, divisor, *args, pscale=None):
    """Safe division in linear space: num / max(tiny, den), then rescaled."""
    pscales = [None, None]
    if len(args):
        assert len(args) == 2, "Both pscales must be specified if at all"
        pscales = [eval_pscale(args[0]), eval_pscale(args[1])]
        pscale = pscale or pscales[0]
    num = rescale(dividend, pscales[0], None)
    den = rescale(divisor, pscales[1], None)
    with np.errstate(over='ignore'):
        quotient = num / np.maximum(NEARLY_POSITIVE_ZERO, den)
    return rescale(quotient, None, pscale)
