"""RV -- a random variable: name, value type, value set (domain), optional
invertible domain transform (``ufun``) and the default box-uniform prior.

Mirror of the reference interface for the float variables the hot path uses
(probayes/rv.py:57-76,153-166,305-320; variable.py:169-219,249-368,466-497,
558-587; rv_utils.py:8-47).  Non-float vtypes, symbolic (sympy) ufuns and
per-variable Markov transitions are outside the device catalogue and raise
``NotImplementedError``.
"""
import collections
import numpy as np

from .pscales import eval_pscale, iscomplex, rescale, log_prob
from .constants import NEARLY_NEGATIVE_INF
from .vtypes import isscalar, isunitsetint, uniform

_LOG_UFUNS = (np.log,)
_EXP_UFUNS = (np.exp,)


class RV:
    """RV(name, vtype=float, vset=None, pscale=None)

    vset for floats: ``(lo, hi)`` tuple -> both ends open; ``[lo, hi]`` -> closed;
    ``[(lo,), hi]`` / ``[lo, (hi,)]`` -> mixed; a two-element ``set`` -> closed.
    """

    def __init__(self, name, vset=None, vtype=None, prob=None, *args, pscale=None, **kwds):
        # positional order of the reference (probayes/rv.py:57-62): name, vset, vtype, prob
        assert isinstance(name, str) and name.isidentifier(), \
            "Variable name must be a valid identifier: {}".format(name)
        if prob is not None or args or kwds:
            raise NotImplementedError("per-variable probability expressions are outside the "
                                      "device catalogue (box-uniform priors are)")
        if vtype not in (float, np.float64, None):
            raise NotImplementedError(
                "only float random variables are in the device catalogue (got {})".format(vtype))
        self._name = name
        self._vtype = float
        self._ufun = None
        self._log_ufun = False
        self._pscale = eval_pscale(pscale)
        self._delta = None
        self._set_vset(vset)

    # ---- domain --------------------------------------------------------------------
    def _set_vset(self, vset):
        if vset is None:
            vset = (-np.inf, np.inf)
        if isinstance(vset, (set, frozenset)):
            vset = sorted(vset)
        if isinstance(vset, tuple):
            assert len(vset) == 2, "Tuple vsets contain pairs of values, not {}".format(vset)
            lo, hi = sorted(float(v) for v in vset)
            vset = [(lo,), (hi,)]
        assert isinstance(vset, list) and len(vset) == 2, \
            "Floating point vset must be two elements, not {}".format(vset)
        ends, opens = [], []
        for v in vset:
            is_open = isinstance(v, tuple)
            ends.append(float(v[0] if is_open else v))
            opens.append(is_open)
        if ends[1] < ends[0]:
            ends, opens = ends[::-1], opens[::-1]
        self._vset = [(e,) if o else e for e, o in zip(ends, opens)]
        self._vlims = np.array(ends)
        self._open = tuple(opens)
        self._eval_ulims()

    def _eval_ulims(self):
        self._ulims = np.log(self._vlims) if self._log_ufun else self._vlims
        self._isfinite = bool(np.all(np.isfinite(self._ulims)))
        self._length = float(max(self._ulims) - min(self._ulims))
        self._lhv = log_prob(self._length) if self._isfinite else np.inf
        self._nlhv = -self._lhv

    @property
    def name(self):
        return self._name

    @property
    def vtype(self):
        return self._vtype

    @property
    def vset(self):
        return self._vset

    @property
    def vlims(self):
        return self._vlims

    @property
    def ulims(self):
        return self._ulims

    @property
    def open_ends(self):
        """(lower_open, upper_open)"""
        return self._open

    @property
    def length(self):
        return self._length

    @property
    def isfinite(self):
        return self._isfinite

    @property
    def pscale(self):
        return self._pscale

    @property
    def ufun(self):
        return self._ufun

    @property
    def log_ufun(self):
        return self._log_ufun

    def inside(self, x):
        lo = (x > self._vlims[0]) if self._open[0] else (x >= self._vlims[0])
        hi = (x < self._vlims[1]) if self._open[1] else (x <= self._vlims[1])
        return np.logical_and(lo, hi)

    def __contains__(self, x):
        return bool(np.all(self.inside(x)))

    def set_ufun(self, ufun=None, *args, **kwds):
        """Monotonic invertible domain transform as a ``(forward, inverse)`` tuple.
        The device catalogue knows ``(np.log, np.exp)``; anything else (including
        sympy expressions) is refused."""
        if ufun is None:
            self._ufun, self._log_ufun = None, False
            self._eval_ulims()
            return
        message = "Non-iconic input ufun be a two-sized tuple of callable functions"
        assert isinstance(ufun, tuple) and len(ufun) == 2 and callable(ufun[0]) \
            and callable(ufun[1]), message
        if args or kwds or ufun[0] not in _LOG_UFUNS or ufun[1] not in _EXP_UFUNS:
            raise NotImplementedError(
                "only the (np.log, np.exp) ufun pair is in the device catalogue")
        assert self._vlims[0] > 0., "log ufun requires positive limits"
        self._ufun, self._log_ufun = ufun, True
        self._eval_ulims()

    # ---- prior --------------------------------------------------------------------------
    def eval_prob(self, values=None):
        """Default prior: 1/length (length in ufun space, no Jacobian) inside the
        domain, 0 / NEARLY_NEGATIVE_INF outside, in this RV's pscale."""
        use_logs = iscomplex(self._pscale)
        inside_val = rescale(self._nlhv, 'log', self._pscale)
        if values is None:
            return inside_val
        outside_val = NEARLY_NEGATIVE_INF if use_logs else 0.
        if isscalar(values):
            return inside_val if self.inside(values) else outside_val
        values = np.asarray(values, dtype=float)
        out = np.full(values.shape, outside_val)
        out[self.inside(values)] = inside_val
        return out

    # ---- sampling grids -----------------------------------------------------------------
    def evaluate(self, values=None):
        """{n}, n > 0 -> deterministic grid of n points (uniform in ufun space,
        mapped back); {0} / {-n} -> random; arrays / scalars pass through."""
        if values is None:
            values = {0}
        if not isunitsetint(values):
            return values if isscalar(values) else np.asarray(values, dtype=float)
        n = list(values)[0]
        assert self._isfinite, \
            "Cannot evaluate {} values for bounds: {}".format(values, self._vlims)
        lo, hi = min(self._ulims), max(self._ulims)
        vals = uniform(lo, hi, n, self._open[0], self._open[1])
        return np.exp(vals) if self._log_ufun else vals

    def __call__(self, values=None):
        """PD of this RV's default prior over ``values`` ({n} grids included):
        probayes/rv.py:323-343."""
        from .pd import PD
        vals = self.evaluate(values)
        if isunitsetint(vals):
            vals = self.evaluate(vals)
        prob = self.eval_prob(vals)
        if not isscalar(vals):
            vals = np.ravel(vals)
            prob = np.ravel(prob)
        name = "{}={}".format(self._name, vals) if isscalar(vals) else self._name + "=[]"
        return PD(name, {self._name: vals}, prob=prob, pscale=self._pscale)

    # ---- deltas ---------------------------------------------------------------------------
    def set_delta(self, delta=None, scale=False, bound=False):
        self._delta = delta
        self._delta_kwds = {'scale': scale, 'bound': bound}

    def __and__(self, other):
        from .rf import RF
        if isinstance(other, RV):
            return RF(self, other)
        if isinstance(other, RF):
            return RF(self, *other.varlist)
        raise TypeError("Unrecognised post-operand type {}".format(type(other)))

    def __repr__(self):
        return self._name
