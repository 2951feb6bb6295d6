"""RF -- a random field: an ordered collection of RVs with a joint probability, a
proposal (``delta``) and a transition specification.

Mirror of the reference interface on the hot path (probayes/rf.py:91-113,169-239,
247-304; field.py:220-317): ``set_prob``, ``set_tran``, ``set_tfun``,
``set_delta`` keep their names, argument meaning and reserved keywords
(``pscale``, ``order``, ``tsteps``, ``scale``, ``bound``).  The objects only
*record* the specification; ``catalogue.py`` maps it onto a device kernel or
refuses (there is no Python-callable fallback on the GPU).
"""
import collections
import numpy as np

from .rv import RV
from .pscales import eval_pscale, prod_pscale, prod_rule, iscomplex


class RF:

    def __init__(self, *args):
        rvs = []
        for arg in args:
            if isinstance(arg, RF):
                rvs.extend(arg.varlist)
            else:
                assert isinstance(arg, RV), \
                    "Input not a RV instance but of type: {}".format(type(arg))
                rvs.append(arg)
        names = [rv.name for rv in rvs]
        assert len(set(names)) == len(names), "Duplicate variable names: {}".format(names)
        self._vars = collections.OrderedDict((rv.name, rv) for rv in rvs)
        self._name = ','.join(names)
        self._id = '_and_'.join(names)
        self._delta_type = collections.namedtuple('Delta', names)
        self.Delta = self._delta_type
        self._pscale = prod_pscale([rv.pscale for rv in rvs]) if rvs else 1.
        self._prob, self._prob_args, self._prob_kwds = None, (), {}
        self._order = None
        self._tran, self._tran_args, self._tran_kwds = None, (), {}
        self._tfun = None
        self._tsteps = None
        self._sym_tran = False
        self._delta, self._delta_args, self._delta_kwds = None, (), {}
        self._cond_cov = None

    # ---- members ------------------------------------------------------------------------
    @property
    def name(self):
        return self._name

    @property
    def vars(self):
        return self._vars

    @property
    def varlist(self):
        return list(self._vars.values())

    @property
    def keylist(self):
        return list(self._vars.keys())

    @property
    def keyset(self):
        return frozenset(self._vars.keys())

    @property
    def nvars(self):
        return len(self._vars)

    @property
    def pscale(self):
        return self._pscale

    @property
    def lengths(self):
        return np.array([rv.length for rv in self.varlist], dtype=float)

    def __getitem__(self, key):
        if isinstance(key, int):
            return self.varlist[key]
        return self._vars[key]

    def __len__(self):
        return len(self._vars)

    def __and__(self, other):
        if isinstance(other, RV):
            return RF(*self.varlist, other)
        if isinstance(other, RF):
            return RF(*self.varlist, *other.varlist)
        raise TypeError("Unrecognised post-operand type {}".format(type(other)))

    def parse_key(self, key):
        """RV-object keys -> names (probayes/variable_utils.py:7-51)."""
        return key.name if isinstance(key, RV) else key

    # ---- probability -----------------------------------------------------------------------
    @property
    def prob(self):
        return self._prob

    def set_prob(self, prob=None, *args, **kwds):
        """prob: a scipy.stats callable / distribution (or a Python callable the
        catalogue can identify); reserved keywords ``pscale`` and ``order``
        (``{'x': 0, 'mu': 'loc', 'sigma': 'scale'}``: variable -> positional index or
        keyword of the callable, probayes/expression.py:422-433)."""
        kwds = dict(kwds)
        if 'pscale' in kwds:
            self._pscale = eval_pscale(kwds.pop('pscale'))
        elif self.nvars:
            self._pscale = prod_pscale([rv.pscale for rv in self.varlist])
        self._order = kwds.pop('order', None)
        kwds.pop('passdims', None)
        self._prob, self._prob_args, self._prob_kwds = prob, tuple(args), kwds

    def __call__(self, values=None):
        """Joint PD of the independent default RV priors over a dictionary of values /
        {n} grid requests (probayes/rf.py:565-581 without a set_prob: rf_utils.py:10-42)."""
        from .pd import product
        values = {} if values is None else values
        values = {self.parse_key(k): v for k, v in values.items()}
        return product(*[rv(values.get(rv.name)) for rv in self.varlist])

    def eval_prior(self, values):
        """Product of the independent default RV priors (rf_utils.py:10-42)."""
        rvs = self.varlist
        probs = [rv.eval_prob(values[rv.name]) for rv in rvs]
        return prod_rule(*probs, pscales=[rv.pscale for rv in rvs], pscale=self._pscale)

    # ---- transitions -------------------------------------------------------------------------
    @property
    def tran(self):
        return self._tran

    @property
    def tfun(self):
        return self._tfun

    @property
    def tsteps(self):
        return self._tsteps

    def set_tran(self, tran=None, *args, **kwds):
        """tran: callable proposal density q(**kwds) (symmetric), a 2-tuple (q, r)
        of callables (asymmetric pair), a covariance matrix (proposal = Cholesky
        factor times the delta draw) or ``scipy.stats.multivariate_normal`` with
        (mean, cov) arguments and ``tsteps`` for Gibbs sampling."""
        kwds = dict(kwds)
        self._tsteps = kwds.pop('tsteps', None)
        if self._tsteps:
            assert type(self._tsteps) is int, "Input tsteps must be int"
        self._tran, self._tran_args, self._tran_kwds = tran, tuple(args), kwds
        self._tfun, self._cond_cov = None, None
        self._sym_tran = not isinstance(tran, tuple)
        if tran is None:
            return
        if isinstance(tran, np.ndarray):
            message = "Non-callable non-scalar tran objects must be a square 2D Numpy " \
                      "array of size corresponding to number of variables {}".format(self.nvars)
            assert tran.ndim == 2 and tran.shape == (self.nvars, self.nvars), message
            assert self._tsteps is None, \
                "Setting tsteps not supported for covariance transitions"
            self._tfun = np.linalg.cholesky(tran)
            return
        from . import catalogue
        if catalogue.is_scipy_mvn(tran):
            from .cond_cov import CondCov
            mean, cov = catalogue.mvn_args(args, kwds)
            lims = np.array([rv.ulims for rv in self.varlist])
            self._cond_cov = CondCov(mean, cov, lims)

    def set_tfun(self, tfun=None, *args, **kwds):
        if tfun is None:
            self._tfun = None
            return
        if isinstance(tfun, np.ndarray):
            message = "Non-callable tran objects must be a triangular 2D Numpy array " \
                      "of size corresponding to number of variables {}".format(self.nvars)
            assert tfun.ndim == 2 and tfun.shape == (self.nvars, self.nvars), message
            assert np.allclose(tfun, np.tril(tfun)) or np.allclose(tfun, np.triu(tfun)), message
            self._tfun = tfun
            return
        raise NotImplementedError(
            "callable conditional samplers (set_tfun with Python functions) cannot run on "
            "the device; use a scipy.stats.multivariate_normal transition for Gibbs")

    # ---- deltas ---------------------------------------------------------------------------------
    @property
    def delta(self):
        return self._delta

    def set_delta(self, delta=None, *args, **kwds):
        """delta: ``[d]`` -> per-RV uniform in (-d, d); ``(d,)`` -> spherical (a cube
        sample rescaled to length d); a frozen ``scipy.stats.norm`` -> iid normal
        draws with that scale; a dict / Delta of per-RV specs.  Keywords ``scale``
        (multiply by the RV lengths) and ``bound`` as in the reference
        (field.py:220-317).  Arbitrary callables cannot run on the device."""
        self._delta = delta
        self._delta_args = tuple(args)
        self._delta_kwds = {'scale': bool(kwds.pop('scale', False)),
                            'bound': kwds.pop('bound', False)}
        assert not kwds, "Unknown delta keywords: {}".format(list(kwds))

    def __repr__(self):
        return "RF({})".format(self._name)
