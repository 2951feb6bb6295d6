"""PD -- probability distribution: named values + dims + prob + pscale.

Host-side mirror of the reference's result container (probayes/pd.py,
distribution.py, named_dict.py) for what the hot path returns, with the
array algebra of the path running in libpbx:

  * ``prob`` may be backed by a device tensor (a DGEI grid of 134 MB stays in HBM;
    ``.prob`` copies to the host on first access, ``.prob_device`` does not);
  * ``conditionalise`` / ``marginal`` / ``marginalise`` / ``rescaled`` on a
    2-D device-backed PD are the K4 kernels (probayes/pd.py:136-165,168-211,
    214-295,496-499);
  * ``expectation`` / ``quantile`` / ``sorted`` are the "next" rows of the scope
    table; here they operate on marginals / sample sets on the host.

Names follow the reference: ``"mu=[],sigma=[]|x={60}"`` -- ``key=[]`` for array
values, ``key={n}`` for an iid-reduced set, ``key=value`` for scalars; marginal
keys before the ``|``, conditional keys after it.
"""
import collections
import numpy as np

from .pscales import eval_pscale, iscomplex, rescale, div_prob, log_prob, exp_logp
from .vtypes import isscalar, isunitset


def str_margcond(name):
    marg, cond = collections.OrderedDict(), collections.OrderedDict()
    if not name:
        return marg, cond
    left, _, right = name.partition('|')
    for part, dst in ((left, marg), (right, cond)):
        for item in (part.split(',') if part else []):
            key = item.split('=')[0]
            dst[key] = item
    return marg, cond


def margcond_str(marg, cond):
    m = ','.join(marg.values() if isinstance(marg, dict) else marg)
    c = ','.join(cond.values() if isinstance(cond, dict) else cond)
    return m + '|' + c if c else m


def _label(key, val):
    if isinstance(val, set):
        return "{}={}".format(key, val)
    if isscalar(val):
        return "{}={}".format(key, val)
    return key + "=[]"


def _is_tensor(a):
    return type(a).__module__.startswith("torch")


class PD(collections.OrderedDict):
    """PD(name, vals, dims=None, prob=None, pscale=None)."""

    def __init__(self, name, vals=None, dims=None, prob=None, pscale=None, **kwds):
        super().__init__()
        if vals is not None:
            self.update(vals)
        self.update(kwds)
        self._marg, self._cond = str_margcond(name)
        for key in self.keys():
            assert key in self._marg or key in self._cond, \
                "Variable {} not accounted for in name {}".format(key, name)
        self._set_dims(dims)
        self._pscale = eval_pscale(pscale)
        self._prob_dev = None
        self._prob = None
        self._cache = {}            # device-side by-products (marginals of a posterior)
        self.prob = prob

    # ---- bookkeeping ------------------------------------------------------------
    def _set_dims(self, dims):
        self._aresingleton = [isscalar(v) or isunitset(v) for v in self.values()]
        if dims is None:
            dims, k = collections.OrderedDict(), 0
            for key, single in zip(self.keys(), self._aresingleton):
                dims[key] = None if single else k
                k += 0 if single else 1
        self._dims = collections.OrderedDict((k, dims.get(k)) for k in self.keys())
        sizes = {}
        for key, single in zip(self.keys(), self._aresingleton):
            if single:
                continue
            shp = np.shape(self[key])
            if len(shp) > 1 and int(np.prod(shp)) != max(shp):
                # batched samples [C, R]: the array spans consecutive dims from dims[key]
                for off, n in enumerate(shp):
                    sizes[self._dims[key] + off] = int(n)
            else:
                sizes[self._dims[key]] = int(np.size(self[key]))
        self._shape = [sizes[k] for k in sorted(sizes)]
        # refresh labels of scalar / set values in the name
        for key, single in zip(self.keys(), self._aresingleton):
            label = _label(key, self[key]) if single else None
            for group in (self._marg, self._cond):
                if key in group and label is not None:
                    group[key] = label
        self._name = margcond_str(self._marg, self._cond)

    @property
    def name(self):
        return self._name

    @property
    def marg(self):
        return self._marg

    @property
    def cond(self):
        return self._cond

    @property
    def dims(self):
        return self._dims

    @property
    def shape(self):
        return self._shape

    @property
    def ndim(self):
        return len(self._shape)

    @property
    def size(self):
        return int(np.prod(self._shape)) if self._shape else 1

    @property
    def issingleton(self):
        return all(self._aresingleton)

    @property
    def short_name(self):
        m, c = ','.join(self._marg.keys()), ','.join(self._cond.keys())
        return m + '|' + c if c else m

    @property
    def pscale(self):
        return self._pscale

    # ---- probabilities ------------------------------------------------------------
    @property
    def prob(self):
        """numpy array / scalar (device-backed grids are copied on first access)."""
        if self._prob is None and self._prob_dev is not None:
            self._prob = self._prob_dev.detach().cpu().numpy()
        return self._prob

    @prob.setter
    def prob(self, prob):
        self._prob, self._prob_dev = None, None
        if prob is None:
            return
        if _is_tensor(prob):
            if prob.is_cuda:
                self._prob_dev = prob
            else:
                self._prob = prob.numpy()
            shape = list(prob.shape)
        else:
            self._prob = prob
            shape = [] if isscalar(prob) else list(np.shape(prob))
        if self.issingleton:
            assert shape == [], "Singleton vals with non-scalar prob"
        else:
            assert shape == self._shape, \
                "Mismatch in dimensions between values {} and probabilities {}".format(
                    self._shape, shape)

    @property
    def prob_device(self):
        """The device tensor behind ``prob`` (None for host-backed PDs)."""
        return self._prob_dev

    def _new(self, name, vals, dims, prob, pscale=None):
        return PD(name, vals, dims=dims, prob=prob,
                  pscale=self._pscale if pscale is None else pscale)

    # ---- device helpers ---------------------------------------------------------------
    def _engine(self):
        from .engine import get_engine
        return get_engine(self._prob_dev.device.index)

    def rescaled(self, pscale=None):
        """PD with prob converted to ``pscale`` (default linear): pd.py:496-499."""
        dst = eval_pscale(pscale)
        if self._prob_dev is not None and iscomplex(self._pscale) != iscomplex(dst) \
                and self._pscale in (0j, 1.) and dst in (0j, 1.):
            eng = self._engine()
            out = self._prob_dev.clone()
            out = eng.exp_logp_(out) if iscomplex(self._pscale) else eng.log_prob_(out)
            return self._new(self._name, collections.OrderedDict(self), self._dims, out, dst)
        prob = rescale(np.copy(self.prob), self._pscale, dst)
        return self._new(self._name, collections.OrderedDict(self), self._dims, prob, dst)

    # ---- marginal-sum algebra ----------------------------------------------------------
    def _split_keys(self, keys):
        if isinstance(keys, str):
            keys = [keys]
        for key in keys:
            assert key in self._marg, \
                "Key {} not marginal in distribution {}".format(key, self._name)
        return set(keys)

    def marginalise(self, keys):
        """from p(A, key | B) returns p(A | B): exp -> sum over key axes -> clamped
        log (pd.py:136-165)."""
        keys = self._split_keys(keys)
        marg = collections.OrderedDict(self._marg)
        vals, dims, axes, shift = collections.OrderedDict(), collections.OrderedDict(), set(), 0
        for key, single in zip(self.keys(), self._aresingleton):
            if key in keys:
                assert not single, "Cannot marginalise along scalar for key {}".format(key)
                axes.add(self._dims[key])
                marg.pop(key)
                shift += 1
            else:
                if not single:
                    dims[key] = self._dims[key] - shift
                vals[key] = self[key]
        name = margcond_str(marg, self._cond)
        cached = self._cache.get(("marginalise", frozenset(keys)))
        if cached is not None:
            return self._new(name, vals, dims, cached)
        if self._prob_dev is not None and self.ndim == 2 and len(axes) == 1 \
                and iscomplex(self._pscale) and self._pscale == 0j:
            eng = self._engine()
            import torch
            one = torch.ones(1, dtype=torch.float64, device=self._prob_dev.device)
            zero = torch.zeros(1, dtype=torch.float64, device=self._prob_dev.device)
            # sum_axis exp_logp(prob): the posterior kernel with gmax = 0, gsum = 1
            # recomputes log_prob(exp_logp(p)) = p up to the clamp, then row/col sums
            _, rows, cols = eng.grid_posterior(self._prob_dev, zero, one, want_post=False)
            lin = rows if 1 in axes else cols
            return self._new(name, vals, dims, eng.log_prob_(lin))
        prob = rescale(self.prob, self._pscale, 1.)
        prob = rescale(np.sum(prob, axis=tuple(axes), keepdims=False), 1., self._pscale)
        return self._new(name, vals, dims, prob)

    def marginal(self, keys):
        """from p(A, key | B) returns p(key | B) (pd.py:168-211)."""
        keys = self._split_keys(keys)
        others = set()
        for key, single in zip(self.keys(), self._aresingleton):
            if key in self._marg and not single and key not in keys:
                others.add(key)
        scalars = {k for k, s in zip(self.keys(), self._aresingleton) if s and k in self._marg}
        if scalars:
            assert scalars.issubset(keys), \
                "If evaluating marginal, must include all marginal scalars in {}".format(
                    list(self._marg.keys()))
        return self.marginalise(others)

    def conditionalise(self, keys):
        """from p(A, key | B) returns p(A | B, key).  Conditioning on a scalar /
        iid-reduced key normalises the whole array: prob - max; exp; / max(tiny,
        sum); clamped log (pd.py:214-295, arithmetic 285-295)."""
        keys = self._split_keys(keys)
        marg = collections.OrderedDict(self._marg)
        cond = collections.OrderedDict(self._cond)
        normalise = False
        for key, single in zip(self.keys(), self._aresingleton):
            if key in keys:
                cond[key] = marg.pop(key)
                if single:
                    normalise = True
                else:
                    raise NotImplementedError(
                        "conditionalising on an array-valued key is outside the device "
                        "catalogue (only the scalar / iid-reduced normalisation is)")
        name = margcond_str(marg, cond)
        vals = collections.OrderedDict(self)
        if not normalise:
            return self._new(name, vals, self._dims, self._prob_dev
                             if self._prob_dev is not None else self.prob)
        if self._prob_dev is not None and self.ndim == 2 and self._pscale == 0j:
            eng = self._engine()
            r = eng.grid_conditionalise(self._prob_dev)
            out = self._new(name, vals, self._dims, r["post"])
            # the posterior pass already produced both marginal sums
            akeys = [k for k, s in zip(self.keys(), self._aresingleton) if not s]
            k0 = [k for k in akeys if self._dims[k] == 0]
            k1 = [k for k in akeys if self._dims[k] == 1]
            out._cache[("marginalise", frozenset(k1))] = r["marg_mu"]
            out._cache[("marginalise", frozenset(k0))] = r["marg_sigma"]
            return out
        prob = np.asarray(self.prob, dtype=float)
        if iscomplex(self._pscale):
            prob = prob - prob.max()
        prob = rescale(prob, self._pscale, 1.)
        prob = div_prob(prob, np.sum(prob))
        return self._new(name, vals, self._dims, rescale(prob, 1., self._pscale))

    def prod(self, keys):
        """iid product over ``keys``: sum (log pscale) / product (linear) along their
        axis; the values become the set {n} (pd.py:332-370)."""
        keys = self._split_keys(keys)
        marg = collections.OrderedDict(self._marg)
        vals, dims, axes, shift = collections.OrderedDict(), collections.OrderedDict(), [], 0
        for key, single in zip(self.keys(), self._aresingleton):
            if key in keys:
                assert not single, "Cannot apply product along scalar for key {}".format(key)
                if self._dims[key] not in axes:
                    axes.append(self._dims[key])
                    shift += 1
                marg[key] = key + "={}"
                vals[key] = {int(np.size(self[key]))}
            else:
                if not single:
                    dims[key] = self._dims[key] - shift
                vals[key] = self[key]
        prob = np.sum(self.prob, axis=tuple(axes)) if iscomplex(self._pscale) \
            else np.prod(self.prob, axis=tuple(axes))
        return self._new(margcond_str(marg, self._cond), vals, dims, prob)

    def expectation(self, keys=None, exponent=None):
        """E[key] = sum(prob*val) / max(tiny, sum(prob)) over all array axes
        (pd.py:373-405)."""
        keys = list(self._marg.keys()) if keys is None else \
            ([keys] if isinstance(keys, str) else list(keys))
        prob = rescale(self.prob, self._pscale, 1.)
        total = np.sum(prob)
        out = collections.OrderedDict()
        for key, single in zip(self.keys(), self._aresingleton):
            if key in keys:
                val = self[key] if not exponent else self[key] ** exponent
                if single:
                    out[key] = val
                else:
                    shape = [1] * self.ndim
                    shape[self._dims[key]] = -1
                    v = np.asarray(val, dtype=float).reshape(shape)
                    out[key] = div_prob(np.sum(prob * v), total)
            elif key in self._cond:
                out[key] = self[key]
        return out

    def quantile(self, q=0.5):
        """Quantiles of a 1-D distribution from the cumulative probability with
        linear interpolation inside the bracketing cell (pd.py:408-461)."""
        quants = [q] if isscalar(q) else list(q)
        if self.issingleton:
            res = [collections.OrderedDict(self)] * len(quants)
            return res[0] if isscalar(q) else res
        assert self.ndim == 1, "quantile() is implemented for 1-D distributions"
        rav = rescale(np.ravel(self.prob), self._pscale, 1.)
        cum = np.cumsum(rav)
        cum = div_prob(cum, cum[-1])
        idxs = np.maximum(0, np.digitize(np.array(quants), cum) - 1).tolist()
        res = []
        for qq, i in zip(quants, idxs):
            item = collections.OrderedDict()
            for key, single in zip(self.keys(), self._aresingleton):
                if single:
                    item[key] = self[key]
                    continue
                val = np.ravel(self[key])
                i = int(min(i, len(val) - 1))
                if i == len(val) - 1:
                    item[key] = val[i]
                elif abs(rav[i + 1] - rav[i]) < min(qq, 1. - qq):
                    item[key] = float(np.interp(qq, cum[i:i + 2], val[i:i + 2]))
                else:
                    w = rav[i:i + 2]
                    item[key] = float(np.sum(w * val[i:i + 2]) / np.sum(w))
            res.append(item)
        return res[0] if isscalar(q) else res

    def sorted(self, key):
        """Distribution re-ordered by ascending ``key`` (pd.py:464-493)."""
        dim = self._dims[key]
        if dim is None:
            return self._new(self._name, collections.OrderedDict(self), self._dims, self.prob)
        order = np.argsort(np.ravel(self[key]))
        vals = collections.OrderedDict()
        for k, v in self.items():
            vals[k] = np.ravel(v)[order] if self._dims[k] == dim else v
        prob = np.take(self.prob, order, axis=dim)
        return self._new(self._name, vals, self._dims, prob)

    def __repr__(self):
        return "PD({!r}, shape={}, pscale={})".format(self._name, self._shape, self._pscale)
