"""CondCov -- covariance-matrix-based conditional sampling, batched on the GPU.

Mirrors the reference class of the same name (probayes/cond_cov.py:11-65): the
constructor derives, per coordinate i of a multivariate normal N(mean, cov),

    coef_i = cov[i, -i] cov[-i, -i]^-1           (regression row, cond_cov.py:30-35)
    stdv_i = sqrt(cov[i, i] - coef_i cov[-i, i])  (conditional stdv, 36)
    cdfs_i = Phi((lims_i - mean_i) / stdv_i)      (fixed truncation limits, 38-39)

and ``interp`` draws x_i = ndtri(U(cdf_lo, cdf_hi)) stdv_i + mean_i +
coef_i (x_-i - mean_-i) (cond_cov.py:42-65).  Here the constants are computed on
the host exactly as the reference does (n small inverses of size n-1), and the
draws run on the device for whole batches of chains (``Engine.gibbs_mvn``).
"""
import numpy as np
import scipy.stats


class CondCov:

    def __init__(self, mean, cov, lims):
        self._mean = np.atleast_1d(np.asarray(mean, dtype=np.float64))
        self._cov = np.atleast_2d(np.asarray(cov, dtype=np.float64))
        self._inv = np.linalg.inv(self._cov)
        self._lims = np.atleast_2d(lims) - np.expand_dims(self._mean, -1)
        self._n = len(self._mean)
        assert len(self._cov) == self._n, \
            "Means and covariance matrix incommensurate"
        n = self._n
        self._stdv = np.empty(n, dtype=float)
        self._coef = [None] * n
        for i in range(n):
            idx = [j for j in range(n) if j != i]
            ll = self._cov[idx, i].reshape(n - 1, 1)
            ru = self._cov[i, idx].reshape(1, n - 1)
            sub = self._cov[np.ix_(idx, idx)]
            self._coef[i] = ru.dot(np.linalg.inv(sub))
            self._stdv[i] = np.sqrt(self._cov[i, i] - float(self._coef[i].dot(ll)[0, 0]))
        self._cdfs = np.array([scipy.stats.norm.cdf(lim, loc=0., scale=self._stdv[i])
                               for i, lim in enumerate(self._lims)])

    # -- dense forms used by the device kernel ---------------------------------
    @property
    def n(self):
        return self._n

    @property
    def mean(self):
        return self._mean

    @property
    def cov(self):
        return self._cov

    @property
    def stdv(self):
        return self._stdv

    @property
    def cdfs(self):
        return self._cdfs

    def coef_matrix(self):
        """[n, n] regression rows with a zero diagonal."""
        n = self._n
        out = np.zeros((n, n))
        for i in range(n):
            idx = [j for j in range(n) if j != i]
            out[i, idx] = np.ravel(self._coef[i])
        return out

    def interp(self, *args, cond_pdf=False):
        """Reference signature (cond_cov.py:42): exactly one argument is the set
        ``{0}`` marking the coordinate to draw; the others are scalars or arrays
        of a common length (one entry per chain).  Runs one coordinate update on
        the device and returns the drawn values (array, or scalar for scalar
        inputs)."""
        from .engine import get_engine
        idx = None
        vals = []
        for i, arg in enumerate(args):
            if isinstance(arg, set):
                if idx is not None:
                    raise ValueError("Only one argument can be interpolated at a time")
                idx = i
                vals.append(None)
            else:
                vals.append(np.atleast_1d(np.asarray(arg, dtype=np.float64)))
        assert idx is not None, "No variable specified for interpolation"
        if cond_pdf:
            raise NotImplementedError("cond_pdf=True is not in the device catalogue")
        C = max(len(v) for v in vals if v is not None)
        x = np.empty((self._n, C))
        for i, v in enumerate(vals):
            x[i] = self._mean[i] if v is None else np.broadcast_to(v, (C,))
        eng = get_engine()
        state = eng.to_device(x)
        seed = int(np.random.randint(0, 2 ** 31 - 1))      # the reference draws from np.random
        eng.gibbs_mvn(state, self, 1, seed=seed, step0=idx, record=False)
        out = state[idx].cpu().numpy()
        scalar = all(np.ndim(a) == 0 for a in args if not isinstance(a, set))
        return float(out[0]) if scalar else out
