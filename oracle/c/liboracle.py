"""ctypes loader + numpy wrappers for oracle/_build/liboracle.so.
TEST INFRASTRUCTURE (see oracle/__init__.py)."""
import ctypes as C
import os
import numpy as np
from . import build as _build

_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def load():
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.exists(path):
            path = _build.build()
        _lib = C.CDLL(path)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def num_threads():
    return int(load().orc_num_threads())


def use_all_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm wants every core it may use."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    load().orc_set_num_threads(C.c_int(n))
    return num_threads()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _mvn_consts(mean, cov, reorder=True):
    """Same host-side constants the device path derives, computed independently
    from the numpy oracle (scipy _PSD whitening + prob.py:349-358 ordering)."""
    from .. import np_oracle as o
    mean = np.asarray(mean, dtype=float)
    d = len(mean)
    U, log_pdet = o.mvn_whiten(cov)
    order = o.mvn_value_order(d) if reorder else list(range(d))
    W = np.empty_like(U)
    mean_nat = np.empty_like(mean)
    for j, i in enumerate(order):
        W[i] = U[j]
        mean_nat[i] = mean[j]
    return np.ascontiguousarray(mean_nat), np.ascontiguousarray(W), d * o.LOG_2PI + log_pdet


def mh_mvn_walk(init, mean, cov, steps, seed, step0=0, chain0=0, scale=1.0, log_pscale=False,
                accept="reference", record=True, reorder=True):
    """init [C, D] -> dict(x [T, D, C], prob [T, C], u [T, C], final [C, D], accept_count)."""
    lib = load()
    init = np.ascontiguousarray(init, dtype=float)
    Cn, D = init.shape
    mean_nat, W, norm_c = _mvn_consts(mean, cov, reorder)
    sc = np.ascontiguousarray(np.broadcast_to(np.asarray(scale, float), (D,)))
    T = int(steps)
    ox = np.empty((T, D, Cn)) if record else None
    op = np.empty((T, Cn)) if record else None
    ou = np.empty((T, Cn), dtype=np.uint8) if record else None
    fin = np.empty((Cn, D))
    acc = np.zeros(Cn, dtype=np.int64)
    lib.orc_mh_mvn_walk(C.c_int(Cn), C.c_int(D), C.c_int(T), _p(init), _p(mean_nat), _p(W),
                        C.c_double(norm_c), _p(sc), C.c_uint64(seed), C.c_int64(step0),
                        C.c_int64(chain0), C.c_int(1 if log_pscale else 0),
                        C.c_int(0 if accept == "reference" else 1), _p(ox), _p(op), _p(ou),
                        _p(fin), _p(acc))
    return dict(x=ox, prob=op, u=None if ou is None else ou.astype(bool), final=fin,
                accept_count=acc)


def normreg_logjoint(theta, x_obs, y_obs, lims, ex, log_ufun):
    lib = load()
    theta = np.ascontiguousarray(theta, dtype=float)
    Cn, P = theta.shape
    y = np.ascontiguousarray(y_obs, dtype=float)
    x = np.ascontiguousarray(x_obs, dtype=float) if x_obs is not None else y
    out = np.empty(Cn)
    lib.orc_normreg_logjoint(C.c_int(Cn), C.c_int(P), _p(theta), C.c_int64(len(y)), _p(x), _p(y),
                             _p(np.ascontiguousarray(lims, dtype=float)),
                             _p(np.ascontiguousarray(ex, dtype=np.int32)),
                             _p(np.ascontiguousarray(log_ufun, dtype=np.int32)), _p(out))
    return out


def grid_norm_logjoint(x_obs, mu, sigma, lp_mu, lp_sigma):
    lib = load()
    x = np.ascontiguousarray(x_obs, dtype=float)
    mu = np.ascontiguousarray(mu, dtype=float)
    sg = np.ascontiguousarray(sigma, dtype=float)
    out = np.empty((len(mu), len(sg)))
    lib.orc_grid_norm_logjoint(C.c_int64(len(x)), _p(x), C.c_int(len(mu)), _p(mu),
                               C.c_int(len(sg)), _p(sg),
                               _p(np.ascontiguousarray(lp_mu, dtype=float)),
                               _p(np.ascontiguousarray(lp_sigma, dtype=float)), _p(out))
    return out


def grid_posterior(lj):
    lib = load()
    lj = np.ascontiguousarray(lj, dtype=float)
    M, S = lj.shape
    post = np.empty((M, S)); mm = np.empty(M); ms = np.empty(S)
    lib.orc_grid_posterior(C.c_int(M), C.c_int(S), _p(lj), _p(post), _p(mm), _p(ms))
    return post, mm, ms


def gibbs_mvn_walk(x, mean, coef, stdv, cdfs, steps, seed=0, step0=0, chain0=0, runif=None):
    """x [C, d] (copied); returns final x [C, d]."""
    lib = load()
    x = np.array(x, dtype=float, order="C")
    Cn, d = x.shape
    ru = None if runif is None else np.ascontiguousarray(runif, dtype=float)
    lib.orc_gibbs_mvn_walk(C.c_int(Cn), C.c_int(d), C.c_int(int(steps)), _p(x),
                           _p(np.ascontiguousarray(mean, dtype=float)),
                           _p(np.ascontiguousarray(coef, dtype=float)),
                           _p(np.ascontiguousarray(stdv, dtype=float)),
                           _p(np.ascontiguousarray(cdfs[:, 0], dtype=float)),
                           _p(np.ascontiguousarray(cdfs[:, 1], dtype=float)),
                           C.c_uint64(seed), C.c_int64(step0), C.c_int64(chain0), _p(ru))
    return x
