"""Value-type helpers (reference: probayes/vtypes.py): scalar / unit-set
predicates and the ``uniform`` grid / random sampler that defines DGEI grids."""
import numpy as np

OO = np.inf


def isscalar(var):
    return np.isscalar(var) or (isinstance(var, np.ndarray) and var.ndim == 0)


def isunitset(var, vtype=None):
    if not isinstance(var, set) or len(var) != 1:
        return False
    if vtype is None:
        return True
    return isinstance(list(var)[0], vtype)


def isunitsetint(var):
    """{n} with integer n: the reference's "sample n values" request."""
    return isunitset(var) and isinstance(list(var)[0], (int, np.integer)) \
        and not isinstance(list(var)[0], bool)


def issingleton(var):
    if isunitset(var):
        return True
    return isscalar(var)


def uniform(v_0=0, v_1=1, n=None, ex_0=False, ex_1=False):
    """n > 0: n points on a regular grid between the limits, dropping excluded
    (open) ends -- closed-closed linspace(n) (n == 1 gives the midpoint), open-open
    linspace(n+2)[1:-1], half-open linspace(n+1) minus the open end; n == 0: one
    random scalar; n < 0: -n random points.  (probayes/vtypes.py:169-204)"""
    if not n:
        return np.random.uniform(v_0, v_1)
    if n < 0:
        return np.random.uniform(v_0, v_1, size=-n)
    if ex_0 and ex_1:
        return np.linspace(v_0, v_1, n + 2)[1:-1]
    if ex_0:
        return np.linspace(v_0, v_1, n + 1)[1:]
    if ex_1:
        return np.linspace(v_0, v_1, n + 1)[:-1]
    if n == 1:
        return np.linspace(v_0, v_1, 3)[1:-1]
    return np.linspace(v_0, v_1, n)
