"""Import shim for the LIVE reference -- development container only.

TEST INFRASTRUCTURE (see oracle/__init__.py).  /root/reference does not exist
on the GPU box; callers must check ``available()`` and fall back to the
committed fixtures under tests/golden/.

Two harness-side shims, reference files untouched:
  * a stub ``h5py`` (imported at module top by probayes/pd_utils.py:8, used only
    by the HDF5 serialisers which are off the hot path);
  * drop ``'fit'`` from ``probayes.prob.SCIPY_DIST_METHODS`` -- scipy >= 1.1x
    frozen ``multivariate_normal`` has no ``.fit`` (probayes/prob.py:212-217).
"""
import os
import sys
import tempfile
import contextlib

REFERENCE_ROOT = os.environ.get("PROBAYES_REFERENCE", "/root/reference")
_pb = None


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "probayes"))


def load():
    """Returns the reference ``probayes`` module (cached)."""
    global _pb
    if _pb is not None:
        return _pb
    if not available():
        raise RuntimeError("live reference not present at " + REFERENCE_ROOT)
    try:
        import h5py  # noqa: F401
    except ImportError:
        shim = os.path.join(tempfile.gettempdir(), "pbx_ref_shim")
        os.makedirs(os.path.join(shim, "h5py"), exist_ok=True)
        open(os.path.join(shim, "h5py", "__init__.py"), "a").close()
        sys.path.insert(0, shim)
    sys.path.insert(0, REFERENCE_ROOT)
    import probayes as pb
    import probayes.prob as _pp
    if "fit" in _pp.SCIPY_DIST_METHODS:
        _pp.SCIPY_DIST_METHODS.remove("fit")
    _pb = pb
    return pb


@contextlib.contextmanager
def injected_uniform(stream):
    """Replaces ``np.random.uniform`` by ``low + (high-low)*next(stream)`` (the
    same affine map numpy applies to its raw double) for the duration of the
    block, so every uniform the reference draws -- proposal deltas
    (probayes/variable.py:631-633), thresholds (sp_utils.py:30-31), CondCov cdf
    draws (vtypes.py:186) -- comes from a recorded stream in call order."""
    import numpy as np
    it = iter(stream)
    orig = np.random.uniform

    def fake(low=0.0, high=1.0, size=None):
        if size is None:
            return low + (high - low) * next(it)
        n = int(np.prod(size))
        return (low + (high - low) * np.array([next(it) for _ in range(n)])).reshape(size)

    np.random.uniform = fake
    try:
        yield
    finally:
        np.random.uniform = orig
