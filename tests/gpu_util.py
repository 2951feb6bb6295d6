"""Shared helpers for the -m gpu parity tests (all go through the C ABI)."""
import numpy as np
import pytest


def engine():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from probayes_b200.engine import get_engine
    return get_engine(0)


def dev(eng, a):
    return eng.to_device(a)


def tcd_to_tdc(a):
    """oracle layout [T, C, D] -> device layout [T, D, C]."""
    return np.ascontiguousarray(np.transpose(a, (0, 2, 1)))


def host(t):
    return t.detach().cpu().numpy()
