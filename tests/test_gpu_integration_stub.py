"""The reference-side ctypes stub printed in INTEGRATION.md must actually work: it is
extracted from the document, pointed at the built libpbx.so, and its walk compared with
the engine's for the same seed."""
import os
import re
import numpy as np
import pytest
from gpu_util import engine, dev, host

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_stub_runs_and_matches_engine():
    eng = engine()
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    code = next(b for b in blocks if "def walk_mvn" in b)
    lib = os.path.join(ROOT, "probayes_b200", "csrc", "libpbx.so")
    code = code.replace('C.CDLL("libpbx.so")', 'C.CDLL(%r)' % lib)
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    cov = [[2.0, 1.2], [1.2, 2.0]]
    T, C_ = 2000, 64
    x, prob, acc = ns["walk_mvn"]([0., 1.], [0., 0.], cov, T, chains=C_, seed=5)
    state = dev(eng, np.tile(np.array([[0.], [1.]]), (1, C_)))
    out = eng.mh_mvn(state, [0., 0.], cov, T, seed=5, accept="log")
    eng.sync()
    assert np.array_equal(x, host(out["x"]))
    assert np.array_equal(prob, host(out["prob"]))
    assert np.array_equal(acc, host(out["accept_count"]))
    assert 0.4 < acc.mean() / T < 0.8
