"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/pbx.h declares (no compute calls here)."""
import ctypes as C
import os
import re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pbx.h")).read()
    return sorted(set(re.findall(r"PBX_API\s+[\w\s\*]+?\b(pbx_\w+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    from probayes_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from probayes_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libpbx.so does not export " + n
        assert n in _lib.SIGNATURES, "no ctypes signature for " + n
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_error_string(lib):
    assert lib.pbx_version() == 100
    n = C.c_int(-1)
    rc = lib.pbx_device_count(C.byref(n))
    if rc != 0:                      # CPU-only box: fails loudly with a message
        assert n.value == 0
        assert len(lib.pbx_last_error()) > 0


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout (checked against a gcc-compiled probe)."""
    import subprocess
    import tempfile
    from probayes_b200 import _lib
    src = r'''
#include <stdio.h>
#include "pbx.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(pbx_mh_mvn_params), sizeof(pbx_mh_normreg_params),
         sizeof(pbx_gibbs_mvn_params), sizeof(pbx_devinfo));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"),
                        "-o", os.path.join(d, "p")], check=True)
        out = subprocess.run([os.path.join(d, "p")], capture_output=True, text=True,
                             check=True).stdout.split()
    sizes = [int(v) for v in out]
    assert sizes == [C.sizeof(_lib.MhMvnParams), C.sizeof(_lib.MhNormregParams),
                     C.sizeof(_lib.GibbsMvnParams), C.sizeof(_lib.DevInfo)]


def test_no_cpu_fallback_without_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from probayes_b200.engine import get_engine
    from probayes_b200._lib import PbxError
    with pytest.raises(PbxError):
        get_engine()
