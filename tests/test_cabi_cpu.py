"""CPU-side checks of the drop-in boundary: libmsp_b200.so loads without a GPU driver and exports every
symbol include/msp_b200.h declares (and nothing is bound that the header does not declare)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "msp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from medsegpretrainimagenet_b200 import _lib
    names = _header_symbols()
    assert len(names) >= 45
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/msp_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding table and header disagree"
    assert lib.msp_version() == 3


def test_errors_are_reported_not_swallowed():
    """Argument validation happens before any CUDA call: a bad descriptor returns a negative status and a message."""
    from medsegpretrainimagenet_b200 import _lib
    d = _lib.ConvDesc(1, 8, 8, 7, 7, 8, 8, 8, 8, 3, 3, 1, 1, 1, 0, 0, 0)   # C = 7 is not a multiple of 8
    rc = _lib.lib.msp_conv_wgrad_splits(ctypes.byref(d))
    assert rc < 0 and "multiples of 8" in _lib.last_error()
