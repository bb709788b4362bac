"""The C-ABI library loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from livelyspeaker_b200 import _cabi

HEADER = os.path.join(ROOT, "include", "livelyspeaker_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ls_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), "libls_b200.so does not export %s" % n
    assert sorted(_cabi.EXPORTS) == names, "binding list and header disagree"
    assert _cabi.load_library().ls_abi_version() == 1


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_cabi.LsConfig) == 12 * 4
    assert ctypes.sizeof(_cabi.LsStepParams) == 4 * 4 + 8 * 4
    assert _cabi.LsStepParams.c.offset == 16
    assert ctypes.sizeof(_cabi.LsStepIO) == 64 and _cabi.LsStepIO.x_prev.offset == 48      # static_assert'ed in ls_api.cu
    from livelyspeaker_b200 import sag
    assert ctypes.sizeof(sag.LsSagLayer) == 18 * 8
    assert ctypes.sizeof(sag.LsSagWeights) == 8 * 4 + 6 * 8 + sag.SAG_MAX_LAYERS * 18 * 8 == 1232   # ... in ls_sag.cu
    assert sag.LsSagWeights.layer.offset == 80


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _cabi.load_library()
    h = ctypes.c_void_p()
    cfg = _cabi.LsConfig(9, 3, 34, 1, 512, 8, 36267, 1400, 0, 4, 1000, 0)
    rc = lib.ls_create(ctypes.byref(h), ctypes.byref(cfg))
    assert rc < 0 and not h.value
    assert b"not available" in lib.ls_last_error(None)
    bad = _cabi.LsConfig(9, 3, 34, 1, 256, 8, 36267, 1400, 0, 4, 1000, 0)
    assert lib.ls_create(ctypes.byref(h), ctypes.byref(bad)) == -5   # LS_EUNSUPPORTED
