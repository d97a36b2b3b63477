"""CPU checks of the C-ABI boundary: the shared library loads, exports every function include/monosdf_b200.h declares,
and the ctypes binding covers exactly that set.  No compute calls (there is no GPU in the CPU suite)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "monosdf_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msdf_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_surface():
    names = _declared()
    for must in ("msdf_sampler_round", "msdf_hash_encode_forward", "msdf_field_forward", "msdf_field_backward",
                 "msdf_render_forward", "msdf_render_backward", "msdf_fused_adam", "msdf_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from monosdf_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -m monosdf_b200.build"
    h = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(h, name), "libmonosdf_b200.so does not export %s" % name
    assert h.msdf_abi_version() == 3


def test_binding_covers_the_header():
    from monosdf_b200 import _lib
    assert sorted(_lib.exported_symbols()) == _declared()
    _lib.lib()   # sets restype/argtypes for every symbol; raises on a missing one


def test_argument_errors_are_reported_not_thrown():
    """Status code + msdf_last_error() instead of exceptions/aborts (no kernel is launched: n_layers is invalid)."""
    from monosdf_b200 import _lib
    d = _lib.MlpDesc()
    d.n_layers = 99
    e = _lib.EncodingDesc()
    n = _lib.lib().msdf_field_workspace_bytes(ctypes.byref(d), ctypes.byref(e), None, None, 128, _lib.MODE_SDF_ONLY, 0)
    assert n == 0
    assert b"n_layers" in _lib.lib().msdf_last_error()


def test_cpu_tensors_are_refused():
    import torch
    from monosdf_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.ptr(torch.zeros(4))
