"""CPU: the C-ABI shared library loads and exports every symbol include/*.h declares; entry points
fail loudly (no fallback) when there is no GPU."""
import ctypes
import glob
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from sam_audio_b200 import _capi
    return _capi.lib()


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(sab_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"


def test_python_binding_covers_header():
    from sam_audio_b200 import _capi
    assert set(declared_symbols()) == set(_capi.EXPORTS)


def test_version_and_error_string(lib):
    assert lib.sab_version() == 1
    assert isinstance(lib.sab_last_error(), bytes)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(lib):
    from sam_audio_b200 import _capi
    from sam_audio_b200.config import stand_in_config
    cfg = _capi.make_config(stand_in_config("sam-audio-tiny"))
    h = ctypes.c_void_p()
    rc = lib.sab_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc != 0 and b"no CUDA device" in lib.sab_last_error()
    with pytest.raises(RuntimeError):
        from sam_audio_b200.model import build_synthetic_model
        build_synthetic_model("sam-audio-tiny")


def test_config_struct_layout_matches_header():
    """sizeof(sab_config) from the ctypes mirror == what the C side compiled (13 scalars + 5 + 2x8 ints)."""
    from sam_audio_b200 import _capi
    assert ctypes.sizeof(_capi.SabConfig) == 4 * (13 + 5 + 16)
