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


def test_native_solver_grid_matches_the_oracle_evaluation_times(lib):
    """Host-only seam (no GPU): the evaluation times sab_solve uploads are the ones the oracle's fixed-grid solvers
    visit, in order — midpoint: exact multiples of 1/32 for 16 steps; euler; rk4 (3/8 rule stages)."""
    import numpy as np
    from oracle import restate
    from sam_audio_b200 import _capi
    for name, steps in (("midpoint", 16), ("euler", 8), ("rk4", 4), ("midpoint", 5)):
        seen = []

        def f(t, y):
            seen.append(float(t))
            return -y
        restate.odeint_fixed(f, torch.ones(1), steps, name)
        n = ctypes.c_int(0)
        times = np.zeros(128, np.float32)
        _capi.check(lib.sab_test_solver_grid(_capi.Engine.ODE_METHODS[name], steps, 128, ctypes.addressof(n), times.ctypes.data))
        assert n.value == len(seen)
        assert np.allclose(times[: n.value], np.array(seen, np.float32), rtol=0, atol=1e-7), name
    n = ctypes.c_int(0)
    times = np.zeros(128, np.float32)
    _capi.check(lib.sab_test_solver_grid(0, 16, 128, ctypes.addressof(n), times.ctypes.data))
    assert times[:32].tolist() == [k / 32 for k in range(32)]            # bit-exact multiples of 1/32
    assert lib.sab_test_solver_grid(7, 4, 128, ctypes.addressof(n), times.ctypes.data) != 0 and b"unknown ODE method" in lib.sab_last_error()
