"""CPU-side checks of the C-ABI boundary: the library builds (nvcc cross-compiles
without a GPU), loads, and exports every symbol include/b2rl.h declares — no
compute calls here."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(REPO, "include", "b2rl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2rl_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from distributed_rl_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_declares_expected_surface():
    syms = _declared_symbols()
    for must in ("b2rl_replay_create", "b2rl_replay_push", "b2rl_tree_sample", "b2rl_tree_update",
                 "b2rl_replay_gather", "b2rl_apex_target", "b2rl_r2d2_target", "b2rl_vtrace"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    from distributed_rl_b200 import _lib
    for s in _declared_symbols():
        assert hasattr(lib, s), f"libb2rl.so does not export {s}"
        assert s in _lib.SIGNATURES, f"ctypes binding missing for {s}"
    assert set(_lib.SIGNATURES) == set(_declared_symbols())


def test_version_and_error_string(lib):
    assert lib.b2rl_version() >= 100
    assert isinstance(lib.b2rl_last_error(), bytes)
    assert lib.b2rl_launch_count() >= 0


def test_invalid_arguments_are_reported_not_crashed(lib):
    from distributed_rl_b200._lib import ReplayDesc
    # null pointers are rejected before any CUDA call
    assert lib.b2rl_replay_create(None, None) < 0
    assert b"null" in lib.b2rl_last_error()
    d = ReplayDesc(); d.capacity = 0; d.n_fields = 0; d.device = 0
    h = ctypes.c_void_p()
    assert lib.b2rl_replay_create(ctypes.byref(d), ctypes.byref(h)) < 0
    assert lib.b2rl_apex_target(*([None] * 7), 0, 0, 0.0, 0.0, *([None] * 6)) < 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from distributed_rl_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.B2RLError, match="no CPU fallback"):
        _lib.load()


def test_device_replay_refuses_cpu_device(lib):
    import torch
    from distributed_rl_b200 import replay
    from distributed_rl_b200._lib import B2RLError
    if torch.cuda.is_available():
        pytest.skip("GPU present; covered by the gpu tests")
    with pytest.raises((B2RLError, RuntimeError, AssertionError)):
        replay.DeviceReplay(16, fields=(), device="cpu")
