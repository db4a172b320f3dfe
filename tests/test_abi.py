"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/mlsim.h declares.
No compute call is made here (no GPU); host-only entry points are exercised."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mlsim.h")).read()
    return sorted(set(re.findall(r"\b(mlsim_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(built_lib):
    lib = C.CDLL(built_lib)
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mlsim.h but not exported"


def test_python_binding_covers_the_header(built_lib):
    from ml_accelerated_simulation_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()
    lib = _lib.load()
    assert lib.mlsim_version().startswith(b"mlsim-b200")


def test_library_contains_sm100a_tensor_core_and_bulk_copy_code(built_lib):
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert mnemonic in sass, f"{mnemonic} missing: not a tcgen05/bulk-copy build"


def test_plan_create_fails_loudly_without_a_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ml_accelerated_simulation_b200 import _lib
    lib = _lib.load()
    cfg = _lib.MlsimConfig(64, 64, 1, 2 * 3.141592653589793, 2 * 3.141592653589793, 1e-3, 1e-3, 1.0, 0.1, 4, 1.0, 0, 1.0, 0)
    h = C.c_void_p()
    rc = lib.mlsim_plan_create(C.byref(cfg), C.byref(h))
    assert rc != 0 and not h.value
    assert len(lib.mlsim_last_error()) > 0
    from ml_accelerated_simulation_b200 import Plan, PlanConfig
    with pytest.raises(RuntimeError):
        Plan(PlanConfig())


def test_plan_create_rejects_bad_geometry(built_lib):
    from ml_accelerated_simulation_b200 import _lib
    lib = _lib.load()
    for nx, ny, batch in ((48, 64, 1), (64, 32, 1), (16384, 64, 1), (64, 64, 0)):
        cfg = _lib.MlsimConfig(nx, ny, batch, 6.28, 6.28, 1e-3, 1e-3, 1.0, 0.1, 4, 1.0, 0, 1.0, 0)
        h = C.c_void_p()
        assert lib.mlsim_plan_create(C.byref(cfg), C.byref(h)) == -1       # MLSIM_EINVAL, before any CUDA call
        assert b"power of two" in lib.mlsim_last_error() or b"batch" in lib.mlsim_last_error()


def test_pack_host_sizes_and_bias(built_lib):
    from ml_accelerated_simulation_b200 import _lib
    lib = _lib.load()
    assert lib.mlsim_cnn_pack_host(0, 2, 64, None, None, None, 0) == 4 * 64 * 16 + 256
    assert lib.mlsim_cnn_pack_host(1, 64, 64, None, None, None, 0) == 9 * 64 * 64 * 2 + 256 + 64 * 2 * 4     # + bias + 1x1 shortcut slot
    assert lib.mlsim_cnn_pack_host(2, 64, 2, None, None, None, 0) == 3 * 8 * 48 * 16 + 64 + 16 * 64 * 4
    w = np.arange(2 * 64 * 9, dtype=np.float64).reshape(2, 64, 3, 3) / 1000.0
    b = np.array([0.25, -1.5])
    n = lib.mlsim_cnn_pack_host(2, 64, 2, None, None, None, 0)
    out = np.zeros(n, dtype=np.uint8)
    assert lib.mlsim_cnn_pack_host(2, 64, 2, w.ctypes.data, b.ctypes.data, out.ctypes.data, n) == n
    bias = out[3 * 8 * 48 * 16:3 * 8 * 48 * 16 + 64].view(np.float32)
    assert bias[0] == 0.25 and bias[1] == -1.5 and not bias[2:].any()
    assert lib.mlsim_cnn_pack_host(2, 64, 2, w.ctypes.data, b.ctypes.data, out.ctypes.data, 8) < 0
