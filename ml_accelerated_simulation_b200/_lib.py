"""ctypes binding of the C ABI declared in include/mlsim.h.

There is NO fallback: if the shared library is missing or does not load, importing the hot-path
modules raises.  Build it with ``python -m ml_accelerated_simulation_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmlsim_b200.so")

MLSIM_OK = 0


class MlsimConfig(C.Structure):
    """mlsim_config of include/mlsim.h (field order and types must match)."""
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("batch", C.c_int32),
        ("lx", C.c_double), ("ly", C.c_double), ("dt", C.c_double),
        ("viscosity", C.c_double), ("density", C.c_double), ("drag", C.c_double),
        ("forcing_wavenumber", C.c_int32), ("forcing_scale", C.c_double),
        ("advection", C.c_int32), ("lw_dt_scale", C.c_double),
        ("device", C.c_int32),
        ("slab_rank", C.c_int32), ("slab_world", C.c_int32),
        ("solver_blocks", C.c_int32), ("host_blocks", C.c_int32), ("separate_first_layer", C.c_int32),
        ("cluster_projection", C.c_int32), ("step_graph", C.c_int32),
    ]


class MlsimError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mlsim error {code}: {msg}")
        self.code = code


_SIGNATURES = {
    "mlsim_plan_create": (C.c_int, [C.POINTER(MlsimConfig), C.POINTER(C.c_void_p)]),
    "mlsim_plan_destroy": (C.c_int, [C.c_void_p]),
    "mlsim_cnn_begin": (C.c_int, [C.c_void_p, C.c_int32]),
    "mlsim_cnn_set_layer": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_int32, C.c_int32]),
    "mlsim_cnn_set_shortcut": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "mlsim_cnn_finalize": (C.c_int, [C.c_void_p]),
    "mlsim_cnn_pack_host": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "mlsim_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double,
                             C.c_void_p]),
    "mlsim_step_f64": (C.c_int, [C.c_void_p] * 6 + [C.c_int32, C.c_int32, C.c_double, C.c_void_p]),
    "mlsim_rollout_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_double]),
    "mlsim_set_host_blocks": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int32]),
    "mlsim_explicit_terms": (C.c_int, [C.c_void_p] * 5 + [C.c_void_p]),
    "mlsim_project": (C.c_int, [C.c_void_p] * 4 + [C.c_void_p]),
    "mlsim_spectral_multiply": (C.c_int, [C.c_void_p] * 4 + [C.c_void_p]),
    "mlsim_normalize_speed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "mlsim_divergence": (C.c_int, [C.c_void_p] * 4 + [C.c_void_p]),
    "mlsim_cnn_forward": (C.c_int, [C.c_void_p] * 4 + [C.c_double, C.c_void_p]),
    "mlsim_cnn_correct": (C.c_int, [C.c_void_p] * 3 + [C.c_double, C.c_void_p]),
    "mlsim_slab_geometry": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "mlsim_slab_stage_a": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 6),
    "mlsim_slab_divergence_rfft": (C.c_int, [C.c_void_p] * 4),
    "mlsim_slab_stage_b": (C.c_int, [C.c_void_p] * 4),
    "mlsim_slab_stage_c": (C.c_int, [C.c_void_p] * 5),
    "mlsim_slab_set_peers": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "mlsim_slab_halo_push": (C.c_int, [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_void_p]),
    "mlsim_slab_cnn_correct": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "mlsim_downsample_staggered": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 4 + [C.c_void_p]),
    "mlsim_datagen_pair": (C.c_int, [C.c_void_p] * 6 + [C.c_int32, C.c_int32, C.c_void_p]),
    "mlsim_finite_check_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "mlsim_finite_check_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mlsim_launches_per_step": (C.c_int, [C.c_void_p, C.c_int32]),
    "mlsim_step_graph_launches": (C.c_int64, [C.c_void_p]),
    "mlsim_workspace_bytes": (C.c_int64, [C.c_void_p]),
    "mlsim_time_kernel": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.c_void_p]),
    "mlsim_conv_debug": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "mlsim_conv_debug_first": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "mlsim_profile_enable": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "mlsim_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "mlsim_last_error": (C.c_char_p, []),
    "mlsim_version": (C.c_char_p, []),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load() -> C.CDLL:
    """Load libmlsim_b200.so (once) and declare every entry point of include/mlsim.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found. The B200 hot path has no CPU/PyTorch fallback; build the CUDA library with "
            "`python -m ml_accelerated_simulation_b200.build` (needs nvcc, sm_100a).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != MLSIM_OK:
        raise MlsimError(rc, load().mlsim_last_error().decode("utf-8", "replace"))
