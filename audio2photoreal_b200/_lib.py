"""ctypes binding of liba2p_b200.so (include/a2p_b200.h).  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liba2p_b200.so")

# every symbol include/a2p_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "a2p_abi_version", "a2p_last_error", "a2p_has_tcgen05", "a2p_denoiser_create", "a2p_denoiser_destroy",
    "a2p_packed_weight_bytes", "a2p_denoiser_bind_weights", "a2p_kv_cache_bytes", "a2p_denoiser_set_conditioning",
    "a2p_conditioning_workspace_bytes", "a2p_encode_workspace_bytes", "a2p_denoiser_encode_conditioning", "a2p_workspace_bytes", "a2p_denoiser_forward", "a2p_sampler_step",
    "a2p_sample_loop", "a2p_sample_loop_rng", "a2p_sampler_step_rng", "a2p_profile_forward", "a2p_profile_forward_rows", "a2p_loop_row_groups", "a2p_launch_count",
]


class ModelCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("fmt", "C", "D", "L", "H", "FF", "S2", "max_pos", "split_terms", "reserved")]


class Weight(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ptr", C.c_void_p), ("numel", C.c_int64)]


class A2PError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the in-tree library (rebuilding it first when its source-hash stamp does not match csrc/ + include/)."""
    global _lib
    if _lib is not None:
        return _lib
    from .csrc import build as _b
    if _b._stale():            # content hash of csrc/ + include/ against the stamp written at build time
        import shutil
        if shutil.which(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
            _b.build()
        elif os.path.exists(LIB_PATH):
            raise A2PError(f"{LIB_PATH} was built from different sources and nvcc is not available to rebuild it")
    if not os.path.exists(LIB_PATH):
        raise A2PError(f"{LIB_PATH} is missing and could not be built: the a2p_b200 CUDA extension is required")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    P = C.POINTER
    lib.a2p_abi_version.restype = i32
    lib.a2p_last_error.restype = C.c_char_p
    lib.a2p_has_tcgen05.restype = i32
    lib.a2p_denoiser_create.argtypes = [P(vp), P(ModelCfg)]
    lib.a2p_denoiser_destroy.argtypes = [vp]
    lib.a2p_denoiser_destroy.restype = None
    lib.a2p_packed_weight_bytes.argtypes = [P(ModelCfg)]
    lib.a2p_packed_weight_bytes.restype = sz
    lib.a2p_denoiser_bind_weights.argtypes = [vp, P(Weight), i32, vp, sz, vp]
    lib.a2p_kv_cache_bytes.argtypes = [P(ModelCfg), i32, i32]
    lib.a2p_kv_cache_bytes.restype = sz
    lib.a2p_conditioning_workspace_bytes.argtypes = [P(ModelCfg), i32, i32]
    lib.a2p_conditioning_workspace_bytes.restype = sz
    lib.a2p_encode_workspace_bytes.argtypes = [P(ModelCfg), i32, i32, i32]
    lib.a2p_encode_workspace_bytes.restype = sz
    lib.a2p_denoiser_encode_conditioning.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.a2p_workspace_bytes.argtypes = [P(ModelCfg), i32, i32]
    lib.a2p_workspace_bytes.restype = sz
    lib.a2p_denoiser_set_conditioning.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, sz, vp, sz, vp]
    lib.a2p_denoiser_forward.argtypes = [vp, i32, i32, vp, i32, vp, i32, vp, vp, vp, sz, vp]
    lib.a2p_sampler_step.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp]
    lib.a2p_sample_loop.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, sz, vp]
    lib.a2p_sample_loop_rng.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, C.c_uint64, i64, i32, i32, i32, vp, sz, vp]
    lib.a2p_sampler_step_rng.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp, C.c_uint64, i64, i64, i32, vp, vp, vp]
    lib.a2p_profile_forward.argtypes = [vp, i32, i32, vp, vp, i32, vp, sz, vp, vp, vp, i32]
    lib.a2p_profile_forward.restype = i32
    lib.a2p_profile_forward_rows.argtypes = [vp, i32, i32, i32, i32, vp, vp, i32, vp, sz, vp, vp, vp, i32]
    lib.a2p_profile_forward_rows.restype = i32
    lib.a2p_loop_row_groups.argtypes = [vp, i32, i32]
    lib.a2p_loop_row_groups.restype = i32
    lib.a2p_launch_count.argtypes = [vp]
    lib.a2p_launch_count.restype = i64
    for name in ("a2p_denoiser_create", "a2p_denoiser_bind_weights", "a2p_denoiser_set_conditioning",
                 "a2p_denoiser_encode_conditioning", "a2p_denoiser_forward", "a2p_sampler_step", "a2p_sample_loop", "a2p_sample_loop_rng",
                 "a2p_sampler_step_rng"):
        getattr(lib, name).restype = i32
    if lib.a2p_abi_version() != 1:
        raise A2PError("liba2p_b200.so ABI version mismatch")
    _lib = lib
    return lib


_tlib: Optional[C.CDLL] = None


def load_testing() -> C.CDLL:
    """dlopen liba2p_b200_testing.so: per-kernel test / measurement hooks (include/a2p_b200_testing.h).  Not part of the
    product: only tests/ and scripts/ call this."""
    global _tlib
    if _tlib is None:
        load()                                            # builds both libraries when stale
        path = os.path.join(_HERE, "liba2p_b200_testing.so")
        if not os.path.exists(path):
            raise A2PError(f"{path} is missing")
        _tlib = C.CDLL(path)
        _tlib.a2p_test_last_error.restype = C.c_char_p
    return _tlib


def check_testing(rc: int) -> None:
    if rc != 0:
        raise A2PError(load_testing().a2p_test_last_error().decode("utf-8", "replace"))


def check(rc: int) -> None:
    if rc != 0:
        raise A2PError(load().a2p_last_error().decode("utf-8", "replace"))
