"""ctypes binding of the C-ABI library (include/mask_bev_b200.h).

There is NO CPU fallback and no pure-PyTorch path: if the shared library is missing or a symbol does not
resolve this module raises, loudly, at first use.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

MAX_BATCH = 128
MAX_LAYERS = 4
MAX_UNITS = 128
MAX_POINT_DIM = 8
ABI_VERSION = 11
GEMM_AUTO, GEMM_FMA, GEMM_TCGEN05, GEMM_TCGEN05_BF16 = 0, 1, 2, 3
LN_WALK_RUNS, LN_WALK_FRAMES = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libmask_bev_b200.so")


class MbevGeometry(ctypes.Structure):
    _fields_ = [("range", c_float * 6), ("voxel", c_float * 3), ("grid", c_int32 * 3),
                ("max_points", c_int32), ("max_voxels", c_int32), ("num_feats", c_int32),
                ("strict_filter", c_int32)]


class MbevPfnParams(ctypes.Structure):
    _fields_ = [("num_layers", c_int32), ("in_dim", c_int32 * MAX_LAYERS), ("units", c_int32 * MAX_LAYERS),
                ("weight", c_void_p * MAX_LAYERS), ("scale", c_void_p * MAX_LAYERS),
                ("shift", c_void_p * MAX_LAYERS),
                ("with_cluster_center", c_int32), ("with_voxel_center", c_int32), ("with_distance", c_int32),
                ("legacy", c_int32), ("voxel_center_dims", c_int32),
                ("vx", c_float), ("vy", c_float), ("vz", c_float),
                ("x_offset", c_float), ("y_offset", c_float), ("z_offset", c_float),
                ("gemm_path", c_int32)]


class MbevFrameAugment(ctypes.Structure):
    _fields_ = [("cos_t", c_double), ("sin_t", c_double), ("drop_prob", c_float), ("jitter_std", c_float * 4),
                ("jitter_max", c_float * 4), ("flip_x", c_int32), ("flip_y", c_int32), ("rotate", c_int32),
                ("jitter", c_int32)]


class MbevAugment(ctypes.Structure):
    _fields_ = [("frames", c_void_p), ("drop_u", c_void_p), ("noise", c_void_p), ("seed", c_uint64),
                ("points_out", c_void_p)]


_PTRS = c_void_p * MAX_LAYERS
_G = POINTER(MbevGeometry)
_P = POINTER(MbevPfnParams)
_v = c_void_p

# name -> (restype, argtypes); must list every symbol include/mask_bev_b200.h declares
SIGNATURES = {
    "mbev_abi_version": (c_int, []),
    "mbev_build_info": (c_char_p, []),
    "mbev_status_string": (c_char_p, [c_int]),
    "mbev_pillar_capacity": (c_int64, [POINTER(c_int64), c_int, _G]),
    "mbev_voxelize_workspace_bytes": (c_int, [_G, c_int, c_int64, POINTER(c_size_t)]),
    "mbev_voxelize": (c_int, [_v, POINTER(c_int64), c_int, _G, _v, _v, _v, _v, _v, c_int64, _v, c_size_t, _v]),
    "mbev_voxelize_augmented": (c_int, [_v, POINTER(c_int64), c_int, _G, _v, _v, _v, _v, _v, _v, c_int64, _v, c_size_t, _v]),
    "mbev_gather_voxels": (c_int, [_v, _v, _v, _v, c_int64, c_int, c_int, _v, _v]),
    "mbev_pfn_workspace_bytes": (c_int, [_P, c_int, c_int64, c_int, POINTER(c_size_t)]),
    "mbev_pfn_path": (c_int, [_P, c_int]),
    "mbev_pfn_forward": (c_int, [_v, c_int, _v, _v, _v, _v, c_int64, c_int, _P, _v, _v, c_size_t, _v]),
    "mbev_pfn_forward_train": (c_int, [_v, c_int, _v, _v, _v, _v, c_int64, c_int, _P, POINTER(_PTRS),
                                       POINTER(_PTRS), c_float, _v, _v, _v, _v, c_size_t, _v]),
    "mbev_pfn_backward_workspace_bytes": (c_int, [_P, c_int, c_int64, c_int64, POINTER(c_size_t)]),
    "mbev_pfn_backward": (c_int, [_v, c_int, _v, _v, _v, _v, c_int64, c_int, c_int64, _P, _v, _v, c_float, c_int, _v,
                                  POINTER(_PTRS), POINTER(_PTRS), POINTER(_PTRS), _v, c_size_t, _v]),
    "mbev_pfn_forward_train_rows": (c_int, [_v, c_int, _v, _v, _v, _v, c_int64, c_int, c_int64, _P, POINTER(_PTRS),
                                            POINTER(_PTRS), c_float, _v, _v, _v, _v, c_size_t, _v]),
    "mbev_pfn_backward_rows": (c_int, [_v, c_int64, c_int, c_int, c_int64, _P, c_float, _v, POINTER(_PTRS),
                                       POINTER(_PTRS), POINTER(_PTRS), _v, c_size_t, _v]),
    "mbev_build_cell_table": (c_int, [_v, _v, c_int64, c_int, c_int, c_int, _v, _v]),
    "mbev_scatter_forward": (c_int, [_v, _v, c_int, c_int, c_int, c_int, _v, _v]),
    "mbev_scatter_stream_supported": (c_int, [c_int, c_int, c_int, _v]),
    "mbev_scatter_forward_stream": (c_int, [_v, _v, c_int, c_int, c_int, c_int, _v, c_int, _v]),
    "mbev_scatter_forward_bf16": (c_int, [_v, _v, c_int, c_int, c_int, c_int, _v, _v]),
    "mbev_scatter_forward_nhwc": (c_int, [_v, _v, c_int, c_int, c_int, c_int, _v, _v]),
    "mbev_scatter_backward_nhwc": (c_int, [_v, _v, _v, _v, c_int64, c_int, c_int, c_int, c_int, _v, _v]),
    "mbev_scatter_backward": (c_int, [_v, _v, c_int, c_int, c_int, c_int, _v, _v]),
    "mbev_scatter_layernorm_supported": (c_int, [c_int, c_int, c_int, c_int, _v, _v, _v]),
    "mbev_scatter_layernorm_workspace_bytes": (c_int, [c_int, POINTER(c_size_t)]),
    "mbev_scatter_layernorm_forward": (c_int, [_v, _v, _v, c_int, c_int, c_int, c_int, _v, _v, c_float, c_int, _v, _v,
                                               _v, c_size_t, _v]),
    "mbev_scatter_layernorm_backward_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "mbev_scatter_layernorm_backward_workspace_bytes": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "mbev_scatter_layernorm_backward": (c_int, [_v, _v, _v, _v, _v, c_int64, c_int, c_int, c_int, c_int, _v, _v, _v,
                                                _v, _v, _v, c_size_t, _v]),
    "mbev_patch_embed_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "mbev_patch_embed_workspace_bytes": (c_int, [c_int, c_int64, c_int, c_int, POINTER(c_size_t)]),
    "mbev_patch_embed_prepare_weights": (c_int, [_v, c_int, c_int, c_int, _v, _v]),
    "mbev_patch_embed_forward": (c_int, [_v, _v, _v, _v, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, _v, c_float,
                                         _v, _v, _v, _v, _v, c_float, _v, _v, _v, c_size_t, _v]),
    "mbev_encode_batch_workspace_bytes": (c_int, [_G, _P, c_int, c_int64, c_int64, POINTER(c_size_t)]),
    "mbev_encode_batch": (c_int, [_v, POINTER(c_int64), c_int, _G, _P, _v, _v, _v, _v, _v, c_int64, _v, _v, _v,
                                  c_size_t, _v]),
    "mbev_encode_batch_host": (c_int, [_v, _v, POINTER(c_int64), c_int, _G, _P, _v, _v, _v, _v, _v, c_int64, _v,
                                       _v, _v, c_size_t, _v]),
    "mbev_event_create": (c_int, [POINTER(c_void_p)]),
    "mbev_event_destroy": (c_int, [_v]),
    "mbev_encode_batch_pipelined": (c_int, [_v, _v, POINTER(c_int64), c_int, _G, _P, _v, _v, _v, _v, _v, c_int64, _v, _v,
                                            _v, c_size_t, _v, c_size_t, c_int, _v, _v, _v, _v, _v, _v]),
    "mbev_launch_count": (c_int64, []),
}

_lib = None


class MbevError(RuntimeError):
    pass


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load libmask_bev_b200.so, building it in-tree with nvcc if it is missing. Raises if neither works."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise MbevError(f"{LIB_PATH} is missing; run `python -m mask_bev_b200.build` (no CPU fallback exists)")
        from . import build as _build
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise MbevError(f"cannot build {LIB_PATH}: {e}. mask_bev_b200 has no CPU fallback.") from e
    else:
        try:
            from . import build as _build
            if _build.is_stale():
                import warnings
                warnings.warn(f"{LIB_PATH} is older than its sources; rebuild with `python -m mask_bev_b200.build --force`")
        except Exception:  # noqa: BLE001 - a missing source tree next to a shipped library is fine
            pass
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise MbevError(f"{LIB_PATH} does not export {name}; rebuild with `python -m mask_bev_b200.build --force`") from e
        fn.restype = res
        fn.argtypes = args
    if lib.mbev_abi_version() != ABI_VERSION:
        raise MbevError(f"ABI mismatch: library {lib.mbev_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().mbev_status_string(int(status)).decode()
        raise MbevError(f"{what} failed: status {status} ({msg})")


def ptr(t) -> c_void_p:
    """Device (or host) address of a torch tensor, None -> NULL."""
    return c_void_p(None) if t is None else c_void_p(t.data_ptr())


def ptr_array(tensors) -> _PTRS:
    arr = _PTRS()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def launch_count() -> int:
    return int(load().mbev_launch_count())
