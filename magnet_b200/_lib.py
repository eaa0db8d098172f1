"""ctypes binding of libmagnet_b200.so (the C ABI in include/magnet_b200.h).

There is NO fallback: if the shared library is missing or does not export the ABI, importing
the ops raises.  Device pointers are passed as integers (``tensor.data_ptr()``), the stream as
``torch.cuda.current_stream().cuda_stream``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

PKG = Path(__file__).resolve().parent
# MAGNET_B200_LIB selects a tuning build (magnet_b200.build.build(defines=..., tag=...)); default = production
LIB_PATH = Path(os.environ["MAGNET_B200_LIB"]) if os.environ.get("MAGNET_B200_LIB") else PKG / "libmagnet_b200.so"

MAGNET_ABI_VERSION = 3
MAGNET_MAX_PLANES = 256

OK, ERR_NULL, ERR_SHAPE, ERR_UNSUPPORTED, ERR_CUDA, ERR_ALIGN = 0, -1, -2, -3, -4, -5
DEPTH_VOLUME, DEPTH_GAUSS, DEPTH_PLANES = 0, 1, 2
SRC_NCHW, SRC_TILED32, SRC_PIXC, SRC_SPLIT16 = 0, 1, 2, 3
VARIANT_AUTO, VARIANT_DIRECT, VARIANT_CELLS, VARIANT_CELLS_NOREUSE, VARIANT_TMA, VARIANT_MMA = 0, 1, 2, 3, 4, 5

# every symbol include/magnet_b200.h declares (tests check the library exports all of them)
EXPORTS = (
    "magnet_abi_version", "magnet_strerror", "magnet_last_cuda_error", "magnet_launch_count",
    "magnet_cost_launch_info", "magnet_cost_volume_f32", "magnet_cost_volume_f_bwd_f32", "magnet_pack_cameras_f32",
    "magnet_repack_tiled32_f32", "magnet_repack_pixc_f32", "magnet_repack_split16_f32", "magnet_split16_bytes", "magnet_sample_depths_f32", "magnet_gaussian_update_fwd_f32",
    "magnet_gaussian_update_bwd_f32", "magnet_convex_upsample_fwd_f32", "magnet_convex_upsample_bwd_f32",
    "magnet_relative_poses_f32", "magnet_camera_rays_f32",
    "magnet_upsample_nll_partials", "magnet_upsample_nll_fwd_f32", "magnet_upsample_nll_bwd_f32",
)


class CostArgs(C.Structure):
    """Mirror of ``struct magnet_cost_args``."""
    _fields_ = [
        ("B", C.c_int32), ("V", C.c_int32), ("D", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("depth_mode", C.c_int32), ("src_layout", C.c_int32), ("consistency", C.c_int32), ("softmax", C.c_int32),
        ("variant", C.c_int32), ("kappa", C.c_float),
        ("ref_feat", C.c_void_p), ("src_feat", C.c_void_p), ("src_gmm", C.c_void_p), ("rays", C.c_void_p),
        ("cams", C.c_void_p), ("d_volume", C.c_void_p), ("ref_gmm", C.c_void_p), ("k_host", C.c_void_p),
        ("out", C.c_void_p),
    ]


class CostFBwdArgs(C.Structure):
    """Mirror of ``struct magnet_cost_f_bwd_args``."""
    _fields_ = [("fwd", C.POINTER(CostArgs)), ("prob", C.c_void_p), ("grad_out", C.c_void_p), ("workspace", C.c_void_p),
                ("grad_ref", C.c_void_p), ("grad_src", C.c_void_p)]


class MagnetError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load (once) and type the shared library.  Raises if it is missing — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MagnetError(
            f"{LIB_PATH} not found: build it with `python -m magnet_b200.build` (or __graft_entry__.build()). "
            "magnet_b200 has no CPU / PyTorch fallback for the matching path.")
    L = C.CDLL(str(LIB_PATH))
    for name in EXPORTS:
        if not hasattr(L, name):
            raise MagnetError(f"{LIB_PATH} does not export {name}")
    L.magnet_abi_version.restype = C.c_int
    L.magnet_strerror.restype = C.c_char_p
    L.magnet_strerror.argtypes = [C.c_int]
    L.magnet_last_cuda_error.restype = C.c_char_p
    L.magnet_launch_count.restype = C.c_uint64
    L.magnet_cost_launch_info.restype = C.c_int
    L.magnet_cost_launch_info.argtypes = [C.POINTER(CostArgs), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.magnet_cost_volume_f32.restype = C.c_int
    L.magnet_cost_volume_f32.argtypes = [C.POINTER(CostArgs), C.c_void_p]
    L.magnet_cost_volume_f_bwd_f32.restype = C.c_int
    L.magnet_cost_volume_f_bwd_f32.argtypes = [C.POINTER(CostFBwdArgs), C.c_void_p]
    L.magnet_pack_cameras_f32.restype = C.c_int
    L.magnet_pack_cameras_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                          C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int32,
                                          C.c_int32, C.c_void_p, C.c_void_p]
    L.magnet_repack_tiled32_f32.restype = C.c_int
    L.magnet_repack_tiled32_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.magnet_repack_pixc_f32.restype = C.c_int
    L.magnet_repack_pixc_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_void_p]
    L.magnet_split16_bytes.restype = C.c_size_t
    L.magnet_split16_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.magnet_repack_split16_f32.restype = C.c_int
    L.magnet_repack_split16_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_void_p]
    L.magnet_sample_depths_f32.restype = C.c_int
    L.magnet_sample_depths_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.magnet_gaussian_update_fwd_f32.restype = C.c_int
    L.magnet_gaussian_update_fwd_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.magnet_gaussian_update_bwd_f32.restype = C.c_int
    L.magnet_gaussian_update_bwd_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.magnet_convex_upsample_fwd_f32.restype = C.c_int
    L.magnet_convex_upsample_fwd_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                 C.c_int32, C.c_void_p, C.c_void_p]
    L.magnet_convex_upsample_bwd_f32.restype = C.c_int
    L.magnet_convex_upsample_bwd_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                                 C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.magnet_relative_poses_f32.restype = C.c_int
    L.magnet_relative_poses_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.magnet_camera_rays_f32.restype = C.c_int
    L.magnet_camera_rays_f32.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.magnet_upsample_nll_partials.restype = C.c_int
    L.magnet_upsample_nll_partials.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    L.magnet_upsample_nll_fwd_f32.restype = C.c_int
    L.magnet_upsample_nll_fwd_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                              C.c_int32, C.c_void_p, C.c_void_p]
    L.magnet_upsample_nll_bwd_f32.restype = C.c_int
    L.magnet_upsample_nll_bwd_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int32,
                                              C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    if L.magnet_abi_version() != MAGNET_ABI_VERSION:
        raise MagnetError(f"ABI version mismatch: library {L.magnet_abi_version()} != binding {MAGNET_ABI_VERSION}")
    _lib = L
    return L


def check(status: int, what: str) -> None:
    """Status -> RuntimeError (the reference's operators raise plain Python exceptions)."""
    if status == OK:
        return
    L = lib()
    msg = L.magnet_strerror(status).decode()
    if status == ERR_CUDA:
        msg += ": " + L.magnet_last_cuda_error().decode()
    raise MagnetError(f"{what} failed ({status}): {msg}")


def launch_count() -> int:
    return int(lib().magnet_launch_count())
