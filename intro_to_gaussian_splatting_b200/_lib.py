"""ctypes binding of libgsb_b200.so (C ABI in include/gsb.h).

There is deliberately NO fallback: if the shared library is missing, or no CUDA device is present
when a context is created, the calls raise.  Nothing in this package imports `oracle/`.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# GSB_LIB_PATH: load another build of the same library (A/B runs of kernel variants); never a different backend
LIB_PATH = os.environ.get("GSB_LIB_PATH") or os.path.join(_PKG_DIR, "libgsb_b200.so")

GSB_SEM_REF_CPU = 0
GSB_SEM_REF_CU = 1
GSB_SORT_AUTO = 0
GSB_SORT_FULL = 1
GSB_SORT_SPLIT = 2
GSB_API_VERSION = 2
GSB_NUM_STAGES = 8
STAGE_NAMES = ("project", "depth_sort", "scan", "emit", "sort", "ranges", "composite", "expand")

# every symbol include/gsb.h declares (tests check that the .so exports exactly these)
EXPORTED_SYMBOLS = (
    "gsb_version", "gsb_error_string", "gsb_default_params", "gsb_create", "gsb_destroy", "gsb_upload",
    "gsb_render", "gsb_render_wh", "gsb_render_u8", "gsb_preprocess", "gsb_render_image", "gsb_frame_info",
    "gsb_debug_projection", "gsb_debug_sorted_keys", "gsb_debug_emitted_keys", "gsb_debug_tile_ranges",
    "gsb_stage_times", "gsb_sort_pairs_u64", "gsb_join_host_copies", "gsb_render_backward",
)


class GsbCamera(C.Structure):
    _fields_ = [
        ("world2view", C.c_float * 16),
        ("full_proj", C.c_float * 16),
        ("f_x", C.c_float),
        ("f_y", C.c_float),
        ("tan_fovx", C.c_float),
        ("tan_fovy", C.c_float),
        ("width", C.c_int32),
        ("height", C.c_int32),
    ]


class GsbParams(C.Structure):
    _fields_ = [
        ("tile_size", C.c_int32),
        ("minimum_z", C.c_float),
        ("fov_clamp", C.c_float),
        ("det_min", C.c_float),
        ("lambda_floor", C.c_float),
        ("sigma_extent", C.c_float),
        ("min_weight", C.c_float),
        ("alpha_max", C.c_float),
        ("semantics", C.c_int32),
        ("full_cover", C.c_int32),
        ("sort_mode", C.c_int32),
        ("collect_stage_times", C.c_int32),
        ("async_host_copy", C.c_int32),
        ("save_for_backward", C.c_int32),
        ("cull_alpha", C.c_float),
    ]


class GsbFrameInfo(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("m_in_view", C.c_int64),
        ("k_instances", C.c_int64),
        ("tiles_x", C.c_int32),
        ("tiles_y", C.c_int32),
        ("sort_passes", C.c_int32),
        ("depth_passes", C.c_int32),
        ("kernel_launches", C.c_int32),
        ("key_bits", C.c_int32),
        ("k_sorted", C.c_int64),
        ("frame_id", C.c_int64),
        ("super_w", C.c_int32),
        ("super_h", C.c_int32),
        ("v_with_tiles", C.c_int64),
        ("tail_requeued", C.c_int32),
        ("graph_launch", C.c_int32),
    ]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libgsb_b200.so; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C intro_to_gaussian_splatting_b200/csrc`."
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.gsb_version.restype = C.c_int
    lib.gsb_error_string.restype = C.c_char_p
    lib.gsb_error_string.argtypes = [C.c_int]
    lib.gsb_default_params.restype = None
    lib.gsb_default_params.argtypes = [C.POINTER(GsbParams)]
    lib.gsb_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.gsb_destroy.restype = None
    lib.gsb_destroy.argtypes = [vp]
    lib.gsb_upload.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp]
    for name in ("gsb_render", "gsb_render_wh", "gsb_render_u8"):
        getattr(lib, name).argtypes = [vp, C.POINTER(GsbCamera), C.POINTER(GsbParams), vp, vp]
    lib.gsb_preprocess.argtypes = [vp, C.POINTER(GsbCamera), C.POINTER(GsbParams), C.POINTER(i64)] + [vp] * 13
    lib.gsb_render_image.argtypes = [vp, i32, i32, i32, i64] + [vp] * 8 + [C.POINTER(GsbParams), vp, vp]
    lib.gsb_frame_info.argtypes = [vp, C.POINTER(GsbFrameInfo)]
    lib.gsb_debug_projection.argtypes = [vp] * 7
    lib.gsb_debug_sorted_keys.argtypes = [vp, vp, vp]
    lib.gsb_debug_emitted_keys.argtypes = [vp, vp, vp]
    lib.gsb_debug_tile_ranges.argtypes = [vp, vp]
    lib.gsb_stage_times.argtypes = [vp, C.POINTER(C.c_float * GSB_NUM_STAGES)]
    lib.gsb_sort_pairs_u64.argtypes = [vp, i64, vp, vp, vp, vp, i32, i32, vp]
    lib.gsb_join_host_copies.argtypes = [vp, vp]
    lib.gsb_render_backward.argtypes = [vp, C.POINTER(GsbCamera), C.POINTER(GsbParams), i64] + [vp] * 7
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("gsb_error_string", "gsb_default_params", "gsb_destroy"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def error_string(status: int) -> str:
    return load().gsb_error_string(int(status)).decode()


def check(status: int, what: str = "gsb") -> None:
    """Map the ABI's int status to the error behaviour of the reference op (a RuntimeError,
    like the c10::Error of torch::checkAllSameGPU at splat/c/render.cu:112)."""
    if status != 0:
        raise RuntimeError(f"{what} failed: {error_string(status)} (status {status})")


def default_params(**over) -> GsbParams:
    p = GsbParams()
    load().gsb_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown GsbParams field {k!r}")
        setattr(p, k, v)
    return p
