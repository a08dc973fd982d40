"""TEST INFRASTRUCTURE -- numpy/ctypes front-end of the C oracle (oracle/gs_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.  See gs_oracle.c for what each function
restates (reference file:line) and how the restatement is pinned.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgs_oracle.so")


class Camera(C.Structure):
    """Mirror of GsbCamera (include/gsb.h)."""

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


class Params(C.Structure):
    """Mirror of GsbParams (include/gsb.h)."""

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


def default_params(**over) -> Params:
    p = Params(16, 0.2, 1.3, 1e-3, 0.1, 3.0, 1e-6, 0.99, 0, 0, 0, 0, 0, 0, 2.0 ** -30)
    for k, v in over.items():
        setattr(p, k, v)
    return p


def make_camera(world2view, full_proj, f_x, f_y, tan_fovx, tan_fovy, width, height) -> Camera:
    cam = Camera()
    w = np.ascontiguousarray(np.asarray(world2view, dtype=np.float32)).reshape(16)
    f = np.ascontiguousarray(np.asarray(full_proj, dtype=np.float32)).reshape(16)
    for i in range(16):
        cam.world2view[i] = float(w[i])
        cam.full_proj[i] = float(f[i])
    cam.f_x = float(np.float32(f_x))
    cam.f_y = float(np.float32(f_y))
    cam.tan_fovx = float(np.float32(tan_fovx))
    cam.tan_fovy = float(np.float32(tan_fovy))
    cam.width = int(width)
    cam.height = int(height)
    return cam


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_project.restype = C.c_int64
        _lib.orc_depth_order.restype = C.c_int64
        _lib.orc_emit_keys.restype = C.c_int64
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


@dataclass
class Projection:
    """Per-Gaussian records in Gaussian-index order (rows with in_view == 0 are zero)."""

    m: int
    in_view: np.ndarray
    depth: np.ndarray
    pxy: np.ndarray
    cov2d: np.ndarray
    conic: np.ndarray
    radius: np.ndarray
    bbox: np.ndarray  # (n,4) min_x, min_y, max_x, max_y
    sig_op: np.ndarray
    rect: np.ndarray  # (n,4) tx0, tx1, ty0, ty1
    count: np.ndarray
    colors: np.ndarray


def grid(cam: Camera, prm: Params):
    ntx, nty = C.c_int32(), C.c_int32()
    lib().orc_grid(C.byref(cam), C.byref(prm), C.byref(ntx), C.byref(nty))
    return ntx.value, nty.value


def project(cam: Camera, prm: Params, xyz, scales, quats, colors, opacity_logit) -> Projection:
    xyz, scales, quats, colors, op = map(_f32, (xyz, scales, quats, colors, opacity_logit))
    n = xyz.shape[0]
    out = Projection(
        0,
        np.zeros(n, np.uint8), np.zeros(n, np.float32), np.zeros((n, 2), np.float32),
        np.zeros((n, 2, 2), np.float32), np.zeros((n, 2, 2), np.float32), np.zeros(n, np.float32),
        np.zeros((n, 4), np.float32), np.zeros(n, np.float32), np.zeros((n, 4), np.int32),
        np.zeros(n, np.uint32), colors,
    )
    out.m = int(lib().orc_project(
        C.byref(cam), C.byref(prm), C.c_int64(n), _p(xyz), _p(scales), _p(quats), _p(colors), _p(op),
        _p(out.in_view), _p(out.depth), _p(out.pxy), _p(out.cov2d), _p(out.conic), _p(out.radius),
        _p(out.bbox), _p(out.sig_op), _p(out.rect), _p(out.count)))
    return out


def depth_order(pr: Projection) -> np.ndarray:
    n = pr.in_view.shape[0]
    order = np.zeros(max(n, 1), np.int32)
    m = int(lib().orc_depth_order(C.c_int64(n), _p(pr.in_view), _p(pr.depth), _p(order)))
    return order[:m].copy()


def preprocess(cam: Camera, prm: Params, xyz, scales, quats, colors, opacity_logit) -> dict:
    """The 12 PreprocessedScene fields (splat/schema.py:13-25), depth-sorted, + source_index."""
    pr = project(cam, prm, xyz, scales, quats, colors, opacity_logit)
    o = depth_order(pr)
    pts = pr.pxy[o]
    return dict(
        points=pts, colors=pr.colors[o], covariance_2d=pr.cov2d[o], depths=pr.depth[o],
        inverse_covariance_2d=pr.conic[o], radius=pr.radius[o], points_xy=pts.copy(),
        min_x=pr.bbox[o, 0], min_y=pr.bbox[o, 1], max_x=pr.bbox[o, 2], max_y=pr.bbox[o, 3],
        sigmoid_opacity=pr.sig_op[o].reshape(-1, 1), source_index=o,
    )


def emit_keys(pr: Projection, ntx: int):
    n = pr.in_view.shape[0]
    k = int(lib().orc_emit_keys(C.c_int64(n), _p(pr.in_view), _p(pr.depth), _p(pr.rect), C.c_int32(ntx), None, None))
    keys = np.zeros(max(k, 1), np.uint64)
    payload = np.zeros(max(k, 1), np.uint32)
    lib().orc_emit_keys(C.c_int64(n), _p(pr.in_view), _p(pr.depth), _p(pr.rect), C.c_int32(ntx), _p(keys), _p(payload))
    return keys[:k], payload[:k]


def sort_pairs(keys: np.ndarray, payload: np.ndarray):
    keys = np.ascontiguousarray(keys, np.uint64)
    payload = np.ascontiguousarray(payload, np.uint32)
    ko, vo = np.empty_like(keys), np.empty_like(payload)
    if keys.shape[0]:
        lib().orc_sort_pairs(C.c_int64(keys.shape[0]), _p(keys), _p(payload), _p(ko), _p(vo))
    return ko, vo


def tile_ranges(sorted_keys: np.ndarray, ntiles: int) -> np.ndarray:
    r = np.zeros((max(ntiles, 1), 2), np.uint32)
    sk = np.ascontiguousarray(sorted_keys, np.uint64)
    lib().orc_tile_ranges(C.c_int64(sk.shape[0]), _p(sk), C.c_int64(ntiles), _p(r))
    return r[:ntiles]


def composite(cam: Camera, prm: Params, ranges, payload, pr: Projection, steps_per_pixel=None):
    """REF_CPU compositing -> ((H,W,3) image, executed steps)."""
    img = np.zeros((cam.height, cam.width, 3), np.float32)
    steps = C.c_int64(0)
    ranges = np.ascontiguousarray(ranges, np.uint32)
    payload = np.ascontiguousarray(payload, np.uint32)
    lib().orc_composite(C.byref(cam), C.byref(prm), _p(ranges), _p(payload), _p(pr.pxy), _p(pr.conic),
                        _p(pr.colors), _p(pr.sig_op), _p(img), C.byref(steps), _p(steps_per_pixel))
    return img, steps.value


@dataclass
class Frame:
    proj: Projection
    keys: np.ndarray
    payload: np.ndarray
    sorted_keys: np.ndarray
    sorted_payload: np.ndarray
    ranges: np.ndarray
    image: Optional[np.ndarray]
    steps: int
    ntx: int
    nty: int


def render(cam: Camera, prm: Params, xyz, scales, quats, colors, opacity_logit, with_image: bool = True) -> Frame:
    """Whole path on the CPU: projection -> keys -> stable sort -> ranges -> compositing."""
    ntx, nty = grid(cam, prm)
    pr = project(cam, prm, xyz, scales, quats, colors, opacity_logit)
    keys, payload = emit_keys(pr, ntx)
    sk, sp = sort_pairs(keys, payload)
    rng = tile_ranges(sk, ntx * nty)
    img, steps = (None, 0)
    if with_image:
        img, steps = composite(cam, prm, rng, sp, pr)
    return Frame(pr, keys, payload, sk, sp, rng, img, steps, ntx, nty)


def composite_cu(width, height, prm: Params, ntx, nty, ranges, payload, means, conic, colors, opacity,
                 min_x, max_x, min_y, max_y):
    img = np.zeros((height, width, 3), np.float32)
    args = [np.ascontiguousarray(a, np.float32) for a in (means, conic, colors, opacity, min_x, max_x, min_y, max_y)]
    ranges = np.ascontiguousarray(ranges, np.uint32)
    payload = np.ascontiguousarray(payload, np.uint32)
    lib().orc_composite_cu(C.c_int32(width), C.c_int32(height), C.byref(prm), C.c_int32(ntx), C.c_int32(nty),
                           _p(ranges), _p(payload), *[_p(a) for a in args], _p(img))
    return img


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(n))
