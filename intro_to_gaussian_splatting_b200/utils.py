"""Camera conventions of the reference, frozen (splat/utils.py).

These run on the HOST, once per view, and produce the fp32 constants that cross the C ABI in
GsbCamera.  They are written with the same torch/`math` operations in the same order as the
reference so the constants are bit-identical (tests/test_camera_golden.py checks them against
tensors produced by the reference's own GaussianImage).  Everything per-Gaussian that the
reference computes in this file (in_view_frustum, compute_2d_covariance, ...) lives in
csrc/project.cu instead.
"""

from __future__ import annotations

import math

import torch

from .colmap_io import read_camera_file, read_image_file  # noqa: F401  (same import surface as splat.utils)
from .ply_io import fetchPly, storePly  # noqa: F401


def inverse_sigmoid(x: torch.Tensor) -> torch.Tensor:
    """logit, splat/utils.py:128-129."""
    return torch.log(x / (1 - x))


def build_rotation(r: torch.Tensor) -> torch.Tensor:
    """(N,4) wxyz quaternions -> (N,3,3), normalising first (splat/utils.py:132-155)."""
    n = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / n[:, None]
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rot = torch.zeros((q.size(0), 3, 3), device=r.device, dtype=r.dtype)
    rot[:, 0, 0] = 1 - 2 * (y * y + z * z)
    rot[:, 0, 1] = 2 * (x * y - w * z)
    rot[:, 0, 2] = 2 * (x * z + w * y)
    rot[:, 1, 0] = 2 * (x * y + w * z)
    rot[:, 1, 1] = 1 - 2 * (x * x + z * z)
    rot[:, 1, 2] = 2 * (y * z - w * x)
    rot[:, 2, 0] = 2 * (x * z - w * y)
    rot[:, 2, 1] = 2 * (y * z + w * x)
    rot[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return rot


def focal2fov(focal: torch.Tensor, pixels: torch.Tensor) -> torch.Tensor:
    """2*atan(pixels / (2 focal)): the ratio in fp32, atan in double, result fp32 (splat/utils.py:158-159)."""
    return torch.Tensor([2 * math.atan(pixels / (2 * focal))])


def getWorld2View(R: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """[[R, t], [0, 1]] (splat/utils.py:162-172); GaussianImage transposes it to row-vector form."""
    m = torch.zeros((4, 4))
    m[:3, :3] = R
    m[:3, 3] = t
    m[3, 3] = 1.0
    return m.float()


def getProjectionMatrix(znear: torch.Tensor, zfar: torch.Tensor, fovX: torch.Tensor, fovY: torch.Tensor) -> torch.Tensor:
    """Perspective matrix of the 3DGS code base as the reference uses it (splat/utils.py:189-225)."""
    tan_y = math.tan((fovY / 2))
    tan_x = math.tan((fovX / 2))
    top = tan_y * znear
    bottom = -top
    right = tan_x * znear
    left = -right
    P = torch.zeros(4, 4)
    z_sign = 1.0
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = z_sign
    P[2, 2] = z_sign * zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def ndc2Pix(points: torch.Tensor, dimension) -> torch.Tensor:
    """(v + 1) * (S - 1) * 0.5 (splat/utils.py:313-317)."""
    return (points + 1) * (dimension - 1) * 0.5


# ---- The functions below are the per-Gaussian conventions of splat/utils.py that the frozen scene API still
# exposes through public methods (GaussianImage.project_point_to_camera_perspective_projection,
# GaussianScene.get_2d_covariance).  They are torch restatements for callers of those methods -- debugging helpers in
# the reference too (splat/gaussian_scene.py:44-51) -- and deliberately keep the reference's operator order, because
# that order IS the contract.  The render path does not call them: csrc/project.cu does this
# arithmetic for every frame.
def get_intrinsic_matrix(f_x, f_y, c_x, c_y) -> torch.Tensor:
    """3x4 pinhole matrix [[fx 0 cx 0] [0 fy cy 0] [0 0 1 0]] (splat/utils.py:19-37)."""
    k = torch.zeros((3, 4))
    k[0, 0], k[0, 2], k[1, 1], k[1, 2], k[2, 2] = float(f_x), float(c_x), float(f_y), float(c_y), 1.0
    return k


def get_extrinsic_matrix(R: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """4x4 [[R t] [0 1]] (splat/utils.py:40-52)."""
    return getWorld2View(R, t)


def in_view_frustum(points: torch.Tensor, view_matrix: torch.Tensor, minimum_z: float = 0.2) -> torch.Tensor:
    """z of [p 1] @ view_matrix >= minimum_z: the only cull of the reference (splat/utils.py:293-310)."""
    hom = torch.cat([points, torch.ones((points.shape[0], 1), device=points.device, dtype=points.dtype)], dim=1)
    return (hom @ view_matrix)[:, 2] >= minimum_z


def compute_2d_covariance(points: torch.Tensor, extrinsic_matrix: torch.Tensor, covariance_3d: torch.Tensor,
                          tan_fovY: torch.Tensor, tan_fovX: torch.Tensor, focal_x: torch.Tensor,
                          focal_y: torch.Tensor) -> torch.Tensor:
    """EWA projection J W Sigma W^T J^T [:2,:2] with x/z, y/z clamped to +-1.3 tan(fov/2) and NO low-pass term
    (splat/utils.py:320-354; argument order as there: tan_fovY before tan_fovX)."""
    hom = torch.cat([points, torch.ones((points.shape[0], 1), device=points.device, dtype=points.dtype)], dim=1)
    v = (hom @ extrinsic_matrix)[:, :3]
    z = v[:, 2]
    x = torch.clamp(v[:, 0] / z, -1.3 * tan_fovX, 1.3 * tan_fovX) * z
    y = torch.clamp(v[:, 1] / z, -1.3 * tan_fovY, 1.3 * tan_fovY) * z
    J = torch.zeros((v.shape[0], 3, 3), device=covariance_3d.device)
    J[:, 0, 0] = focal_x / z
    J[:, 0, 2] = -(focal_x * x) / (z**2)
    J[:, 1, 1] = focal_y / z
    J[:, 1, 2] = -(focal_y * y) / (z**2)
    W = extrinsic_matrix[:3, :3].T
    return (J @ W @ covariance_3d @ W.T @ J.transpose(1, 2))[:, :2, :2]
