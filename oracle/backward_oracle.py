"""backward_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

float64 torch restatement of the reference's forward render path, written so that torch autograd can
differentiate it; it is the checker of gsb_render_backward.

  projection   GaussianScene.preprocess            splat/gaussian_scene.py:70-144
               Gaussians.get_3d_covariance_matrix  splat/gaussians.py:54-69, build_rotation splat/utils.py:132-155
               compute_2d_covariance               splat/utils.py:320-354
               compute_inverted_covariance         splat/utils.py:368-393
  compositing  render_pixel                        splat/gaussian_scene.py:146-171
               compute_gaussian_weight             splat/utils.py:357-365

PINNING.  The reference has NO backward pass (README.md:3 announces training; compute_gaussian_weight returns
`.item()`, which cuts its autograd graph), so there is nothing of the reference's to pin the compositing gradient
against: **parity unpinned** for that half, and DESIGN.md says so.  What IS pinned:
  * this restatement's forward image against oracle/gs_oracle.c (itself pinned bit-exact / 1e-6 to the reference)
    -- tests/test_backward_oracle.py;
  * the projection half of the gradient against the REFERENCE'S OWN autograd through GaussianScene.preprocess
    (which is differentiable torch code) -- fixtures tests/golden/preprocess_grad_*.npz made by
    tests/golden/make_golden.py with the unmodified reference.
Tile membership, the depth order inside each tile and the early termination are taken as constants (from the C
oracle's forward / decided under no_grad), exactly what the CUDA backward assumes.

Only tests/ and __graft_entry__.smoke() may import this module.
"""

from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch

DT = torch.float64


def _cam_tensors(cam):
    V = torch.tensor(list(cam.world2view), dtype=DT).reshape(4, 4)
    P = torch.tensor(list(cam.full_proj), dtype=DT).reshape(4, 4)
    return V, P


def build_rotation(q: torch.Tensor) -> torch.Tensor:
    """splat/utils.py:132-155 (normalises its argument once more)."""
    n = torch.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])
    q = q / n[:, None]
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rows = [
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y),
    ]
    return torch.stack(rows, dim=1).reshape(-1, 3, 3)


def project(cam, prm, xyz, scales, quats, opacity_logit) -> Dict[str, torch.Tensor]:
    """Differentiable per-Gaussian quantities for ALL rows (callers mask by the C oracle's in_view)."""
    V, P = _cam_tensors(cam)
    n = xyz.shape[0]
    hom = torch.cat([xyz, torch.ones(n, 1, dtype=DT)], dim=1)
    view = hom @ V
    clip = hom @ P
    ndc = clip[:, :2] / clip[:, 3:4]
    px = (ndc[:, 0] + 1.0) * (cam.width - 1) * 0.5  # ndc2Pix, splat/utils.py:313-317
    py = (ndc[:, 1] + 1.0) * (cam.height - 1) * 0.5
    qn = torch.nn.functional.normalize(quats, p=2, dim=1)
    R = build_rotation(qn)
    M = R * scales[:, None, :]  # R @ diag(s)
    cov3 = M @ M.transpose(1, 2)
    limx = prm.fov_clamp * cam.tan_fovx
    limy = prm.fov_clamp * cam.tan_fovy
    z = view[:, 2]
    x = torch.clamp(view[:, 0] / z, -limx, limx) * z
    y = torch.clamp(view[:, 1] / z, -limy, limy) * z
    J = torch.zeros(n, 3, 3, dtype=DT)
    J[:, 0, 0] = cam.f_x / z
    J[:, 0, 2] = -(cam.f_x * x) / (z ** 2)
    J[:, 1, 1] = cam.f_y / z
    J[:, 1, 2] = -(cam.f_y * y) / (z ** 2)
    W = V[:3, :3].T
    cov2 = (J @ W @ cov3 @ W.T @ J.transpose(1, 2))[:, :2, :2]
    det = cov2[:, 0, 0] * cov2[:, 1, 1] - cov2[:, 0, 1] * cov2[:, 1, 0]
    det = torch.clamp(det, min=float(prm.det_min))
    inv = torch.stack([cov2[:, 1, 1] / det, -cov2[:, 0, 1] / det, -cov2[:, 1, 0] / det, cov2[:, 0, 0] / det], dim=1)
    op1 = torch.sigmoid(opacity_logit.reshape(-1))  # PreprocessedScene.sigmoid_opacity, gaussian_scene.py:143
    op2 = torch.sigmoid(op1)                        # render_pixel applies sigmoid AGAIN, gaussian_scene.py:164
    return dict(px=px, py=py, inv=inv, op1=op1, op2=op2, depth=z, cov2=cov2)


def render(cam, prm, xyz, scales, quats, colors, opacity_logit, ranges: np.ndarray, payload: np.ndarray,
           ntx: int, nty: int) -> torch.Tensor:
    """(H,W,3) float64 image; `ranges`/`payload` are the C oracle's per-tile depth-ordered lists."""
    # project only the rows that appear in some tile list: culled rows (z < minimum_z, possibly z <= 0) would put
    # inf/NaN into the graph, and they receive no gradient anyway
    k_total = int(ranges[:, 1].max()) if len(ranges) else 0
    used = np.unique(np.asarray(payload[:k_total], dtype=np.int64))
    local = {int(g): j for j, g in enumerate(used)}
    sel = torch.as_tensor(used, dtype=torch.long)
    pr = project(cam, prm, xyz[sel], scales[sel], quats[sel], opacity_logit.reshape(-1)[sel])
    colors_u = colors[sel]
    H, W, T = cam.height, cam.width, prm.tile_size
    minw = float(prm.min_weight)
    image = torch.zeros(H, W, 3, dtype=DT)
    for tile in range(ntx * nty):
        s, e = int(ranges[tile, 0]), int(ranges[tile, 1])
        if e <= s:
            continue
        x0, y0 = (tile % ntx) * T, (tile // ntx) * T
        x1, y1 = min(x0 + T, W), min(y0 + T, H)
        ys, xs = torch.meshgrid(torch.arange(y0, y1, dtype=DT), torch.arange(x0, x1, dtype=DT), indexing="ij")
        Tw = torch.ones_like(xs)
        live = torch.ones_like(xs, dtype=torch.bool)
        col = torch.zeros(*xs.shape, 3, dtype=DT)
        for g in (local[int(v)] for v in payload[s:e]):
            dx = pr["px"][g] - xs
            dy = pr["py"][g] - ys
            i = pr["inv"][g]
            power = -0.5 * (i[0] * dx * dx + (i[1] + i[2]) * dx * dy + i[3] * dy * dy)
            alpha = torch.exp(power) * pr["op2"][g]
            test = Tw * (1.0 - alpha)
            with torch.no_grad():
                upd = live & ~(test < minw)  # the Gaussian that trips the threshold is NOT added
            col = col + torch.where(upd, Tw * alpha, torch.zeros_like(alpha))[..., None] * colors_u[g]
            Tw = torch.where(upd, test, Tw)
            live = upd
            if not bool(live.any()):
                break
        image[y0:y1, x0:x1] = col
    return image


def gradients(cam, prm, arrays: Sequence[np.ndarray], grad_image: np.ndarray, ranges, payload, ntx, nty):
    """-> (image float64 (H,W,3), dict of float64 gradient arrays named like the reference's attributes)."""
    names = ("points", "scales", "quaternions", "colors", "opacity")
    ts = [torch.tensor(np.asarray(a, dtype=np.float64), dtype=DT, requires_grad=True) for a in arrays]
    img = render(cam, prm, ts[0], ts[1], ts[2], ts[3], ts[4], ranges, payload, ntx, nty)
    loss = (img * torch.tensor(np.asarray(grad_image, dtype=np.float64), dtype=DT)).sum()
    loss.backward()
    grads = {k: (t.grad if t.grad is not None else torch.zeros_like(t)).numpy() for k, t in zip(names, ts)}
    return img.detach().numpy(), grads
