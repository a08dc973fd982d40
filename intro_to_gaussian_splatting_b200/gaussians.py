"""Gaussian parameter container (splat/gaussians.py:9-69): same constructor, same attributes."""

from __future__ import annotations

import os

import torch
from torch import nn

from .utils import build_rotation, inverse_sigmoid


class Gaussians(nn.Module):
    """points (N,3); colors (N,3) stored as rgb/256; scales (N,3) LINEAR, default 0.001;
    quaternions (N,4) wxyz, default identity; opacity (N,1) logit, default logit(0.9999)
    (splat/gaussians.py:19-33).  Overwrite the attributes to load a trained / synthetic set.

    The reference ctor also writes `<model_path>/point_cloud.ply` through plyfile and a Python tuple loop
    (splat/gaussians.py:17-18, splat/utils.py:102-125); nothing on the render path reads it back, so here it is
    opt-in (`write_ply=True`, ply_io.storePly: same vertex layout, one numpy structured array)."""

    def __init__(self, points: torch.Tensor, colors: torch.Tensor, model_path: str = ".", write_ply: bool = False) -> None:
        super().__init__()
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.point_cloud_path = os.path.join(model_path, "point_cloud.ply")
        if write_ply:  # the reference always does this (splat/gaussians.py:17-18); here it is opt-in and fast
            from .ply_io import storePly

            storePly(self.point_cloud_path, points.detach().cpu().numpy(), colors.detach().cpu().numpy())
        self.points = points.clone().requires_grad_(True).to(self.device).float()
        self.colors = (colors / 256).clone().requires_grad_(True).to(self.device).float()
        self.scales = torch.ones((len(self.points), 3)).to(self.device).float() * 0.001
        self.quaternions = torch.zeros((len(self.points), 4)).to(self.device)
        self.quaternions[:, 0] = 1.0
        self.opacity = inverse_sigmoid(0.9999 * torch.ones((self.points.shape[0], 1), dtype=torch.float)).to(self.device)

    @classmethod
    def from_ply(cls, path: str, model_path: str = ".") -> "Gaussians":
        """A complete Gaussian set from a PLY: this package's native record or the 3DGS trainers' layout
        (ply_io.load_gaussians).  What the reference cannot do: its loader restores positions and colours only
        (splat/utils.py:93-99) and the constructor hard-codes the rest (splat/gaussians.py:23-33)."""
        from .ply_io import load_gaussians

        a = load_gaussians(path)
        g = cls(points=torch.from_numpy(a["points"]), colors=torch.from_numpy(a["colors"]) * 256, model_path=model_path)
        g.scales = torch.from_numpy(a["scales"]).to(g.device).float()
        g.quaternions = torch.from_numpy(a["quaternions"]).to(g.device).float()
        g.opacity = torch.from_numpy(a["opacity"]).to(g.device).float()
        return g

    def save_ply(self, path: str, layout: str = "native") -> None:
        """Persist all five attribute tensors (ply_io.save_gaussians)."""
        from .ply_io import save_gaussians

        save_gaussians(path, self.points, self.scales, self.quaternions, self.colors, self.opacity, layout=layout)

    def get_3d_covariance_matrix(self) -> torch.Tensor:
        """R S S^T R^T per Gaussian (splat/gaussians.py:54-69).  Kept for API parity; the render
        path computes this inside csrc/project.cu and never calls it."""
        q = nn.functional.normalize(self.quaternions, p=2, dim=1)
        rot = build_rotation(q)
        s = torch.zeros((len(self.points), 3, 3)).to(self.device)
        s[:, 0, 0] = self.scales[:, 0]
        s[:, 1, 1] = self.scales[:, 1]
        s[:, 2, 2] = self.scales[:, 2]
        m = rot @ s
        return m @ m.transpose(1, 2)
