"""Per-view camera constants: the host-side input provider of the render path (splat/image.py:18-70)."""

from __future__ import annotations

import numpy as np
import torch

from ._lib import GsbCamera
from .colmap_io import Camera, Image
from .utils import build_rotation, focal2fov, getProjectionMatrix, getWorld2View


class GaussianImage(torch.nn.Module):
    """Same attributes as the reference class (1-element / 4x4 fp32 tensors on `self.device`).

    Only PINHOLE-style params[0..3] = fx, fy, cx, cy are read, as in splat/image.py:28-31."""

    def __init__(self, camera: Camera, image: Image) -> None:
        super().__init__()
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        dev = self.device
        self.f_x = torch.Tensor([camera.params[0]]).to(dev)
        self.f_y = torch.Tensor([camera.params[1]]).to(dev)
        self.c_x = torch.Tensor([camera.params[2]]).to(dev)
        self.c_y = torch.Tensor([camera.params[3]]).to(dev)
        self.R = build_rotation(torch.Tensor(image.qvec).unsqueeze(0)).to(dev)
        self.T = torch.Tensor(image.tvec).to(dev)
        self.height = torch.Tensor([camera.height]).to(dev)
        self.width = torch.Tensor([camera.width]).to(dev)
        self.fovX = focal2fov(self.f_x, self.width).to(dev)
        self.fovY = focal2fov(self.f_y, self.height).to(dev)
        self.tan_fovX = torch.tan(self.fovX / 2).to(dev)
        self.tan_fovY = torch.tan(self.fovY / 2).to(dev)
        self.zfar = torch.Tensor([100.0]).to(dev)   # splat/image.py:46-47
        self.znear = torch.Tensor([0.001]).to(dev)
        self.name = image.name
        # row-vector convention: row = [x y z 1] @ M   (splat/image.py:51-65)
        self.world2view = getWorld2View(R=self.R[0], t=self.T).transpose(0, 1).to(dev)
        self.projection_matrix = (
            getProjectionMatrix(znear=self.znear, zfar=self.zfar, fovX=self.fovX, fovY=self.fovY).transpose(0, 1).to(dev)
        )
        self.full_proj_transform = (
            self.world2view.unsqueeze(0).bmm(self.projection_matrix.unsqueeze(0)).squeeze(0).to(dev)
        )
        self.camera_center = self.world2view.inverse()[3, :3].to(dev)
        self._packed = None

    def pack(self) -> GsbCamera:
        """The GsbCamera struct that crosses the C ABI: the tensors above, bit for bit."""
        if self._packed is None:
            cam = GsbCamera()
            w2v = self.world2view.detach().cpu().contiguous().numpy().astype(np.float32).reshape(16)
            fpt = self.full_proj_transform.detach().cpu().contiguous().numpy().astype(np.float32).reshape(16)
            for i in range(16):
                cam.world2view[i] = float(w2v[i])
                cam.full_proj[i] = float(fpt[i])
            cam.f_x = float(self.f_x.item())
            cam.f_y = float(self.f_y.item())
            cam.tan_fovx = float(self.tan_fovX.item())
            cam.tan_fovy = float(self.tan_fovY.item())
            cam.width = int(self.width.item())
            cam.height = int(self.height.item())
            self._packed = cam
        return self._packed
