"""Per-view camera constants: the host-side input provider of the render path (splat/image.py:18-70)."""

from __future__ import annotations

import numpy as np
import torch

from ._lib import GsbCamera
from .colmap_io import Camera, Image
from .utils import (build_rotation, focal2fov, get_extrinsic_matrix, get_intrinsic_matrix, getProjectionMatrix, getWorld2View,
                    in_view_frustum, ndc2Pix)


class GaussianImage(torch.nn.Module):
    """Same attributes and method as the reference class (1-element / 3x4 / 4x4 fp32 tensors on `self.device`).

    Only PINHOLE-style params[0..3] = fx, fy, cx, cy are read, as in splat/image.py:28-31."""

    def __init__(self, camera: Camera, image: Image) -> None:
        super().__init__()
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        # Every constant is computed ON THE CPU with the reference's op sequence and only then moved to
        # `self.device`.  The reference evaluates tan / bmm on whatever device it runs on, and CUDA's tan and
        # cuBLAS' 4x4 product differ from the CPU's in the last bit; the parity target is the CPU path, so the
        # CPU values are the contract (a 1-ulp tan_fovX moves the fov clamp of compute_2d_covariance).
        f_x = torch.Tensor([camera.params[0]])
        f_y = torch.Tensor([camera.params[1]])
        c_x = torch.Tensor([camera.params[2]])
        c_y = torch.Tensor([camera.params[3]])
        R = build_rotation(torch.Tensor(image.qvec).unsqueeze(0))
        T = torch.Tensor(image.tvec)
        height = torch.Tensor([camera.height])
        width = torch.Tensor([camera.width])
        fovX = focal2fov(f_x, width)
        fovY = focal2fov(f_y, height)
        tan_fovX = torch.tan(fovX / 2)
        tan_fovY = torch.tan(fovY / 2)
        zfar = torch.Tensor([100.0])   # splat/image.py:46-47
        znear = torch.Tensor([0.001])
        # row-vector convention: row = [x y z 1] @ M   (splat/image.py:51-65)
        world2view = getWorld2View(R=R[0], t=T).transpose(0, 1)
        projection_matrix = getProjectionMatrix(znear=znear, zfar=zfar, fovX=fovX, fovY=fovY).transpose(0, 1)
        full_proj_transform = world2view.unsqueeze(0).bmm(projection_matrix.unsqueeze(0)).squeeze(0)
        camera_center = world2view.inverse()[3, :3]

        dev = self.device
        self.f_x, self.f_y, self.c_x, self.c_y = f_x.to(dev), f_y.to(dev), c_x.to(dev), c_y.to(dev)
        self.R, self.T = R.to(dev), T.to(dev)
        self.height, self.width = height.to(dev), width.to(dev)
        self.fovX, self.fovY = fovX.to(dev), fovY.to(dev)
        self.tan_fovX, self.tan_fovY = tan_fovX.to(dev), tan_fovY.to(dev)
        self.zfar, self.znear = zfar.to(dev), znear.to(dev)
        self.name = image.name
        self.world2view = world2view.contiguous().to(dev)
        self.projection_matrix = projection_matrix.contiguous().to(dev)
        self.full_proj_transform = full_proj_transform.contiguous().to(dev)
        self.camera_center = camera_center.to(dev)
        # pinhole matrices the render path never reads; kept because callers of the reference class can (splat/image.py:32-40,:68-70)
        self.intrinsic_matrix = get_intrinsic_matrix(f_x=f_x, f_y=f_y, c_x=c_x, c_y=c_y).to(dev)
        self.extrinsic_matrix = get_extrinsic_matrix(R[0], T).to(dev)
        self.projection = (self.intrinsic_matrix.cpu() @ self.extrinsic_matrix.cpu()).to(dev)
        self._packed = None

    def project_point_to_camera_perspective_projection(self, points: torch.Tensor, colors: torch.Tensor):
        """In-view points -> (pixel x, pixel y, NDC z) and their colours, in Gaussian-index order: the debug scatter of
        splat/image.py:72-90 (clip = [p 1] @ full_proj_transform; xyz / w; ndc2Pix on x and y)."""
        keep = in_view_frustum(points=points, view_matrix=self.world2view.to(points.device))
        pts = points[keep]
        hom = torch.cat([pts, torch.ones((pts.shape[0], 1), device=pts.device, dtype=pts.dtype)], dim=1)
        clip = hom @ self.full_proj_transform.to(pts.device)
        out = clip[:, :3] / clip[:, 3].unsqueeze(1)
        out[:, 0] = ndc2Pix(out[:, 0], self.width.to(pts.device))
        out[:, 1] = ndc2Pix(out[:, 1], self.height.to(pts.device))
        return out, colors[keep]

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
