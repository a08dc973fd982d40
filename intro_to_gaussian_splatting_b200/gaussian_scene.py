"""`GaussianScene`: the scene render API of the reference (splat/gaussian_scene.py:25-285), unchanged
on the outside, with the forward path running in libgsb_b200.so.

    scene = GaussianScene(colmap_path, gaussians)
    scene.preprocess(idx)                 -> PreprocessedScene          (:70-144)
    scene.render_image_cuda(idx, 16)      -> (H,W,3) fp32 CUDA tensor   (:263-285)
    scene.render_image(idx, 16)           -> (W,H,3) fp32 CPU tensor    (:200-238)
    scene.compile_cuda_ext().render_image(H, W, tile, means, ...)       (:240-261, splat/c/render.cu:90-101)

Semantics.  The reference's two paths disagree with each other; the parity
target is the CPU/torch path, so BOTH render_image and render_image_cuda composite with those
semantics (`semantics="ref_cpu"`): same pixels, different layout/device, as their names promise.
The arithmetic of render.cu itself (per-pixel bbox test, int-truncated means, single sigmoid, 0.99
clamp, 1e-3 cut-off) is available as `semantics="ref_cu"` and through `compile_cuda_ext()`.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib
from .gaussians import Gaussians
from .image import GaussianImage
from .rasterizer import Rasterizer, ViewRenderer
from .schema import PreprocessedScene
from .utils import compute_2d_covariance, read_camera_file, read_image_file


class _ExtShim:
    """What `compile_cuda_ext()` returns: an object with the reference op's `render_image`."""

    def __init__(self, rast: Rasterizer) -> None:
        self._rast = rast

    def render_image(self, image_height, image_width, tile_size, point_means, point_colors, inverse_covariance_2d,
                     min_x, max_x, min_y, max_y, opacity) -> torch.Tensor:
        # height/width arrive as 1-element float tensors in the reference (splat/image.py:37-38)
        return self._rast.render_preprocessed(int(image_height), int(image_width), int(tile_size), point_means,
                                              point_colors, inverse_covariance_2d, min_x, max_x, min_y, max_y, opacity)


class GaussianScene(nn.Module):
    def __init__(self, colmap_path: str, gaussians: Gaussians, full_cover: bool = False,
                 sort_mode: str = "auto") -> None:
        super().__init__()
        camera_dict = read_camera_file(colmap_path)
        image_dict = read_image_file(colmap_path)
        self.images = {}
        for idx in image_dict.keys():
            image = image_dict[idx]
            self.images[idx] = GaussianImage(camera=camera_dict[image.camera_id], image=image)
        self.gaussians = gaussians
        self.full_cover = bool(full_cover)  # False = the reference tile grid (last row/column never rendered)
        self.sort_mode = {"auto": _lib.GSB_SORT_AUTO, "full": _lib.GSB_SORT_FULL, "split": _lib.GSB_SORT_SPLIT}[sort_mode]
        self._rast: Optional[Rasterizer] = None
        self._uploaded_sig = None

    # ---- native context + scene residency ------------------------------------------------------
    @property
    def rasterizer(self) -> Rasterizer:
        if self._rast is None:
            self._rast = Rasterizer()
        return self._rast

    def invalidate(self) -> None:
        """Force a re-upload of the Gaussian set on the next call (needed only after in-place edits made
        through `.data`, which do not bump the tensors' version counters)."""
        self._uploaded_sig = None

    def _sync_gaussians(self) -> Rasterizer:
        g = self.gaussians
        ts = (g.points, g.scales, g.quaternions, g.colors, g.opacity)
        # identity + version counter of the five tensors; the tensors themselves are kept referenced so
        # that neither their id() nor their storage address can be recycled by a replacement
        rast = self.rasterizer
        prev = self._uploaded_sig
        # ... and the rasterizer's upload generation: somebody else (render_differentiable, fit) may have uploaded
        # other tensors into this shared rasterizer since
        same = (prev is not None and prev[0] == rast.upload_generation and
                all(a is b and a._version == v for a, (b, v) in zip(ts, prev[1])))
        if not same:
            rast.upload(*ts)
            self._uploaded_sig = (rast.upload_generation, tuple((t, t._version) for t in ts))
        return rast

    def _params(self, tile_size: int, **over):
        return _lib.default_params(tile_size=int(tile_size), full_cover=int(self.full_cover), sort_mode=self.sort_mode, **over)

    # ---- the reference API ----------------------------------------------------------------------
    def render_points_image(self, image_idx: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(pixel x, pixel y, NDC z) + colours of the in-view Gaussians in index order: the debug scatter of
        splat/gaussian_scene.py:44-51, same delegation to the image as there."""
        return self.images[image_idx].project_point_to_camera_perspective_projection(self.gaussians.points,
                                                                                      self.gaussians.colors)

    def get_2d_covariance(self, image_idx: int, points: torch.Tensor, covariance_3d: torch.Tensor) -> torch.Tensor:
        """(M,2,2) EWA covariance of `points` for this view (splat/gaussian_scene.py:53-68).  A torch helper for
        callers of the public method; frames get theirs from csrc/project.cu."""
        im = self.images[image_idx]
        return compute_2d_covariance(points=points, extrinsic_matrix=im.world2view.to(points.device),
                                     covariance_3d=covariance_3d, tan_fovX=im.tan_fovX.to(points.device),
                                     tan_fovY=im.tan_fovY.to(points.device), focal_x=im.f_x.to(points.device),
                                     focal_y=im.f_y.to(points.device))

    # render_pixel / render_tile (splat/gaussian_scene.py:146-198) are the body of composite_fast_kernel; as public
    # methods they run that kernel on the caller's depth-ordered list: the op's own entry point (gsb_render_image,
    # REF_CPU semantics: one sigmoid on `opacities` inside, exactly like :164) on a canvas that reaches the region,
    # every row's bounding box set to the region so that it is a candidate for all of its pixels.
    def _render_region(self, x0: int, y0: int, w: int, h: int, means, colors, opacities, inverse_covariance,
                       min_weight: float) -> torch.Tensor:
        if x0 < 0 or y0 < 0:
            raise RuntimeError("render_tile / render_pixel: pixel coordinates must be non-negative")
        m = means.shape[0]
        dev = self.rasterizer.device
        box = lambda v: torch.full((m,), float(v), device=dev)  # noqa: E731
        prm = _lib.default_params(full_cover=1, min_weight=float(min_weight))
        img = self.rasterizer.render_preprocessed(y0 + h, x0 + w, 16, means.to(dev), colors.to(dev),
                                                  inverse_covariance.to(dev), box(x0), box(x0 + w - 1), box(y0),
                                                  box(y0 + h - 1), opacities.to(dev), params=prm)
        return img[y0:y0 + h, x0:x0 + w]  # (h, w, 3), [y][x]

    def render_pixel(self, pixel_coords: torch.Tensor, points_in_tile_mean: torch.Tensor, colors: torch.Tensor,
                     opacities: torch.Tensor, inverse_covariance: torch.Tensor, min_weight: float = 0.000001) -> torch.Tensor:
        """Front-to-back blend of a depth-ordered list at one pixel -> (1,1,3) (splat/gaussian_scene.py:146-171)."""
        px, py = (int(round(float(v))) for v in pixel_coords.reshape(-1)[:2])
        if float(pixel_coords.reshape(-1)[0]) != px or float(pixel_coords.reshape(-1)[1]) != py:
            raise RuntimeError("render_pixel: pixel coordinates must be integer-valued (the reference passes integer pixels)")
        out = self._render_region(px, py, 1, 1, points_in_tile_mean, colors, opacities, inverse_covariance, min_weight)
        return out.reshape(1, 1, 3).to(points_in_tile_mean.device)

    def render_tile(self, x_min: int, y_min: int, points_in_tile_mean: torch.Tensor, colors: torch.Tensor,
                    opacities: torch.Tensor, inverse_covariance: torch.Tensor, tile_size: int = 16) -> torch.Tensor:
        """(tile_size, tile_size, 3) CPU tensor indexed [x % tile_size][y % tile_size] for pixels x_min..x_min+T-1,
        y_min..y_min+T-1 (splat/gaussian_scene.py:173-198); the list must be in depth order."""
        T = int(tile_size)
        reg = self._render_region(int(x_min), int(y_min), T, T, points_in_tile_mean, colors, opacities,
                                  inverse_covariance, 0.000001).cpu()
        tile = torch.zeros((T, T, 3))
        xs = (torch.arange(int(x_min), int(x_min) + T) % T)
        ys = (torch.arange(int(y_min), int(y_min) + T) % T)
        tile[xs[:, None], ys[None, :]] = reg.transpose(0, 1)  # reg is [y][x]
        return tile

    def preprocess(self, image_idx: int) -> PreprocessedScene:
        rast = self._sync_gaussians()
        return rast.preprocess(self.images[image_idx].pack(), self._params(16))

    def render_image_cuda(self, image_idx: int, tile_size: int = 16, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        rast = self._sync_gaussians()
        return rast.render(self.images[image_idx].pack(), self._params(tile_size), out=out, layout="hwc")

    def render_image(self, image_idx: int, tile_size: int = 16) -> torch.Tensor:
        rast = self._sync_gaussians()
        img = rast.render(self.images[image_idx].pack(), self._params(tile_size), layout="whc")
        return img.cpu()

    def render_views(self, image_idxs, tile_size: int = 16, out: Optional[torch.Tensor] = None,
                     frames_in_flight: int = 3) -> torch.Tensor:
        """Many views at once -> (V,H,W,3).  Not in the reference (it renders one image per call); this is the
        throughput entry for orbit-style workloads: several frames in flight on independent contexts/streams,
        optional asynchronous egress into a pinned CPU `out`.  Frames equal render_image_cuda(idx) bit for bit."""
        g = self.gaussians
        ts = (g.points, g.scales, g.quaternions, g.colors, g.opacity)
        sig = tuple((id(t), t._version) for t in ts) + (frames_in_flight,)
        if getattr(self, "_view_renderer_sig", None) != sig:
            if getattr(self, "_view_renderer", None) is not None:
                self._view_renderer.close()
            self._view_renderer = ViewRenderer(*ts, frames_in_flight=frames_in_flight)
            self._view_renderer_sig = sig
            self._view_renderer_refs = ts  # keep the tensors alive so that id() cannot be recycled
        cams = [self.images[i].pack() for i in image_idxs]
        return self._view_renderer.render(cams, self._params(tile_size), out=out)

    def compile_cuda_ext(self) -> _ExtShim:
        return _ExtShim(self.rasterizer)
