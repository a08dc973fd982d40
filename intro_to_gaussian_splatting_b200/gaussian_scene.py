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
from .utils import read_camera_file, read_image_file


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
        """Pixel centres + colours of the in-view Gaussians in index order (debug scatter,
        splat/gaussian_scene.py:44-51): served from the projection kernel's records."""
        rast = self._sync_gaussians()
        # rows come back depth-sorted; undo that to return Gaussian-index order like the reference
        pp, src = rast.preprocess(self.images[image_idx].pack(), self._params(16), with_source_index=True)
        inv = torch.argsort(src.long())
        pts = torch.cat([pp.points_xy[inv], pp.depths[inv].unsqueeze(1)], dim=1)
        return pts, pp.colors[inv]

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
