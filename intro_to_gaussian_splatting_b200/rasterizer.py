"""`Rasterizer`: one GsbContext (one CUDA device) behind a small torch-facing class.

PyTorch is plumbing here: it owns device memory and streams; every computation on the render path
happens inside libgsb_b200.so (csrc/).  All methods raise RuntimeError on failure; there is no
fallback path of any kind.
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib
from ._lib import GsbCamera, GsbFrameInfo, GsbParams, check
from .schema import PreprocessedScene


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class Rasterizer:
    def __init__(self, device: Optional[int] = None) -> None:
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("intro_to_gaussian_splatting_b200: no CUDA device; this renderer has no CPU fallback")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        check(self._lib.gsb_create(C.byref(h), self.device_index), "gsb_create")
        self._h = h
        self.n = 0
        self.upload_generation = 0  # bumped by every upload(): owners of a shared rasterizer compare it (GaussianScene)
        self.last_frame_id = 0      # GsbFrameInfo.frame_id of the last render() of this rasterizer

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.gsb_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- scene -------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def upload(self, points, scales, quaternions, colors, opacity) -> None:
        """Gaussian attributes in the reference's layouts (splat/gaussians.py:19-33); CUDA or CPU tensors."""
        ts = [_f32c(t) for t in (points, scales, quaternions, colors, opacity)]
        n = ts[0].shape[0]
        shapes = [(n, 3), (n, 3), (n, 4), (n, 3)]
        for t, s in zip(ts[:4], shapes):
            if tuple(t.shape) != s:
                raise RuntimeError(f"gsb_upload: expected shape {s}, got {tuple(t.shape)}")
        if ts[4].numel() != n:
            raise RuntimeError("gsb_upload: opacity must have N elements")
        for t in ts:
            if t.is_cuda and t.device != self.device:
                raise RuntimeError("gsb_upload: tensors live on a different GPU than the rasterizer")
        with torch.cuda.device(self.device):
            # device tensors (and float/contiguous temporaries of them) are consumed by a kernel on the current
            # stream, which is the order torch's allocator already respects; host tensors are copied to a staging
            # block and synchronised inside gsb_upload -- no synchronisation is needed here
            check(self._lib.gsb_upload(self._h, n, *[_ptr(t) for t in ts], self._stream()), "gsb_upload")
        self.n = n
        self.upload_generation += 1

    # ---- rendering ---------------------------------------------------------------------------
    def render(self, cam: GsbCamera, params: Optional[GsbParams] = None, out: Optional[torch.Tensor] = None,
               layout: str = "hwc") -> torch.Tensor:
        """One frame.  layout 'hwc' -> (H,W,3) like render.cu; 'whc' -> (W,H,3) like the CPU path;
        'u8' -> (H,W,3) uint8.  `out` may be a CUDA tensor on this rasterizer's device or a (pinned) CPU tensor.
        A CPU `out` is complete when the call returns, EXCEPT with params.async_host_copy = 1, where the copy runs on
        the context's copy stream: call join_host_copies() and synchronise the current stream before reading it."""
        params = params or _lib.default_params()
        H, W = cam.height, cam.width
        shape, dtype = {"hwc": ((H, W, 3), torch.float32), "whc": ((W, H, 3), torch.float32),
                        "u8": ((H, W, 3), torch.uint8)}[layout]
        if out is None:
            out = torch.empty(shape, dtype=dtype, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous():
            raise RuntimeError(f"render: out must be contiguous {dtype} of shape {shape}")
        if out.is_cuda and out.device != self.device:
            raise RuntimeError("render: `out` lives on a different GPU than the rasterizer")
        fn = {"hwc": self._lib.gsb_render, "whc": self._lib.gsb_render_wh, "u8": self._lib.gsb_render_u8}[layout]
        with torch.cuda.device(self.device):
            check(fn(self._h, C.byref(cam), C.byref(params), _ptr(out), self._stream()), "gsb_render")
            if not out.is_cuda and not params.async_host_copy:
                torch.cuda.current_stream(self.device).synchronize()  # the caller may read a CPU `out` right away
        self.last_frame_id = int(self.frame_info().frame_id)
        return out

    def render_backward(self, cam: GsbCamera, params: GsbParams, grad_image: torch.Tensor,
                        frame_id: int = 0) -> Dict[str, torch.Tensor]:
        """Gradients of the last `render(cam, params)` of this rasterizer (params.save_for_backward must have been
        1) for dL/d image = grad_image (H,W,3): dict with points (N,3), scales (N,3), quaternions (N,4),
        colors (N,3), opacity (N,1) -- the reference's attribute names (splat/gaussians.py:19-33).
        frame_id: `last_frame_id` as it was right after the render being differentiated (0: do not check); the
        call fails with GSB_E_NO_SAVED when anything else ran on the rasterizer in between."""
        H, W = cam.height, cam.width
        g = _f32c(grad_image)
        if tuple(g.shape) != (H, W, 3):
            raise RuntimeError(f"render_backward: grad_image must have shape {(H, W, 3)}, got {tuple(g.shape)}")
        if g.is_cuda and g.device != self.device:
            raise RuntimeError("render_backward: grad_image lives on a different GPU than the rasterizer")
        n, dev = self.n, self.device
        out = {k: torch.empty((n, w), dtype=torch.float32, device=dev)
               for k, w in (("points", 3), ("scales", 3), ("quaternions", 4), ("colors", 3), ("opacity", 1))}
        with torch.cuda.device(self.device):
            check(self._lib.gsb_render_backward(self._h, C.byref(cam), C.byref(params), int(frame_id), _ptr(g),
                                                *[_ptr(out[k]) for k in ("points", "scales", "quaternions", "colors",
                                                                         "opacity")], self._stream()),
                  "gsb_render_backward")
            if not g.is_cuda:
                torch.cuda.current_stream(self.device).synchronize()  # the host gradient may be a temporary
        return out

    def join_host_copies(self) -> None:
        """Order the current stream after every asynchronous image copy (params.async_host_copy) still in flight."""
        with torch.cuda.device(self.device):
            check(self._lib.gsb_join_host_copies(self._h, self._stream()), "gsb_join_host_copies")

    def preprocess(self, cam: GsbCamera, params: Optional[GsbParams] = None, with_source_index: bool = False):
        params = params or _lib.default_params()
        n = self.n
        dev = self.device
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)  # noqa: E731
        bufs = dict(points_xy=f(n, 2), colors=f(n, 3), covariance_2d=f(n, 2, 2), depths=f(n),
                    inverse_covariance_2d=f(n, 2, 2), radius=f(n), min_x=f(n), min_y=f(n), max_x=f(n),
                    max_y=f(n), sigmoid_opacity=f(n, 1))
        src = torch.empty(n, dtype=torch.int32, device=dev)
        m = C.c_int64(0)
        with torch.cuda.device(self.device):
            check(self._lib.gsb_preprocess(self._h, C.byref(cam), C.byref(params), C.byref(m),
                                           *[_ptr(bufs[k]) for k in ("points_xy", "colors", "covariance_2d", "depths",
                                                                     "inverse_covariance_2d", "radius", "min_x", "min_y",
                                                                     "max_x", "max_y", "sigmoid_opacity")],
                                           _ptr(src), self._stream()), "gsb_preprocess")
        m = m.value
        b = {k: v[:m] for k, v in bufs.items()}
        pp = PreprocessedScene(points=b["points_xy"], colors=b["colors"], covariance_2d=b["covariance_2d"],
                               depths=b["depths"], inverse_covariance_2d=b["inverse_covariance_2d"], radius=b["radius"],
                               points_xy=b["points_xy"].clone(), min_x=b["min_x"], min_y=b["min_y"], max_x=b["max_x"],
                               max_y=b["max_y"], sigmoid_opacity=b["sigmoid_opacity"])
        return (pp, src[:m]) if with_source_index else pp

    def render_preprocessed(self, height: int, width: int, tile_size: int, point_means, point_colors,
                            inverse_covariance_2d, min_x, max_x, min_y, max_y, opacity,
                            params: Optional[GsbParams] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The reference op, argument for argument (splat/c/render.cu:90-101)."""
        if params is None:
            params = _lib.default_params(semantics=_lib.GSB_SEM_REF_CU, min_weight=1e-3)  # render.cu:73
        ts = [_f32c(t) for t in (point_means, point_colors, inverse_covariance_2d, min_x, max_x, min_y, max_y, opacity)]
        devs = {t.device for t in ts}
        if len(devs) != 1:  # torch::checkAllSameGPU, render.cu:104-112
            raise RuntimeError("render_image: all tensors must be on the same device")
        m = ts[0].shape[0]
        H, W = int(height), int(width)
        if out is None:
            out = torch.empty((H, W, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.gsb_render_image(self._h, H, W, int(tile_size), m, *[_ptr(t) for t in ts],
                                             C.byref(params), _ptr(out), self._stream()), "gsb_render_image")
            torch.cuda.current_stream(self.device).synchronize()  # `ts` may be temporaries
        return out

    # ---- parity / debug surface ----------------------------------------------------------------
    def frame_info(self) -> GsbFrameInfo:
        info = GsbFrameInfo()
        check(self._lib.gsb_frame_info(self._h, C.byref(info)), "gsb_frame_info")
        return info

    def debug_projection(self) -> Dict[str, torch.Tensor]:
        n = int(self.frame_info().n)
        dev = self.device
        out = dict(in_view=torch.empty(n, dtype=torch.uint8, device=dev),
                   depth=torch.empty(n, dtype=torch.float32, device=dev),
                   points_xy=torch.empty((n, 2), dtype=torch.float32, device=dev),
                   radius=torch.empty(n, dtype=torch.float32, device=dev),
                   tile_rect=torch.empty((n, 4), dtype=torch.int32, device=dev),
                   tile_count=torch.empty(n, dtype=torch.int32, device=dev))
        check(self._lib.gsb_debug_projection(self._h, *[_ptr(out[k]) for k in
                                                        ("in_view", "depth", "points_xy", "radius", "tile_rect", "tile_count")]),
              "gsb_debug_projection")
        return out

    def debug_sorted_keys(self):
        k = int(self.frame_info().k_instances)
        keys = torch.empty(k, dtype=torch.int64, device=self.device)  # bit pattern of the u64 keys
        payload = torch.empty(k, dtype=torch.int32, device=self.device)
        check(self._lib.gsb_debug_sorted_keys(self._h, _ptr(keys), _ptr(payload)), "gsb_debug_sorted_keys")
        return keys, payload

    def debug_tile_ranges(self) -> torch.Tensor:
        info = self.frame_info()
        r = torch.empty((info.tiles_x * info.tiles_y, 2), dtype=torch.int32, device=self.device)
        check(self._lib.gsb_debug_tile_ranges(self._h, _ptr(r)), "gsb_debug_tile_ranges")
        return r

    def stage_times(self) -> Dict[str, float]:
        arr = (C.c_float * _lib.GSB_NUM_STAGES)()
        check(self._lib.gsb_stage_times(self._h, C.byref(arr)), "gsb_stage_times")
        return {name: float(arr[i]) for i, name in enumerate(_lib.STAGE_NAMES)}

    def sort_pairs(self, keys: torch.Tensor, values: torch.Tensor, begin_bit: int = 0, end_bit: int = 64):
        """Stable LSD onesweep sort of (int64-bit-pattern keys, int32 payload) on bits [begin,end)."""
        assert keys.dtype == torch.int64 and values.dtype == torch.int32 and keys.is_cuda and values.is_cuda
        kin, vin = keys.clone(), values.clone()
        kout, vout = torch.empty_like(kin), torch.empty_like(vin)
        with torch.cuda.device(self.device):
            check(self._lib.gsb_sort_pairs_u64(self._h, kin.numel(), _ptr(kin), _ptr(vin), _ptr(kout), _ptr(vout),
                                               int(begin_bit), int(end_bit), self._stream()), "gsb_sort_pairs_u64")
            torch.cuda.current_stream(self.device).synchronize()
        return kout, vout


class ViewRenderer:
    """Throughput rendering of many views of one Gaussian set: `frames_in_flight` independent contexts, each on
    its own CUDA stream, take the views round-robin, so the latency-bound front of one frame (projection, depth
    sort, binning) overlaps the issue-bound compositing of another.  This is how the view-sharded orbit workload is
    driven on every rank (bench.py); the frames are bit-identical to frames rendered one at a time."""

    def __init__(self, points, scales, quaternions, colors, opacity, frames_in_flight: int = 3,
                 device: Optional[int] = None) -> None:
        self.rasts = [Rasterizer(device) for _ in range(max(1, int(frames_in_flight)))]
        self.device = self.rasts[0].device
        for r in self.rasts:
            r.upload(points, scales, quaternions, colors, opacity)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.rasts]

    def close(self) -> None:
        for r in self.rasts:
            r.close()

    def render(self, cams: Sequence[GsbCamera], params: Optional[GsbParams] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """cams: V packed cameras of one image size.  out: (V,H,W,3) fp32, CUDA or pinned CPU (asynchronous egress);
        allocated on the device when omitted.  Work is ordered after the current stream on entry, and the current
        stream is ordered after all of it on return (synchronise it before reading a CPU `out`)."""
        if len(cams) == 0:
            raise RuntimeError("ViewRenderer.render: no cameras")
        H, W = cams[0].height, cams[0].width
        if out is None:
            out = torch.empty((len(cams), H, W, 3), dtype=torch.float32, device=self.device)
        if tuple(out.shape) != (len(cams), H, W, 3) or out.dtype != torch.float32 or not out.is_contiguous():
            raise RuntimeError("ViewRenderer.render: out must be a contiguous float32 (V,H,W,3) tensor")
        params = params or _lib.default_params()
        if not out.is_cuda:
            params = _lib.default_params(**{f: getattr(params, f) for f, _ in GsbParams._fields_})
            params.async_host_copy = 1
        main = torch.cuda.current_stream(self.device)
        fork = torch.cuda.Event()
        fork.record(main)
        for st in self.streams:
            st.wait_event(fork)
        for i, cam in enumerate(cams):
            f = i % len(self.rasts)
            with torch.cuda.stream(self.streams[f]):
                self.rasts[f].render(cam, params, out=out[i])
        for f, st in enumerate(self.streams):
            with torch.cuda.stream(st):
                if not out.is_cuda:
                    self.rasts[f].join_host_copies()
                done = torch.cuda.Event()
                done.record(st)
            main.wait_event(done)
        return out
