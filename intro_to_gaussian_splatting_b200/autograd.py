"""Differentiable rendering: `render_differentiable` wraps gsb_render / gsb_render_backward in a
torch.autograd.Function, so the five Gaussian attribute tensors can be optimised with any torch optimiser.

This is the training step the reference announces (README.md:3) and prepares for (`requires_grad_` at
splat/gaussians.py:19-21) but never wrote -- its own CPU path cuts the graph with `.item()`
(splat/utils.py:365).  torch is plumbing: both passes run inside libgsb_b200.so, there is no fallback.
"""

from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import GsbCamera, GsbParams
from .rasterizer import Rasterizer


class _RenderFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rast: Rasterizer, cam: GsbCamera, params: GsbParams, points, scales, quaternions, colors,
                opacity):
        rast.upload(points, scales, quaternions, colors, opacity)
        image = rast.render(cam, params)
        # the saved state lives in the rasterizer's per-frame scratch: remember WHICH frame this graph belongs to, so
        # that a backward after any other render / preprocess / upload on the rasterizer fails instead of reading
        # another frame's lists
        ctx.frame_id = rast.last_frame_id
        ctx.rast, ctx.cam, ctx.params = rast, cam, params
        ctx.opacity_shape = tuple(opacity.shape)
        return image

    @staticmethod
    def backward(ctx, grad_image):
        g = ctx.rast.render_backward(ctx.cam, ctx.params, grad_image, frame_id=ctx.frame_id)
        need = ctx.needs_input_grad[3:]
        outs = [g["points"], g["scales"], g["quaternions"], g["colors"], g["opacity"].reshape(ctx.opacity_shape)]
        return (None, None, None) + tuple(o if n else None for o, n in zip(outs, need))


def render_differentiable(rast: Rasterizer, cam: GsbCamera, points: torch.Tensor, scales: torch.Tensor,
                          quaternions: torch.Tensor, colors: torch.Tensor, opacity: torch.Tensor,
                          params: Optional[GsbParams] = None) -> torch.Tensor:
    """(H,W,3) fp32 image on the rasterizer's device, differentiable with respect to the five attribute tensors
    (reference layouts, splat/gaussians.py:19-33; `opacity` is the logit).  The backward pass must run before the
    rasterizer renders anything else (it reuses that frame's sorted tile lists); GSB_E_NO_SAVED is raised otherwise.
    """
    p = _lib.default_params(**{f: getattr(params, f) for f, _ in GsbParams._fields_}) if params is not None \
        else _lib.default_params()
    p.save_for_backward = 1
    return _RenderFunction.apply(rast, cam, p, points, scales, quaternions, colors, opacity)
