"""A minimal training loop over `render_differentiable` -- the loop the reference's README announces
("... and training loop", README.md:3) and its `Gaussians` container prepares for (`requires_grad_`,
splat/gaussians.py:19-21) but never contains.  torch supplies the optimiser and the loss arithmetic on (H,W,3)
images (plumbing); rendering and its gradient run in libgsb_b200.so.
"""

from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch

from . import _lib
from ._lib import GsbCamera, GsbParams
from .autograd import render_differentiable
from .rasterizer import Rasterizer
from .sharding import ViewShard, allreduce_gradients

ATTRIBUTES = ("points", "scales", "quaternions", "colors", "opacity")


def l2_loss(image: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return ((image - target) ** 2).mean()


def fit(gaussians, cameras: Sequence[GsbCamera], targets: Sequence[torch.Tensor], steps: int,
        lr: Optional[Dict[str, float]] = None, params: Optional[GsbParams] = None,
        loss_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor] = l2_loss,
        rasterizer: Optional[Rasterizer] = None, trainable: Iterable[str] = ATTRIBUTES,
        shard: Optional[ViewShard] = None) -> List[float]:
    """Optimise the attribute tensors of `gaussians` (an object with .points .scales .quaternions .colors
    .opacity, e.g. `Gaussians`) so that view k renders like targets[k] ((H,W,3) fp32).  One view per step,
    round-robin; Adam with per-attribute learning rates.  The tensors are replaced by trained leaves on the
    rasterizer's device.  Returns the loss of every step.

    `shard` (view-sharded data parallelism, torch.distributed initialised, every rank starting from the same
    Gaussians): at step s rank r differentiates view shard.view_of_step(s) -- R different views per step -- and the
    gradients are averaged with one packed all-reduce, so all ranks apply the same update; the returned losses
    are this rank's."""
    rates = {"points": 1.6e-4, "scales": 5e-3, "quaternions": 1e-3, "colors": 2.5e-3, "opacity": 5e-2}
    unknown = set(lr or {}) - set(ATTRIBUTES)
    if unknown:
        raise ValueError(f"learning rates for unknown attributes {sorted(unknown)}")
    rates.update(lr or {})
    trainable = set(trainable)
    unknown = trainable - set(ATTRIBUTES)
    if unknown:
        raise ValueError(f"unknown attributes {sorted(unknown)}")
    if len(cameras) == 0 or len(cameras) != len(targets):
        raise ValueError("fit: need one target image per camera")
    for cam, tgt in zip(cameras, targets):
        if tuple(tgt.shape) != (cam.height, cam.width, 3):
            raise ValueError(f"fit: target of shape {tuple(tgt.shape)} for a {cam.height}x{cam.width} camera")
    rast = rasterizer or Rasterizer()  # raises without a CUDA device: there is no CPU path
    dev = rast.device
    leaves = {}
    for k in ATTRIBUTES:
        t = getattr(gaussians, k).detach().to(dev, torch.float32).clone()
        leaves[k] = t.requires_grad_(k in trainable)
    opt = torch.optim.Adam([{"params": [leaves[k]], "lr": rates[k]} for k in ATTRIBUTES if k in trainable], eps=1e-15)
    tg = [t.to(dev, torch.float32) for t in targets]
    prm = params or _lib.default_params()
    history: List[float] = []
    for step in range(int(steps)):
        v = step % len(cameras) if shard is None else shard.view_of_step(step) % len(cameras)
        opt.zero_grad(set_to_none=True)
        img = render_differentiable(rast, cameras[v], leaves["points"], leaves["scales"], leaves["quaternions"],
                                    leaves["colors"], leaves["opacity"], prm)
        loss = loss_fn(img, tg[v])
        loss.backward()
        if shard is not None and shard.world > 1:
            allreduce_gradients([leaves[k].grad for k in ATTRIBUTES if k in trainable], shard.world)
        opt.step()
        history.append(float(loss.detach()))
    for k in ATTRIBUTES:
        setattr(gaussians, k, leaves[k].detach().requires_grad_(k in trainable))
    if rasterizer is None:
        rast.close()
    return history
