"""View-sharded multi-GPU rendering.

The render path shards by camera view: frames are independent (`GaussianScene.images` is already a
dict of views, splat/gaussian_scene.py:35-40), so rank r of R renders views {k : k mod R == r}.
Every rank holds the whole Gaussian set (56 B/Gaussian: 336 MB at 6 M, trivial next to 180 GB of HBM);
it is broadcast ONCE from rank 0 with `torch.distributed` (NCCL over NVLink/NVSwitch on GPUs, gloo in
the CPU tests) and there is NO per-frame collective.  Splitting a single frame across GPUs is not
done: compositing is order dependent and a frame is ~1 ms.

Training is the one place with a real exchange step: with the views sharded, every
rank differentiates its own view and the five gradient arrays are summed across ranks -- ONE packed (N,14)
all-reduce per step (`allreduce_gradients`; 56 B/Gaussian), after which every rank applies the same update.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

# (columns) of the five Gaussian attribute arrays, in upload order
_WIDTHS = (3, 3, 4, 3, 1)


@dataclass(frozen=True)
class ViewShard:
    world: int
    rank: int
    n_views: int

    def my_views(self) -> List[int]:
        return list(range(self.rank, self.n_views, self.world))

    def view_of_step(self, step: int) -> int:
        """The view rendered by this rank at its `step`-th frame (wraps around the orbit)."""
        return (step * self.world + self.rank) % self.n_views

    def owner_of(self, view: int) -> int:
        return view % self.world


def broadcast_gaussians(arrays: Optional[Sequence[torch.Tensor]], n: int, device: torch.device, world: int,
                        rank: int, src: int = 0) -> List[torch.Tensor]:
    """Rank `src` passes the five attribute tensors (points, scales, quaternions, colors, opacity);
    every rank returns them on `device`.  One packed (N,14) broadcast."""
    if world == 1:
        assert arrays is not None
        return [a.to(device).float().contiguous() for a in arrays]
    packed = torch.empty((n, sum(_WIDTHS)), dtype=torch.float32, device=device)
    if rank == src:
        assert arrays is not None and all(a.shape[0] == n for a in arrays)
        packed.copy_(torch.cat([a.reshape(n, -1).float() for a in arrays], dim=1))
    dist.broadcast(packed, src=src)
    out, c = [], 0
    for w in _WIDTHS:
        out.append(packed[:, c:c + w].contiguous())
        c += w
    return out


def allreduce_gradients(grads: Sequence[Optional[torch.Tensor]], world: int, average: bool = True) -> None:
    """In place: every non-None tensor of `grads` (same shapes on every rank) becomes the sum -- or the mean --
    over ranks.  One collective: the tensors are packed into a single flat buffer, reduced, and unpacked."""
    if world == 1:
        return
    live = [g for g in grads if g is not None]
    if not live:
        return
    flat = torch.cat([g.reshape(-1).float() for g in live])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= world
    c = 0
    for g in live:
        k = g.numel()
        g.copy_(flat[c:c + k].reshape(g.shape))
        c += k


def gather_frames(frames: Sequence[torch.Tensor], shard: ViewShard) -> Optional[List[torch.Tensor]]:
    """Optional egress (not part of the fps metric): collect the per-rank frames of one
    image size on rank 0, in view order.  ONE tensor gather -- the frames stay where they are (device tensors travel
    over NVLink with NCCL, CPU tensors with gloo), nothing is pickled.  `frames` are this rank's frames in the order
    of `shard.my_views()`; ranks with one view fewer pad their block.  Returns the list on rank 0, None elsewhere."""
    if shard.world == 1:
        return list(frames)
    if len(frames) != len(shard.my_views()):
        raise RuntimeError("gather_frames: one frame per view of shard.my_views() expected")
    per_rank = (shard.n_views + shard.world - 1) // shard.world
    if per_rank == 0:
        return [] if shard.rank == 0 else None
    if len(frames) == 0:
        raise RuntimeError("gather_frames: a rank without views cannot describe the frame shape; use n_views >= world")
    block = torch.stack(list(frames) + [torch.zeros_like(frames[0])] * (per_rank - len(frames)))
    parts = [torch.empty_like(block) for _ in range(shard.world)] if shard.rank == 0 else None
    dist.gather(block, parts, dst=0)
    if shard.rank != 0:
        return None
    return [parts[v % shard.world][v // shard.world] for v in range(shard.n_views)]
