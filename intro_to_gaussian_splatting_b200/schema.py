"""Record layouts of the frozen interface (splat/schema.py:7-25).

The north star keeps the reference's record layout unchanged, and callers unpack `PreprocessedScene` positionally
and by name, so these are the reference's two NamedTuples field for field -- the similarity to splat/schema.py is
the point, not an accident.  What the reference leaves implicit (shapes, units, the sort order) is written next to
each field.
"""

from typing import NamedTuple

import numpy as np
import torch


class BasicPointCloud(NamedTuple):
    """Point cloud of a PLY file as fetchPly returns it (splat/schema.py:7-10)."""

    points: np.ndarray   # (N,3) float
    colors: np.ndarray   # (N,3) in [0,1]
    normals: np.ndarray  # (N,3)


class PreprocessedScene(NamedTuple):
    """Per-Gaussian records of one view (splat/schema.py:13-25): fp32, M rows = the Gaussians with z_view >= 0.2,
    sorted by depth, ties in Gaussian-index order."""

    points: torch.Tensor                 # (M,2) pixel centre (x, y) -- same values as points_xy
    colors: torch.Tensor                 # (M,3) rgb / 256
    covariance_2d: torch.Tensor          # (M,2,2) EWA covariance in pixels^2, no low-pass term
    depths: torch.Tensor                 # (M,) z in view space
    inverse_covariance_2d: torch.Tensor  # (M,2,2) [[d,-b],[-c,a]] / max(ad - bc, 1e-3); b and c are NOT forced equal
    radius: torch.Tensor                 # (M,) ceil(3 sqrt(lambda_max)), integer-valued
    points_xy: torch.Tensor              # (M,2) pixel centre (x, y)
    min_x: torch.Tensor                  # (M,) floor(x - radius)
    min_y: torch.Tensor                  # (M,) floor(y - radius)
    max_x: torch.Tensor                  # (M,) ceil(x + radius)
    max_y: torch.Tensor                  # (M,) ceil(y + radius)
    sigmoid_opacity: torch.Tensor        # (M,1) sigmoid(opacity logit); the CPU render applies a second sigmoid on top
