"""Record layouts kept from the reference (splat/schema.py:7-25): same names, same field order."""

from typing import NamedTuple

import numpy as np
import torch


class BasicPointCloud(NamedTuple):
    points: np.ndarray
    colors: np.ndarray
    normals: np.ndarray


class PreprocessedScene(NamedTuple):
    """Depth-sorted per-Gaussian records of one view (splat/schema.py:13-25).

    All tensors are fp32 with M rows (M = Gaussians with z_view >= 0.2); `points` and `points_xy`
    hold the same pixel centres; min/max are integer-valued floats; ties in depth keep
    Gaussian-index order."""

    points: torch.Tensor
    colors: torch.Tensor
    covariance_2d: torch.Tensor
    depths: torch.Tensor
    inverse_covariance_2d: torch.Tensor
    radius: torch.Tensor
    points_xy: torch.Tensor
    min_x: torch.Tensor
    min_y: torch.Tensor
    max_x: torch.Tensor
    max_y: torch.Tensor
    sigmoid_opacity: torch.Tensor
