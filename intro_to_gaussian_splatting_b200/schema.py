"""Record layouts of the frozen interface.

The north star keeps the reference's record layout unchanged (splat/schema.py:7-25): `PreprocessedScene` must be a
NamedTuple with exactly these twelve fields in this order, because callers unpack it positionally and by name.  The
layout is stated here as data -- (field, shape, meaning) -- and the tuple types are generated from it, so the
shapes and units the reference leaves implicit are written down next to each field.
"""

from collections import namedtuple
from typing import Tuple

# (field, per-row shape, meaning); every tensor is fp32 with M rows, M = Gaussians with z_view >= 0.2, sorted by
# depth with ties in Gaussian-index order
PREPROCESSED_FIELDS: Tuple[Tuple[str, Tuple[int, ...], str], ...] = (
    ("points", (2,), "pixel centre (x, y) -- same values as points_xy"),
    ("colors", (3,), "rgb / 256"),
    ("covariance_2d", (2, 2), "EWA covariance in pixels^2, no low-pass term"),
    ("depths", (), "z in view space"),
    ("inverse_covariance_2d", (2, 2), "[[d, -b], [-c, a]] / max(ad - bc, 1e-3); b and c are NOT forced equal"),
    ("radius", (), "ceil(3 sqrt(lambda_max)), integer-valued"),
    ("points_xy", (2,), "pixel centre (x, y)"),
    ("min_x", (), "floor(x - radius)"),
    ("min_y", (), "floor(y - radius)"),
    ("max_x", (), "ceil(x + radius)"),
    ("max_y", (), "ceil(y + radius)"),
    ("sigmoid_opacity", (1,), "sigmoid(opacity logit); the CPU render applies a second sigmoid on top"),
)

PreprocessedScene = namedtuple("PreprocessedScene", [name for name, _, _ in PREPROCESSED_FIELDS])
PreprocessedScene.__doc__ = (
    "Depth-sorted per-Gaussian records of one view (layout of splat/schema.py:13-25).\n\n"
    + "\n".join(f"  {name:<22s} (M{''.join(',' + str(d) for d in shape)})  {doc}" for name, shape, doc in PREPROCESSED_FIELDS)
)

# COLMAP point cloud as the reference's fetchPly returns it (splat/schema.py:7-10): three (N,3) numpy arrays
BasicPointCloud = namedtuple("BasicPointCloud", ("points", "colors", "normals"))
BasicPointCloud.__doc__ = "Point cloud of a PLY file: points (N,3) float, colors (N,3) in [0,1], normals (N,3)."
