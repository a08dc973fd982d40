"""B200-native forward Gaussian-splat rasterizer: drop-in for the forward path of
dcaustin33/intro_to_gaussian_splatting (splat/gaussian_scene.py, splat/gaussians.py, splat/c).

Host side: this package (same names as the reference's `splat` package).  Device side:
libgsb_b200.so (csrc/, C ABI in include/gsb.h), hand-written CUDA for sm_100a.
"""

from .autograd import render_differentiable  # noqa: F401
from .gaussian_scene import GaussianScene  # noqa: F401
from .gaussians import Gaussians  # noqa: F401
from .image import GaussianImage  # noqa: F401
from .rasterizer import Rasterizer, ViewRenderer  # noqa: F401
from .schema import BasicPointCloud, PreprocessedScene  # noqa: F401
from .train import fit  # noqa: F401

__all__ = ["GaussianScene", "Gaussians", "GaussianImage", "Rasterizer", "ViewRenderer", "PreprocessedScene",
           "BasicPointCloud", "render_differentiable", "fit"]
