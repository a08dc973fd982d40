"""Synthetic scenes (`synth-v2`) and COLMAP text-model writer.

The reference ships no data (its notebooks fetch the treehill COLMAP scene over
the network, /root/reference/get_data.sh:1), so every workload here is a seeded
synthetic Gaussian cloud.  The generator follows SURVEY.md Appendix E exactly:
all draws come from ONE `torch.Generator().manual_seed(seed)` in the order
xyz, rgb, scales, quat, opacity.  v2 differs from the survey's v1 probe generator in one respect: exp
and the normal draws are evaluated in float64 (see _exp_portable) so that a (name, seed) pair is
bit-reproducible on ANY machine -- v1's torch.exp/torch.randn differ in the last bit between CPU ISAs,
which was caught when the GPU box disagreed with hashes made in the build container.

The scene is handed to the scene API the same way a reference user would do it:
a COLMAP *text* model directory (`cameras.txt`, `images.txt`; the formats parsed
by /root/reference/splat/read_colmap.py:87-114 and :152-189) plus a `Gaussians`
container whose `scales/quaternions/opacity` attributes are overwritten after
construction (the reference ctor hard-codes them, splat/gaussians.py:23-33).
"""

from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

# The only camera pose published in the reference (treehill image 100,
# part_1.ipynb cell 6 output) -- used as view 0 of every config but config 1.
TREEHILL_QVEC = (0.96282662, -0.23562335, 0.12748722, 0.0345476)
TREEHILL_TVEC = (0.0530637, 0.87330016, 3.58750122)


@dataclass
class SceneSpec:
    """One BASELINE.json config, fully determined by these numbers."""

    name: str
    n: int
    width: int
    height: int
    box: float = 6.0
    log_scale_range: Optional[Tuple[float, float]] = (-6.0, -3.0)  # None => keep ctor default
    const_scale: Optional[float] = None
    random_quat: bool = True
    random_opacity: bool = True
    focal_frac: float = 0.78125  # fx = fy = focal_frac * width
    qvec: Tuple[float, float, float, float] = TREEHILL_QVEC
    tvec: Tuple[float, float, float] = TREEHILL_TVEC
    n_views: int = 1
    seed: int = 1


# BASELINE.json `configs`, in order.  "cfg3" is the one the metric is quoted on.
CONFIGS: Dict[str, SceneSpec] = {
    "cfg1": SceneSpec("cfg1", 10_000, 256, 256, box=2.0, log_scale_range=None, const_scale=0.01,
                      random_quat=False, random_opacity=False, focal_frac=0.9,
                      qvec=(1.0, 0.0, 0.0, 0.0), tvec=(0.0, 0.0, 4.0)),
    "cfg2": SceneSpec("cfg2", 100_000, 800, 800),
    "cfg3": SceneSpec("cfg3", 1_000_000, 1920, 1080),
    "cfg4": SceneSpec("cfg4", 3_000_000, 1920, 1080, n_views=256),
    "cfg5": SceneSpec("cfg5", 6_000_000, 3840, 2160, box=3.0, log_scale_range=(-5.0, -2.5)),
    # small cases for tests / smoke (same generator, not BASELINE configs)
    "tiny": SceneSpec("tiny", 300, 64, 64),
    "small": SceneSpec("small", 2_000, 160, 96),
}


@dataclass
class SynthScene:
    spec: SceneSpec
    xyz: torch.Tensor          # (N,3) fp32
    rgb255: torch.Tensor       # (N,3) fp32 in [0,255): what a user passes as `colors`
    scales: torch.Tensor       # (N,3) fp32 linear
    quats: torch.Tensor        # (N,4) fp32 wxyz, unnormalised
    opacity_logit: torch.Tensor  # (N,1) fp32
    views: List[Tuple[Tuple[float, ...], Tuple[float, ...]]] = field(default_factory=list)

    @property
    def fx(self) -> float:
        return self.spec.focal_frac * self.spec.width


def _quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return (
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    )


def orbit_views(spec: SceneSpec, n_views: Optional[int] = None, period: int = 256):
    """View k = view-0 pose composed with a world rotation of 2*pi*k/period about +y:
    R_k = R_0 * R_y(theta), t_k = t_0 ('orbit')."""
    n_views = spec.n_views if n_views is None else n_views
    out = []
    for k in range(n_views):
        th = 2.0 * math.pi * k / period
        qy = (math.cos(th / 2.0), 0.0, math.sin(th / 2.0), 0.0)
        out.append((_quat_mul(spec.qvec, qy), tuple(spec.tvec)))
    return out


# torch's vectorised fp32 exp / randn kernels pick an ISA-specific code path (AVX2 vs AVX-512) and differ
# in the last bit between machines, which would make the "same" scene differ between the build container
# and the GPU box.  Only torch.rand (Mersenne twister + a fixed int->float conversion) and IEEE add/mul are
# used in fp32; every transcendental is evaluated in float64 by numpy and rounded once to fp32.
def _exp_portable(x32: torch.Tensor) -> torch.Tensor:
    return torch.from_numpy(np.exp(x32.numpy().astype(np.float64)).astype(np.float32))


def _randn_portable(n: int, c: int, g: torch.Generator) -> torch.Tensor:
    """Box-Muller on two torch.rand draws, evaluated in float64."""
    u1 = torch.rand(n, c, generator=g).numpy().astype(np.float64)
    u2 = torch.rand(n, c, generator=g).numpy().astype(np.float64)
    z = np.sqrt(-2.0 * np.log1p(-u1)) * np.cos(2.0 * np.pi * u2)
    return torch.from_numpy(z.astype(np.float32))


def make_scene(spec_or_name, n_views: Optional[int] = None, n_override: Optional[int] = None) -> SynthScene:
    spec = CONFIGS[spec_or_name] if isinstance(spec_or_name, str) else spec_or_name
    n = spec.n if n_override is None else n_override
    g = torch.Generator().manual_seed(spec.seed)
    xyz = (torch.rand(n, 3, generator=g) - 0.5) * spec.box
    rgb = torch.rand(n, 3, generator=g) * 255
    if spec.log_scale_range is not None:
        lo, hi = spec.log_scale_range
        scales = _exp_portable(torch.rand(n, 3, generator=g) * (hi - lo) + lo)
    else:
        scales = torch.ones(n, 3) * (spec.const_scale if spec.const_scale is not None else 0.001)
    if spec.random_quat:
        quats = _randn_portable(n, 4, g)
    else:
        quats = torch.zeros(n, 4)
        quats[:, 0] = 1.0
    if spec.random_opacity:
        opacity = _randn_portable(n, 1, g) * 2 + 1
    else:
        x = 0.9999 * torch.ones((n, 1), dtype=torch.float)
        opacity = torch.log(x / (1 - x))  # inverse_sigmoid, splat/utils.py:128-129
    return SynthScene(spec, xyz.float(), rgb.float(), scales.float(), quats.float(), opacity.float(),
                      orbit_views(spec, n_views))


def write_colmap_text(scene: SynthScene, directory: str) -> str:
    """Write cameras.txt / images.txt for `scene` (image ids 1..n_views, camera id 1)."""
    os.makedirs(directory, exist_ok=True)
    s = scene.spec
    fx = scene.fx
    with open(os.path.join(directory, "cameras.txt"), "w") as f:
        f.write("# synth-v1 camera\n")
        f.write(f"1 PINHOLE {s.width} {s.height} {fx!r} {fx!r} {s.width / 2!r} {s.height / 2!r}\n")
    with open(os.path.join(directory, "images.txt"), "w") as f:
        f.write("# synth-v1 poses\n")
        for k, (q, t) in enumerate(scene.views):
            vals = " ".join(repr(float(v)) for v in (*q, *t))
            f.write(f"{k + 1} {vals} 1 v{k}.jpg\n\n")
    return directory
