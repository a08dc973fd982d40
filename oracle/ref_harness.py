"""TEST INFRASTRUCTURE -- drives the UNMODIFIED reference in the build container.

Imports `/root/reference/splat` (read-only, never copied) so that the oracle
restatement (oracle/gs_oracle.c) can be pinned against the reference's own CPU
path, and so that golden fixtures can be generated (tests/golden/make_golden.py).
`/root/reference` does not exist on the GPU box, so nothing under `-m gpu`,
`smoke()` or `bench.py` may import this module; it is only used
  * by tests/golden/make_golden.py (run here, outputs committed), and
  * by `-m "not gpu"` tests that skip themselves when the reference is absent.

How the reference is made to run without its missing dependencies
(SURVEY.md section 8c):
  1. a stub `plyfile` module (imported at splat/utils.py:7, only used by
     fetchPly/storePly which are not on the render path);
  2. `splat.gaussians.storePly` replaced by a no-op (the ctor writes a PLY as a
     side effect, splat/gaussians.py:17-18);
  3. a synthetic COLMAP text model (synth.write_colmap_text);
  4. CPU device (there is no GPU here; the reference picks its device from
     torch.cuda.is_available(), splat/gaussians.py:16, splat/image.py:26);
  5. `torch.argsort` forced to stable=True while `preprocess` runs: the reference
     calls the unstable default (splat/gaussian_scene.py:117); depth ties are
     common and the build's contract is "ties in Gaussian-index order".  No
     reference file is edited.
"""

from __future__ import annotations

import contextlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("GSB_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "splat", "gaussian_scene.py"))


def _install_plyfile_stub() -> None:
    if "plyfile" in sys.modules:
        return
    m = types.ModuleType("plyfile")

    class PlyData:  # pragma: no cover - never exercised on the render path
        def __init__(self, *a, **k):
            pass

        @staticmethod
        def read(path):
            raise RuntimeError("plyfile stub: PLY I/O is out of scope")

        def write(self, path):
            pass

    class PlyElement:  # pragma: no cover
        @staticmethod
        def describe(*a, **k):
            return None

    m.PlyData = PlyData
    m.PlyElement = PlyElement
    sys.modules["plyfile"] = m


_REF = None


def load_reference():
    """Import the reference package `splat` from /root/reference."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_plyfile_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import splat.gaussian_scene as gs  # noqa
    import splat.gaussians as gg  # noqa
    import splat.image as gi  # noqa
    import splat.utils as gu  # noqa

    gg.storePly = lambda *a, **k: None
    _REF = types.SimpleNamespace(gaussian_scene=gs, gaussians=gg, image=gi, utils=gu)
    return _REF


@contextlib.contextmanager
def stable_argsort():
    orig = torch.argsort

    def _stable(x, *a, **k):
        k["stable"] = True
        return orig(x, *a, **k)

    torch.argsort = _stable
    try:
        yield
    finally:
        torch.argsort = orig


def build_reference_scene(scene, workdir=None):
    """scene: intro_to_gaussian_splatting_b200.synth.SynthScene -> reference GaussianScene."""
    from intro_to_gaussian_splatting_b200.synth import write_colmap_text

    ref = load_reference()
    workdir = workdir or tempfile.mkdtemp(prefix="gsb_ref_")
    write_colmap_text(scene, workdir)
    g = ref.gaussians.Gaussians(points=scene.xyz.clone(), colors=scene.rgb255.clone(), model_path=workdir)
    g.points = g.points.detach()
    g.colors = g.colors.detach()
    g.scales = scene.scales.clone()
    g.quaternions = scene.quats.clone()
    g.opacity = scene.opacity_logit.clone()
    return ref.gaussian_scene.GaussianScene(colmap_path=workdir, gaussians=g)


def reference_preprocess(ref_scene, image_idx: int):
    with torch.no_grad(), stable_argsort():
        return ref_scene.preprocess(image_idx)


def reference_render_image(ref_scene, image_idx: int, tile_size: int = 16) -> torch.Tensor:
    """The parity target: GaussianScene.render_image (splat/gaussian_scene.py:200-238), (W,H,3)."""
    with torch.no_grad(), stable_argsort():
        return ref_scene.render_image(image_idx, tile_size=tile_size)


def reference_preprocess_gradients(ref_scene, image_idx: int, seed: int = 0):
    """Autograd of the UNMODIFIED reference through GaussianScene.preprocess (differentiable torch code,
    splat/gaussian_scene.py:70-144): L = sum(w_p * points) + sum(w_i * inverse_covariance_2d) + sum(w_o *
    sigmoid_opacity) + sum(w_c * colors), with the weights drawn per depth-sorted output row from
    numpy default_rng(seed) in that order (float64 standard normals, cast to fp32).
    -> (dict of weights, dict of gradients wrt points / scales / quaternions / colors / opacity)."""
    g = ref_scene.gaussians
    names = ("points", "scales", "quaternions", "colors", "opacity")
    for k in names:
        setattr(g, k, getattr(g, k).detach().clone().requires_grad_(True))
    with stable_argsort():
        pp = ref_scene.preprocess(image_idx)
    rng = np.random.default_rng(seed)
    m = pp.points.shape[0]
    w = dict(points=rng.standard_normal((m, 2)), inverse_covariance_2d=rng.standard_normal((m, 2, 2)),
             sigmoid_opacity=rng.standard_normal((m, 1)), colors=rng.standard_normal((m, 3)))
    w = {k: v.astype(np.float32) for k, v in w.items()}
    loss = sum((getattr(pp, k) * torch.from_numpy(v)).sum() for k, v in w.items())
    loss.backward()
    grads = {k: getattr(g, k).grad.detach().numpy().copy() for k in names}
    for k in names:
        setattr(g, k, getattr(g, k).detach())
    return w, grads
