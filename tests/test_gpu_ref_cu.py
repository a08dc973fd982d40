"""REF_CU semantics pinned to the reference's OWN kernel.

oracle/build_ref_cu.py builds /root/reference/splat/c/render.cu exactly the way the reference's scene does
(load_inline, same declaration, -O1) into oracle/_ref/gsb_ref_render_cu.so, which travels to the GPU box.  Here the
reference op and gsb_render_image(semantics = REF_CU) get the same arguments -- the rows GaussianScene.preprocess
leaves, as render_image_cuda passes them (splat/gaussian_scene.py:263-285) -- and must agree within 1e-4 (the
north star's pixel tolerance; the reference evaluates expf, this library ex2.approx)."""

import importlib.util
import os

import numpy as np
import pytest
import torch

from helpers import scene_and_images
from intro_to_gaussian_splatting_b200 import GaussianScene, Gaussians
from intro_to_gaussian_splatting_b200.synth import SceneSpec

pytestmark = pytest.mark.gpu
SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "gsb_ref_render_cu.so")


def _ref_op():
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/gsb_ref_render_cu.so not built (python oracle/build_ref_cu.py needs /root/reference)")
    spec = importlib.util.spec_from_file_location("gsb_ref_render_cu", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("spec", ["tiny", "small", "cfg2", SceneSpec("cu_big", 4000, 333, 211, box=3.0, log_scale_range=(-4.0, -1.0))],
                         ids=["tiny", "small", "cfg2", "big_splats_odd_size"])
def test_ref_cu_matches_the_reference_kernel(spec):
    ref = _ref_op()
    sc, images, d = scene_and_images(spec)
    g = Gaussians(points=sc.xyz.clone(), colors=sc.rgb255.clone(), model_path=d)
    g.scales, g.quaternions, g.opacity = sc.scales.cuda(), sc.quats.cuda(), sc.opacity_logit.cuda()
    scene = GaussianScene(colmap_path=d, gaussians=g)
    pp = scene.preprocess(1)
    im = scene.images[1]
    H, W = int(im.height.item()), int(im.width.item())
    args = [t.contiguous() for t in (pp.points, pp.colors, pp.inverse_covariance_2d, pp.min_x, pp.max_x, pp.min_y, pp.max_y,
                                     pp.sigmoid_opacity)]
    want = ref.render_image(H, W, 16, *args)
    got = scene.compile_cuda_ext().render_image(im.height, im.width, 16, *args)
    torch.cuda.synchronize()
    assert got.shape == want.shape == (H, W, 3)
    err = float((got - want).abs().max())
    assert err <= 1e-4, err
    assert float(want.abs().max()) > 0.05  # the frame is not trivially empty
