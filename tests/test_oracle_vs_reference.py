"""The oracle against the LIVE reference (only where /root/reference exists, i.e. in the build container; the
GPU box has no reference and skips this file).  The committed goldens were produced the same way
(tests/golden/make_golden.py); this test keeps that link honest for a scene that is NOT among the goldens."""

import numpy as np
import pytest

from helpers import bits, to_oracle_camera
from intro_to_gaussian_splatting_b200.image import GaussianImage  # noqa: F401
from intro_to_gaussian_splatting_b200.synth import SceneSpec, make_scene
from oracle import oracle as orc
from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="/root/reference not present")

FIELDS = ["points", "colors", "covariance_2d", "depths", "inverse_covariance_2d", "radius", "points_xy",
          "min_x", "min_y", "max_x", "max_y"]


def test_live_reference_preprocess_and_render():
    spec = SceneSpec("live", 180, 48, 48, log_scale_range=(-5.5, -2.5), seed=11, tvec=(0.1, -0.2, 3.2))
    sc = make_scene(spec)
    rs = rh.build_reference_scene(sc)
    im = rs.images[1]
    # our host-side camera must equal the reference's GaussianImage bit for bit
    from helpers import scene_and_images
    _, images, _ = scene_and_images(spec)
    mine = images[1]
    for a, b in [(mine.world2view, im.world2view), (mine.full_proj_transform, im.full_proj_transform),
                 (mine.tan_fovX, im.tan_fovX), (mine.tan_fovY, im.tan_fovY), (mine.f_x, im.f_x)]:
        assert np.array_equal(bits(a.cpu()), bits(b.cpu()))
    cam = to_oracle_camera(mine.pack())
    g = rs.gaussians
    pp = rh.reference_preprocess(rs, 1)
    got = orc.preprocess(cam, orc.default_params(), g.points, g.scales, g.quaternions, g.colors, g.opacity)
    assert got["depths"].shape[0] == pp.depths.shape[0]
    for k in FIELDS:
        assert np.array_equal(bits(got[k]), bits(getattr(pp, k))), k
    assert np.abs(got["sigmoid_opacity"] - pp.sigmoid_opacity.numpy()).max() <= 1.2e-7
    img = rh.reference_render_image(rs, 1).numpy()  # (W,H,3), ~2 s
    fr = orc.render(cam, orc.default_params(), g.points, g.scales, g.quaternions, g.colors, g.opacity)
    assert np.abs(fr.image.transpose(1, 0, 2) - img).max() <= 1e-6
