"""The remaining public methods of the frozen scene API against outputs of the reference's own code
(tests/golden/api_small.npz, made by tests/golden/make_api_golden.py from the unmodified reference).
Host-side ones run everywhere; render_tile / render_pixel run the compositing kernel and need the GPU."""

import numpy as np
import pytest
import torch

from helpers import bits, golden, scene_and_images, scene_arrays
from intro_to_gaussian_splatting_b200 import GaussianScene, Gaussians
from intro_to_gaussian_splatting_b200.utils import compute_2d_covariance, in_view_frustum


def _scene():
    sc, images, d = scene_and_images("small")
    return sc, images, d


def test_gaussian_image_matrices_and_point_projection_match_reference():
    sc, images, _ = _scene()
    g = golden("api_small.npz")
    im = images[1]
    for k in ("intrinsic_matrix", "extrinsic_matrix", "projection"):
        assert np.array_equal(bits(getattr(im, k).cpu()), bits(g[k])), k
    pts, cols = im.project_point_to_camera_perspective_projection(sc.xyz.to(im.device), (sc.rgb255 / 256).to(im.device))
    if im.device.type == "cpu":  # same torch CPU kernels as the golden run: bit for bit
        assert np.array_equal(bits(pts), bits(g["points_image"]))
    else:
        assert np.allclose(pts.detach().cpu().numpy(), g["points_image"], rtol=1e-5, atol=1e-4)
    assert np.array_equal(bits(cols.detach().cpu()), bits(g["points_image_colors"]))


def test_get_2d_covariance_matches_reference():
    sc, images, d = _scene()
    g = golden("api_small.npz")
    im = images[1]
    gs = Gaussians(sc.xyz, sc.rgb255, model_path=d)
    gs.scales, gs.quaternions = sc.scales.to(gs.device), sc.quats.to(gs.device)
    keep = in_view_frustum(gs.points.detach(), im.world2view.to(gs.device))
    cov3 = gs.get_3d_covariance_matrix().detach()[keep]
    got = compute_2d_covariance(points=gs.points.detach()[keep], extrinsic_matrix=im.world2view.to(gs.device),
                                covariance_3d=cov3, tan_fovX=im.tan_fovX.to(gs.device), tan_fovY=im.tan_fovY.to(gs.device),
                                focal_x=im.f_x.to(gs.device), focal_y=im.f_y.to(gs.device))
    if gs.device.type == "cpu":
        assert np.array_equal(bits(got), bits(g["cov2d"]))
    else:
        ref = g["cov2d"]
        assert np.abs(got.cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.gpu
def test_render_tile_and_render_pixel_match_reference():
    """splat/gaussian_scene.py:146-198 as public methods: same arguments, same [x % T][y % T] layout, pixels within
    the north star's 1e-4 of the reference's own Python loop."""
    sc, images, d = _scene()
    g = golden("api_small.npz")
    gs = Gaussians(sc.xyz, sc.rgb255, model_path=d)
    gs.scales, gs.quaternions, gs.opacity = sc.scales.cuda(), sc.quats.cuda(), sc.opacity_logit.cuda()
    scene = GaussianScene(colmap_path=d, gaussians=gs)
    pp = scene.preprocess(1)
    rows = torch.from_numpy(g["tile_rows"]).long().cuda()
    args = dict(points_in_tile_mean=pp.points[rows], colors=pp.colors[rows], opacities=pp.sigmoid_opacity[rows],
                inverse_covariance=pp.inverse_covariance_2d[rows])
    t = scene.render_tile(x_min=64, y_min=32, tile_size=16, **args)
    assert tuple(t.shape) == (16, 16, 3) and t.device.type == "cpu"
    assert np.abs(t.numpy() - g["tile_aligned"]).max() <= 1e-4
    t = scene.render_tile(x_min=69, y_min=35, tile_size=16, **args)
    assert np.abs(t.numpy() - g["tile_unaligned"]).max() <= 1e-4
    t = scene.render_tile(x_min=64, y_min=32, tile_size=8, **args)
    assert np.abs(t.numpy() - g["tile_size8"]).max() <= 1e-4
    p = scene.render_pixel(pixel_coords=torch.Tensor([70, 40]).view(1, 2).cuda(), **args)
    assert tuple(p.shape) == (1, 1, 3) and np.abs(p.cpu().numpy() - g["pixel_70_40"]).max() <= 1e-4
    p = scene.render_pixel(pixel_coords=torch.Tensor([70, 40]).view(1, 2).cuda(), min_weight=0.5, **args)
    assert np.abs(p.cpu().numpy() - g["pixel_70_40_minw"]).max() <= 1e-4
    # the scene still renders its own Gaussians afterwards (the shared rasterizer was used for foreign rows)
    img = scene.render_image_cuda(1)
    ref = golden("render_small.npz")["image_wh3"].transpose(1, 0, 2)
    assert np.abs(img.cpu().numpy() - ref).max() <= 1e-4
    pts, cols = scene.render_points_image(1)
    assert np.allclose(pts.detach().cpu().numpy(), g["points_image"], rtol=1e-5, atol=1e-4)
