"""The reference-facing Python API on the GPU: GaussianScene / Gaussians / compile_cuda_ext."""

import numpy as np
import pytest
import torch

from helpers import golden, scene_and_images, to_oracle_params
from intro_to_gaussian_splatting_b200 import GaussianScene, Gaussians, PreprocessedScene, _lib
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _scene(name, **kw):
    sc, _, d = scene_and_images(name)
    g = Gaussians(points=sc.xyz.clone(), colors=sc.rgb255.clone(), model_path=d)
    g.scales = sc.scales.clone().cuda()
    g.quaternions = sc.quats.clone().cuda()
    g.opacity = sc.opacity_logit.clone().cuda()
    return sc, GaussianScene(colmap_path=d, gaussians=g, **kw)


def test_scene_api_matches_reference_outputs():
    sc, scene = _scene("small")
    assert scene.gaussians.points.is_cuda
    img_cuda = scene.render_image_cuda(1)           # (H,W,3) on device, like render.cu
    img_cpu = scene.render_image(1)                 # (W,H,3) on CPU, like the torch path
    assert img_cuda.shape == (96, 160, 3) and img_cuda.is_cuda
    assert img_cpu.shape == (160, 96, 3) and not img_cpu.is_cuda
    ref = golden("render_small.npz")["image_wh3"]
    assert np.abs(img_cpu.numpy() - ref).max() <= 1e-4
    assert np.array_equal(img_cuda.cpu().numpy().transpose(1, 0, 2), img_cpu.numpy())
    pp = scene.preprocess(1)
    assert isinstance(pp, PreprocessedScene) and pp._fields == (
        "points", "colors", "covariance_2d", "depths", "inverse_covariance_2d", "radius", "points_xy",
        "min_x", "min_y", "max_x", "max_y", "sigmoid_opacity")
    assert pp.inverse_covariance_2d.shape[1:] == (2, 2) and pp.sigmoid_opacity.shape[1:] == (1,)
    assert torch.all(pp.depths[1:] >= pp.depths[:-1])


def test_reupload_when_attributes_change():
    sc, scene = _scene("tiny")
    a = scene.render_image_cuda(1).clone()
    scene.gaussians.opacity = scene.gaussians.opacity - 3.0
    b = scene.render_image_cuda(1)
    assert not torch.equal(a, b)
    with torch.no_grad():
        scene.gaussians.colors.mul_(0.5)  # in-place edit bumps the version counter
    c = scene.render_image_cuda(1)
    assert torch.allclose(c, b * 0.5, atol=1e-6)
    scene.gaussians.colors.data.mul_(2.0)  # edits through .data are invisible to autograd's counter ...
    scene.invalidate()                      # ... so the scene has to be told
    d = scene.render_image_cuda(1)
    assert torch.allclose(d, b, atol=1e-6)


def test_reference_op_signature_ref_cu():
    """ext.render_image(H, W, tile, means, colors, inv_cov, min_x, max_x, min_y, max_y, opacity): the
    call render_image_cuda makes in the reference (splat/gaussian_scene.py:270-282), REF_CU semantics."""
    sc, scene = _scene("small")
    pp = scene.preprocess(1)
    im = scene.images[1]
    ext = scene.compile_cuda_ext()
    got = ext.render_image(im.height, im.width, 16, pp.points.contiguous(), pp.colors.contiguous(),
                           pp.inverse_covariance_2d.contiguous(), pp.min_x.contiguous(), pp.max_x.contiguous(),
                           pp.min_y.contiguous(), pp.max_y.contiguous(), pp.sigmoid_opacity.contiguous())
    assert got.shape == (96, 160, 3) and got.is_cuda
    # oracle restatement of render.cu over the same per-tile lists
    W, H = 160, 96
    prm = _lib.default_params(semantics=_lib.GSB_SEM_REF_CU, min_weight=1e-3)
    rast = scene.rasterizer
    keys, payload = rast.debug_sorted_keys()
    rng = rast.debug_tile_ranges().cpu().numpy().view(np.uint32)
    info = rast.frame_info()
    assert (info.tiles_x, info.tiles_y) == (10, 6)
    f = lambda t: t.cpu().numpy()  # noqa: E731
    # every (pixel, Gaussian) pair passing the per-pixel bbox test must be in that pixel's tile list
    want = orc.composite_cu(W, H, to_oracle_params(prm), info.tiles_x, info.tiles_y, rng, f(payload).view(np.uint32),
                            f(pp.points), f(pp.inverse_covariance_2d), f(pp.colors), f(pp.sigmoid_opacity),
                            f(pp.min_x), f(pp.max_x), f(pp.min_y), f(pp.max_y))
    assert np.abs(f(got) - want).max() <= 1e-4
    M = pp.depths.shape[0]
    full = np.stack([np.zeros(1, np.uint32), np.full(1, M, np.uint32)], 1)  # one list with ALL Gaussians per pixel
    brute = orc.composite_cu(W, H, to_oracle_params(_lib.default_params(semantics=1, min_weight=1e-3, tile_size=4096)), 1, 1,
                             full, np.arange(M, dtype=np.uint32), f(pp.points), f(pp.inverse_covariance_2d), f(pp.colors),
                             f(pp.sigmoid_opacity), f(pp.min_x), f(pp.max_x), f(pp.min_y), f(pp.max_y))
    assert np.abs(f(got) - brute).max() <= 1e-4, "binning dropped a candidate the brute-force kernel would test"


def test_errors_are_runtime_errors():
    sc, scene = _scene("tiny")
    with pytest.raises(RuntimeError):
        scene.render_image_cuda(1, tile_size=7)
    with pytest.raises(KeyError):
        scene.render_image_cuda(99)
    ext = scene.compile_cuda_ext()
    pp = scene.preprocess(1)
    with pytest.raises(RuntimeError):  # mixed devices, like torch::checkAllSameGPU
        ext.render_image(64, 64, 16, pp.points.cpu(), pp.colors, pp.inverse_covariance_2d, pp.min_x, pp.max_x,
                         pp.min_y, pp.max_y, pp.sigmoid_opacity)


def test_async_host_copy_matches_synchronous_render():
    """params.async_host_copy: frame i's D2H copy overlaps frame i+1; after join + synchronise the pinned host
    images equal the synchronously rendered ones, also when the two staging images are recycled."""
    from intro_to_gaussian_splatting_b200 import Rasterizer
    from helpers import scene_arrays
    sc, images, _ = scene_and_images("small", n_views=4)
    r = Rasterizer(0)
    try:
        r.upload(*[a.cuda() for a in scene_arrays(sc)])
        ref = [r.render(images[i].pack(), _lib.default_params()).cpu() for i in (1, 2, 3, 4)]
        hosts = [torch.empty((96, 160, 3), dtype=torch.float32).pin_memory() for _ in range(4)]
        prm = _lib.default_params(async_host_copy=1)
        for k, i in enumerate((1, 2, 3, 4)):
            r.render(images[i].pack(), prm, out=hosts[k])
        r.join_host_copies()
        torch.cuda.synchronize()
        for k in range(4):
            assert torch.equal(hosts[k], ref[k]), k
    finally:
        r.close()


def test_three_contexts_in_flight_match_serial_frames():
    """bench.py keeps 3 frames in flight (3 contexts, 3 streams).  Frames rendered that way must be bit-identical
    to the same frames rendered one at a time: the contexts share nothing but the GPU."""
    from intro_to_gaussian_splatting_b200 import Rasterizer
    from helpers import scene_arrays
    sc, images, _ = scene_and_images("cfg2", n_views=6)
    arrs = [a.cuda() for a in scene_arrays(sc)]
    rasts = [Rasterizer(0) for _ in range(3)]
    try:
        for r in rasts:
            r.upload(*arrs)
        prm = _lib.default_params(full_cover=1)
        serial = [rasts[0].render(images[i + 1].pack(), prm).clone() for i in range(6)]
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream() for _ in range(3)]
        outs = [torch.empty_like(serial[0]) for _ in range(6)]
        for rep in range(3):  # several rounds so that scratch reuse across frames is exercised too
            for i in range(6):
                with torch.cuda.stream(streams[i % 3]):
                    rasts[i % 3].render(images[i + 1].pack(), prm, out=outs[i])
            torch.cuda.synchronize()
            for i in range(6):
                assert torch.equal(outs[i], serial[i]), (rep, i)
    finally:
        for r in rasts:
            r.close()


def test_render_views_equals_single_frames():
    sc, _, d = scene_and_images("small", n_views=5)
    g = Gaussians(points=sc.xyz.clone(), colors=sc.rgb255.clone(), model_path=d)
    g.scales, g.quaternions, g.opacity = sc.scales.cuda(), sc.quats.cuda(), sc.opacity_logit.cuda()
    scene = GaussianScene(colmap_path=d, gaussians=g)
    single = torch.stack([scene.render_image_cuda(i) for i in (1, 2, 3, 4, 5)])
    many = scene.render_views([1, 2, 3, 4, 5])
    torch.cuda.synchronize()
    assert torch.equal(many, single)
    host = torch.empty((5, 96, 160, 3), dtype=torch.float32).pin_memory()
    scene.render_views([5, 4, 3, 2, 1], out=host)
    torch.cuda.synchronize()
    assert torch.equal(host, single.flip(0).cpu())
