"""GPU tests of what round 2 added to the render path (all through the C ABI):

* warp-level culling in the compositing kernel: `cull_alpha = 0` frames are BIT-IDENTICAL to frames rendered
  without culling; the default (2^-30) stays inside 1e-6 of them (tolerance from north_star: 1e-4);
* two-level binning (super-tiles + expand): every super-tile shape, one-level binning, 64-bit super-tile keys,
  two radix passes over the super-tile ids -- sorted keys / payload / ranges bit-exact against the oracle;
* the queue-first frame: a count that outgrows the capacities the frame was queued with (tail queued twice);
* the saved-frame guard of gsb_render_backward (frame ids) and the leak check of gsb_destroy.
"""

import os

import numpy as np
import pytest
import torch

from helpers import scene_and_images, scene_arrays, to_oracle_camera, to_oracle_params, u64
from intro_to_gaussian_splatting_b200 import GaussianScene, Gaussians, Rasterizer, _lib, render_differentiable
from intro_to_gaussian_splatting_b200.synth import SceneSpec
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
PIXEL_TOL = 1e-4


def _with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        r = Rasterizer(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    try:
        return fn(r)
    finally:
        r.close()
        Rasterizer(0).close()  # contexts created later read the default knobs again


def _check_lists(rast, fr):
    info = rast.frame_info()
    assert info.m_in_view == fr.proj.m and info.k_instances == fr.keys.shape[0]
    keys, payload = rast.debug_sorted_keys()
    assert np.array_equal(u64(keys), fr.sorted_keys), "sorted keys differ"
    assert np.array_equal(payload.cpu().numpy().view(np.uint32), fr.sorted_payload), "sorted payload differs"
    assert np.array_equal(rast.debug_tile_ranges().cpu().numpy().view(np.uint32), fr.ranges), "tile ranges differ"


@pytest.mark.parametrize("name,full_cover", [("cfg2", 1), ("cfg2", 0), ("small", 1)])
def test_culling_exact_mode_is_bit_identical(name, full_cover):
    sc, images, _ = scene_and_images(name)
    cam = images[1].pack()
    r = Rasterizer(0)
    r.upload(*[a.cuda() for a in scene_arrays(sc)])
    off = r.render(cam, _lib.default_params(full_cover=full_cover, cull_alpha=-1.0)).clone()
    exact = r.render(cam, _lib.default_params(full_cover=full_cover, cull_alpha=0.0)).clone()
    dflt = r.render(cam, _lib.default_params(full_cover=full_cover)).clone()
    loose = r.render(cam, _lib.default_params(full_cover=full_cover, cull_alpha=1e-6)).clone()
    torch.cuda.synchronize()
    assert torch.equal(off, exact), "cull_alpha = 0 must not change a single bit"
    err = float((dflt - off).abs().max())
    assert err <= 1e-6, err
    assert float((loose - off).abs().max()) <= 1e-3  # bound: cull_alpha x list length
    fr = orc.render(to_oracle_camera(cam), to_oracle_params(_lib.default_params(full_cover=full_cover)), *scene_arrays(sc))
    assert np.abs(dflt.cpu().numpy() - fr.image).max() <= PIXEL_TOL
    assert np.abs(off.cpu().numpy() - fr.image).max() <= PIXEL_TOL
    r.close()


def test_culling_never_drops_ill_conditioned_gaussians():
    """Needle-thin, huge and degenerate (det-clamped) conics: the bound must answer "keep" whenever its margin cannot
    cover the cancellation of the reference's fp32 quadratic form; exact mode stays bit-identical."""
    spec = SceneSpec("needles", 4000, 200, 120, box=3.0, log_scale_range=(-9.0, -0.5))
    sc, images, _ = scene_and_images(spec)
    cam = images[1].pack()
    scales = sc.scales.clone()
    scales[::3, 0] *= 50.0     # needles
    scales[1::7] *= 1e-3       # sub-pixel: det clamp / lambda floor conics
    arrs = [sc.xyz, scales, sc.quats, (sc.rgb255 / 256).float(), sc.opacity_logit]
    r = Rasterizer(0)
    r.upload(*[a.cuda() for a in arrs])
    for fc in (0, 1):
        off = r.render(cam, _lib.default_params(full_cover=fc, cull_alpha=-1.0)).clone()
        exact = r.render(cam, _lib.default_params(full_cover=fc, cull_alpha=0.0)).clone()
        dflt = r.render(cam, _lib.default_params(full_cover=fc)).clone()
        torch.cuda.synchronize()
        assert torch.equal(off, exact)
        assert float((dflt - off).abs().max()) <= 2e-6
        fr = orc.render(to_oracle_camera(cam), to_oracle_params(_lib.default_params(full_cover=fc)), *arrs)
        assert np.abs(dflt.cpu().numpy() - fr.image).max() <= PIXEL_TOL
    r.close()


@pytest.mark.parametrize("env", [{"GSB_SUPER": "0,0"}, {"GSB_SUPER": "2,2"}, {"GSB_SUPER": "3,1"}, {"GSB_SUPER": "1,3"},
                                 {"GSB_SUPER": "5,0"}, {"GSB_SUPER": "1,0"}, {"GSB_KEYS32": "0"},
                                 {"GSB_SUPER": "2,2", "GSB_KEYS32": "0"}, {"GSB_SUPER": "0,0", "GSB_KEYS32": "0"}],
                         ids=lambda e: ",".join(f"{k[4:]}={v}" for k, v in e.items()))
def test_binning_variants_are_bit_exact(env):
    """One-level binning, other super-tile shapes (masks of 2 .. 32 tiles), 64-bit super-tile keys, super-tile grids that
    need two radix passes (and global-memory accumulation in the projection): identical lists and pixels."""
    cases = []
    # 2000 x 1200: 125 x 75 tiles -> 16 x 19 super-tiles of 8 x 4 (9 bits: two radix passes); 63 x 75 of 2 x 1 tiles
    wide = SceneSpec("wide2k", 20_000, 2000, 1200, log_scale_range=(-5.5, -2.5))
    for name, fc in (("cfg2", 1), ("small", 0), (wide, 1)):
        sc, images, _ = scene_and_images(name)
        cam = images[1].pack()
        prm = _lib.default_params(full_cover=fc)
        fr = orc.render(to_oracle_camera(cam), to_oracle_params(prm), *scene_arrays(sc))
        cases.append((sc, cam, prm, fr))

    def run(r):
        for sc, cam, prm, fr in cases:
            r.upload(*[a.cuda() for a in scene_arrays(sc)])
            for _ in range(2):  # second frame: steady state (no tail re-queue)
                img = r.render(cam, prm)
                torch.cuda.synchronize()
                _check_lists(r, fr)
                assert np.abs(img.cpu().numpy() - fr.image).max() <= PIXEL_TOL

    _with_env(env, run)


def test_counts_that_outgrow_the_queued_capacities():
    """A fresh context queues its first frames with guessed capacities (16 tile instances and 4 super-tile instances
    per Gaussian).  A few hundred screen-filling Gaussians exceed both by far: the device must refuse to run the
    tail, the host must grow and queue it again, and the frame must be right -- also through gsb_render_image."""
    spec = SceneSpec("huge", 300, 640, 400, box=1.0, log_scale_range=(-1.5, 0.0))
    sc, images, _ = scene_and_images(spec)
    cam = images[1].pack()
    arrs = scene_arrays(sc)
    for fc in (1, 0):
        prm = _lib.default_params(full_cover=fc)
        fr = orc.render(to_oracle_camera(cam), to_oracle_params(prm), *arrs)
        assert fr.keys.shape[0] > 64 * spec.n  # far beyond the first guess
        r = Rasterizer(0)
        r.upload(*[a.cuda() for a in arrs])
        for _ in range(2):
            img = r.render(cam, prm)
            torch.cuda.synchronize()
            _check_lists(r, fr)
            assert np.abs(img.cpu().numpy() - fr.image).max() <= PIXEL_TOL
        # the reference op's own entry point on a fresh context, rows as preprocess leaves them
        pp = r.preprocess(cam, prm)
        r2 = Rasterizer(0)
        p2 = _lib.default_params(full_cover=fc)  # REF_CPU semantics through gsb_render_image
        img2 = r2.render_preprocessed(cam.height, cam.width, 16, pp.points, pp.colors, pp.inverse_covariance_2d, pp.min_x,
                                      pp.max_x, pp.min_y, pp.max_y, pp.sigmoid_opacity, params=p2)
        torch.cuda.synchronize()
        assert np.abs(img2.cpu().numpy() - fr.image).max() <= PIXEL_TOL
        assert r2.frame_info().k_instances == fr.keys.shape[0]
        r.close()
        r2.close()


def test_growing_views_after_a_small_one():
    """Capacities follow the largest frame seen: small view, then a view with 30x the instances, then the small one."""
    small = SceneSpec("far", 3000, 320, 200, box=2.0, log_scale_range=(-6.0, -4.0))
    sc, images, _ = scene_and_images(small)
    cam = images[1].pack()
    arrs = list(scene_arrays(sc))
    big_scales = arrs[1] * 60.0
    prm = _lib.default_params(full_cover=1)
    r = Rasterizer(0)
    for scales in (arrs[1], big_scales, arrs[1]):
        a = [arrs[0], scales, arrs[2], arrs[3], arrs[4]]
        r.upload(*[t.cuda() for t in a])
        img = r.render(cam, prm)
        torch.cuda.synchronize()
        fr = orc.render(to_oracle_camera(cam), to_oracle_params(prm), *a)
        _check_lists(r, fr)
        assert np.abs(img.cpu().numpy() - fr.image).max() <= PIXEL_TOL
    r.close()


def test_backward_refuses_a_frame_that_is_no_longer_the_last():
    """render(save) -> anything else on the rasterizer -> backward must fail with NO_SAVED instead of
    reading another frame's lists; two saved forwards of one view: only the second graph may run backward."""
    sc, images, _ = scene_and_images("tiny")
    cam = images[1].pack()
    prm = _lib.default_params(full_cover=1, save_for_backward=1)
    gi = torch.ones((cam.height, cam.width, 3), device="cuda")
    arrs = [a.cuda() for a in scene_arrays(sc)]
    r = Rasterizer(0)
    r.upload(*arrs)
    r.render(cam, prm)
    fid = r.last_frame_id
    r.render_backward(cam, prm, gi, frame_id=fid)  # fine
    # 1. the reference op's entry overwrites the per-frame records with M-row data
    pp = r.preprocess(cam, _lib.default_params(full_cover=1))
    with pytest.raises(RuntimeError, match="saved"):
        r.render_backward(cam, prm, gi)
    r.render(cam, prm)
    r.render_preprocessed(cam.height, cam.width, 16, pp.points, pp.colors, pp.inverse_covariance_2d, pp.min_x, pp.max_x,
                          pp.min_y, pp.max_y, pp.sigmoid_opacity, params=_lib.default_params(full_cover=1))
    with pytest.raises(RuntimeError, match="saved"):
        r.render_backward(cam, prm, gi)
    # 2. the debug projection rewrites depth keys / counts / records
    r.render(cam, prm)
    r.debug_projection()
    with pytest.raises(RuntimeError, match="saved"):
        r.render_backward(cam, prm, gi)
    # 3. same view, same params, rendered twice: the first frame's id is stale
    r.render(cam, prm)
    first = r.last_frame_id
    r.render(cam, prm)
    with pytest.raises(RuntimeError, match="saved"):
        r.render_backward(cam, prm, gi, frame_id=first)
    r.render_backward(cam, prm, gi, frame_id=r.last_frame_id)
    # 4. through autograd: two graphs over one rasterizer, backward of the older one must raise
    leaves = [a.clone().requires_grad_(True) for a in arrs]
    img_a = render_differentiable(r, cam, *leaves, params=prm)
    img_b = render_differentiable(r, cam, *leaves, params=prm)
    with pytest.raises(RuntimeError, match="saved"):
        img_a.sum().backward()
    img_b.sum().backward()
    assert all(t.grad is not None for t in leaves)
    r.close()


def test_scene_rasterizer_shared_with_training_is_reuploaded():
    """fit()/render_differentiable on scene.rasterizer upload other tensors; the scene must notice."""
    import tempfile

    from intro_to_gaussian_splatting_b200.synth import make_scene, write_colmap_text

    sc = make_scene("tiny")
    d = tempfile.mkdtemp(prefix="gsb_t_")
    write_colmap_text(sc, d)
    g = Gaussians(points=sc.xyz.clone(), colors=sc.rgb255.clone(), model_path=d)
    g.scales, g.quaternions, g.opacity = sc.scales.cuda(), sc.quats.cuda(), sc.opacity_logit.cuda()
    scene = GaussianScene(colmap_path=d, gaussians=g)
    a = scene.render_image_cuda(1).clone()
    other = [t.cuda() for t in scene_arrays(sc)]
    other[3] = torch.zeros_like(other[3])  # black
    render_differentiable(scene.rasterizer, scene.images[1].pack(), *[t.requires_grad_(True) for t in other])
    b = scene.render_image_cuda(1).clone()
    assert torch.equal(a, b)


def test_destroy_returns_all_device_memory():
    """create -> upload -> render(save) -> backward -> destroy, 100 times: free device memory must come back."""
    sc, images, _ = scene_and_images("small")
    cam = images[1].pack()
    prm = _lib.default_params(full_cover=1, save_for_backward=1)
    arrs = [a.cuda() for a in scene_arrays(sc)]
    gi = torch.ones((cam.height, cam.width, 3), device="cuda")

    def cycle():
        r = Rasterizer(0)
        r.upload(*arrs)
        r.render(cam, prm)
        r.render_backward(cam, prm, gi)
        host = torch.empty((cam.height, cam.width, 3))
        r.render(cam, _lib.default_params(full_cover=1), out=host)
        r.render(cam, _lib.default_params(full_cover=1), out=torch.empty((cam.height, cam.width, 3), dtype=torch.uint8), layout="u8")
        r.close()

    cycle()
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(100):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 <= (8 << 20), f"{(free0 - free1) / 2**20:.1f} MiB of device memory leaked over 100 contexts"


def test_cpu_out_is_complete_on_return_and_devices_are_checked():
    sc, images, _ = scene_and_images("small")
    cam = images[1].pack()
    r = Rasterizer(0)
    r.upload(*[a.cuda() for a in scene_arrays(sc)])
    dev = r.render(cam).clone()
    host = torch.zeros((cam.height, cam.width, 3)).pin_memory()
    r.render(cam, out=host)  # no explicit synchronisation by the caller
    assert torch.equal(host, dev.cpu())
    if torch.cuda.device_count() > 1:
        with pytest.raises(RuntimeError, match="different GPU"):
            r.render(cam, out=torch.empty((cam.height, cam.width, 3), device="cuda:1"))
    r.close()
