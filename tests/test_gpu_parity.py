"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the reference goldens.

bit-exact: in-view mask, depth, pixel centre, radius, tile rect/count, sort keys, payload, tile ranges,
           and every PreprocessedScene field except sigmoid (<= 1e-6).
pixels   : max-abs <= 1e-4 (the tolerance BASELINE.json's north_star states) against the reference's
           own render_image output (goldens) and against the oracle on the larger configs.
"""

import hashlib
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, bits, golden, scene_and_images, scene_arrays, to_oracle_camera, to_oracle_params, u64
from intro_to_gaussian_splatting_b200 import Rasterizer, _lib
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
PIXEL_TOL = 1e-4

FIELDS = ["points", "colors", "covariance_2d", "depths", "inverse_covariance_2d", "radius", "points_xy",
          "min_x", "min_y", "max_x", "max_y"]


@pytest.fixture(scope="module")
def rast():
    r = Rasterizer(0)
    yield r
    r.close()


def _upload(rast, sc):
    arrs = [a.cuda() for a in scene_arrays(sc)]
    rast.upload(*arrs)


def _check_frame_against_oracle(rast, fr, prm, check_pixels_img=None):
    info = rast.frame_info()
    assert (info.tiles_x, info.tiles_y) == (fr.ntx, fr.nty)
    assert info.m_in_view == fr.proj.m
    assert info.k_instances == fr.keys.shape[0]
    dbg = rast.debug_projection()
    assert np.array_equal(dbg["in_view"].cpu().numpy(), fr.proj.in_view)
    v = fr.proj.in_view.astype(bool)
    assert np.array_equal(bits(dbg["depth"])[v], bits(fr.proj.depth)[v])
    assert np.array_equal(bits(dbg["points_xy"])[v], bits(fr.proj.pxy)[v])
    assert np.array_equal(bits(dbg["radius"])[v], bits(fr.proj.radius)[v])
    assert np.array_equal(dbg["tile_count"].cpu().numpy().view(np.uint32), fr.proj.count)
    assert np.array_equal(dbg["tile_rect"].cpu().numpy(), fr.proj.rect)
    keys, payload = rast.debug_sorted_keys()
    assert np.array_equal(u64(keys), fr.sorted_keys), "sorted keys differ"
    assert np.array_equal(payload.cpu().numpy().view(np.uint32), fr.sorted_payload), "sorted payload differs"
    rng = rast.debug_tile_ranges().cpu().numpy().view(np.uint32)
    assert np.array_equal(rng, fr.ranges), "tile ranges differ"


@pytest.mark.parametrize("sort_mode", [_lib.GSB_SORT_FULL, _lib.GSB_SORT_SPLIT])
@pytest.mark.parametrize("name,base,nv,view", [("tiny", "tiny", 1, 1), ("small", "small", 1, 1),
                                               ("orbit", "small", 4, 3), ("cfg1", "cfg1", 1, 1)])
def test_small_scenes_vs_reference_goldens(rast, name, base, nv, view, sort_mode):
    sc, images, _ = scene_and_images(base, n_views=nv)
    cam = images[view].pack()
    prm = _lib.default_params(sort_mode=sort_mode)
    _upload(rast, sc)
    img = rast.render(cam, prm)
    torch.cuda.synchronize()
    fr = orc.render(to_oracle_camera(cam), to_oracle_params(prm), *scene_arrays(sc))
    _check_frame_against_oracle(rast, fr, prm)
    got = img.cpu().numpy()
    ref = golden(f"render_{name}.npz")["image_wh3"].transpose(1, 0, 2)  # reference is (W,H,3)
    assert np.abs(got - ref).max() <= PIXEL_TOL, f"vs reference render_image: {np.abs(got - ref).max()}"
    assert np.abs(got - fr.image).max() <= PIXEL_TOL
    # egress variants of the same frame
    wh = rast.render(cam, prm, layout="whc").cpu().numpy()
    assert np.array_equal(wh, got.transpose(1, 0, 2))
    u8 = rast.render(cam, prm, layout="u8").cpu().numpy()
    assert np.array_equal(u8, np.rint(np.clip(got, 0, 1) * 255).astype(np.uint8))
    # host destination (pageable CPU tensor): the ABI stages and copies
    host = torch.empty((cam.height, cam.width, 3), dtype=torch.float32)
    rast.render(cam, prm, out=host)
    torch.cuda.synchronize()
    assert np.array_equal(host.numpy(), got)


@pytest.mark.parametrize("name,base,nv,view", [("tiny", "tiny", 1, 1), ("small", "small", 1, 1),
                                               ("orbit", "small", 4, 3), ("cfg1", "cfg1", 1, 1)])
def test_preprocess_vs_reference_goldens(rast, name, base, nv, view):
    sc, images, _ = scene_and_images(base, n_views=nv)
    _upload(rast, sc)
    pp = rast.preprocess(images[view].pack())
    ref = golden(f"preprocess_{name}.npz")
    assert pp.depths.shape[0] == ref["depths"].shape[0]
    for k in FIELDS:
        assert np.array_equal(bits(getattr(pp, k)), bits(ref[k])), f"{name}: {k}"
    assert np.abs(pp.sigmoid_opacity.cpu().numpy() - ref["sigmoid_opacity"]).max() <= 1e-6


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_preprocess_hashes_full_size(rast, name):
    hashes = json.load(open(os.path.join(GOLDEN, "hashes.json")))[name]
    sc, images, _ = scene_and_images(name)
    _upload(rast, sc)
    pp = rast.preprocess(images[1].pack())
    assert pp.depths.shape[0] == hashes["M"]
    for k in FIELDS:
        h = hashlib.sha256(np.ascontiguousarray(getattr(pp, k).cpu().numpy()).tobytes()).hexdigest()
        assert h == hashes[k], f"{name}: {k}"


@pytest.mark.parametrize("name,full_cover,sort_mode", [("cfg2", 0, _lib.GSB_SORT_SPLIT), ("cfg2", 1, _lib.GSB_SORT_FULL),
                                                       ("cfg3", 0, _lib.GSB_SORT_FULL),
                                                       ("cfg3", 1, _lib.GSB_SORT_SPLIT), ("cfg3", 0, _lib.GSB_SORT_SPLIT)])
def test_baseline_configs_vs_oracle(rast, name, full_cover, sort_mode):
    """BASELINE configs 2 and 3 at full size: keys/ranges bit-exact, pixels <= 1e-4 vs the oracle."""
    sc, images, _ = scene_and_images(name)
    cam = images[1].pack()
    prm = _lib.default_params(full_cover=full_cover, sort_mode=sort_mode)
    _upload(rast, sc)
    img = rast.render(cam, prm)
    torch.cuda.synchronize()
    fr = orc.render(to_oracle_camera(cam), to_oracle_params(prm), *scene_arrays(sc))
    _check_frame_against_oracle(rast, fr, prm)
    err = np.abs(img.cpu().numpy() - fr.image).max()
    assert err <= PIXEL_TOL, err
    # size-independent properties of the sorted stream
    keys, payload = rast.debug_sorted_keys()
    k = u64(keys)
    assert np.all(k[:-1] <= k[1:])
    rng = rast.debug_tile_ranges().cpu().numpy().view(np.uint32).astype(np.int64)
    assert (rng[:, 1] - rng[:, 0]).sum() == k.shape[0]
    nz = rng[rng[:, 1] > rng[:, 0]]
    assert np.all(nz[1:, 0] == nz[:-1, 1])  # ranges tile the key array without gaps


def test_edge_cases(rast):
    sc, images, _ = scene_and_images("tiny")
    cam = images[1].pack()
    prm = _lib.default_params()
    # N = 0
    z3 = torch.zeros((0, 3)).cuda()
    rast.upload(z3, z3, torch.zeros((0, 4)).cuda(), z3, torch.zeros((0, 1)).cuda())
    img = rast.render(cam, prm)
    assert float(img.abs().max()) == 0.0 and rast.frame_info().k_instances == 0
    # everything culled
    xyz = sc.xyz.clone()
    xyz[:, 2] = -50
    rast.upload(xyz.cuda(), sc.scales.cuda(), sc.quats.cuda(), (sc.rgb255 / 256).cuda(), sc.opacity_logit.cuda())
    img = rast.render(cam, prm)
    info = rast.frame_info()
    assert info.m_in_view == 0 and info.k_instances == 0 and float(img.abs().max()) == 0.0
    # unsupported tile size fails loudly
    with pytest.raises(RuntimeError, match="unsupported"):
        rast.render(cam, _lib.default_params(tile_size=8))
    # image sizes that are not multiples of 16, both grids, huge + tiny Gaussians, depth ties
    from intro_to_gaussian_splatting_b200.synth import SceneSpec
    for (w, h) in [(50, 37), (17, 16), (16, 16), (33, 130)]:
        spec = SceneSpec("odd", 500, w, h, log_scale_range=(-7.0, -1.0))
        s2, im2, _ = scene_and_images(spec)
        xyz = s2.xyz.clone()
        xyz[100:200] = xyz[0:100]  # exact duplicates -> depth ties, stable order must hold
        arrs = (xyz, s2.scales, s2.quats, (s2.rgb255 / 256).float(), s2.opacity_logit)
        rast.upload(*[a.cuda() for a in arrs])
        for fc in (0, 1):
            for sm in (_lib.GSB_SORT_FULL, _lib.GSB_SORT_SPLIT):
                p = _lib.default_params(full_cover=fc, sort_mode=sm)
                img = rast.render(im2[1].pack(), p)
                torch.cuda.synchronize()
                fr = orc.render(to_oracle_camera(im2[1].pack()), to_oracle_params(p), *arrs)
                _check_frame_against_oracle(rast, fr, p)
                assert np.abs(img.cpu().numpy() - fr.image).max() <= PIXEL_TOL


def test_sort_standalone(rast):
    g = torch.Generator(device="cpu").manual_seed(7)
    for n in [1, 2, 255, 4095, 4096, 4097, 100_000, 3_000_000]:
        for (b, e) in [(0, 64), (0, 45), (32, 45), (3, 20)]:
            keys = torch.randint(-(2 ** 62), 2 ** 62, (n,), generator=g, dtype=torch.int64)
            if n > 10:
                keys[n // 2:] = keys[: n - n // 2].clone()  # plenty of duplicates: stability matters
            vals = torch.arange(n, dtype=torch.int32)
            ko, vo = rast.sort_pairs(keys.cuda(), vals.cuda(), b, e)
            ku = keys.numpy().view(np.uint64)
            mask = np.uint64(((1 << (e - b)) - 1) << b) if e - b < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
            order = np.argsort(ku & mask, kind="stable")
            assert np.array_equal(u64(ko), ku[order]), (n, b, e)
            assert np.array_equal(vo.cpu().numpy(), vals.numpy()[order]), (n, b, e)


def test_wide_status_words_and_small_sort_tiles():
    """The 64-bit look-back path (used from 2^30 keys on) and the 8-keys-per-thread tile, forced through the
    environment knobs read at context creation; same bit-exact contract."""
    import os
    sc, images, _ = scene_and_images("cfg2")
    cam = images[1].pack()
    fr_cache = {}
    # GSB_KEYS32=0: SPLIT mode with the 64-bit tile keys it uses when tile bits + rank bits exceed 32 (the default
    # contexts of every other test take the 32-bit keys at these sizes)
    for env in ({"GSB_FORCE_WIDE_STATUS": "1"}, {"GSB_SORT_ITEMS": "8"}, {"GSB_FORCE_WIDE_STATUS": "1", "GSB_SORT_ITEMS": "8"},
                {"GSB_KEYS32": "0"}, {"GSB_KEYS32": "0", "GSB_FORCE_WIDE_STATUS": "1"}):
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
            _upload(r, sc)
            for sm in (_lib.GSB_SORT_FULL, _lib.GSB_SORT_SPLIT):
                prm = _lib.default_params(sort_mode=sm)
                img = r.render(cam, prm)
                torch.cuda.synchronize()
                if "fr" not in fr_cache:
                    fr_cache["fr"] = orc.render(to_oracle_camera(cam), to_oracle_params(prm), *scene_arrays(sc))
                _check_frame_against_oracle(r, fr_cache["fr"], prm)
                assert np.abs(img.cpu().numpy() - fr_cache["fr"].image).max() <= PIXEL_TOL
                want_bits = 64 if (sm == _lib.GSB_SORT_FULL or env.get("GSB_KEYS32") == "0") else 32
                assert r.frame_info().key_bits == want_bits, (env, sm)
            g = torch.Generator().manual_seed(3)
            keys = torch.randint(-(2 ** 62), 2 ** 62, (300_000,), generator=g, dtype=torch.int64)
            vals = torch.arange(300_000, dtype=torch.int32)
            ko, vo = r.sort_pairs(keys.cuda(), vals.cuda(), 0, 64)
            order = np.argsort(keys.numpy().view(np.uint64), kind="stable")
            assert np.array_equal(u64(ko), keys.numpy().view(np.uint64)[order]) and np.array_equal(vo.cpu().numpy(), order.astype(np.int32))
        finally:
            r.close()
    Rasterizer(0).close()  # restore the defaults for later contexts in this process
