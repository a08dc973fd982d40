"""Randomised GPU parity: many small scenes with random poses, intrinsics, image sizes, scale ranges and
degenerate inputs, each compared with the oracle (keys / payload / ranges bit-exact, pixels <= 1e-4)."""

import math

import numpy as np
import pytest
import torch

from helpers import scene_and_images, to_oracle_camera, to_oracle_params, u64, bits
from intro_to_gaussian_splatting_b200 import Rasterizer, _lib
from intro_to_gaussian_splatting_b200.synth import SceneSpec
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _rand_quat(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return tuple(float(v) for v in q)


@pytest.fixture(scope="module")
def rast():
    r = Rasterizer(0)
    yield r
    r.close()


def _compare(rast, cam, prm, arrs, tol=1e-4):
    rast.upload(*[a.cuda() for a in arrs])
    img = rast.render(cam, prm)
    torch.cuda.synchronize()
    fr = orc.render(to_oracle_camera(cam), to_oracle_params(prm), *arrs)
    info = rast.frame_info()
    assert info.m_in_view == fr.proj.m and info.k_instances == fr.keys.shape[0]
    keys, payload = rast.debug_sorted_keys()
    assert np.array_equal(u64(keys), fr.sorted_keys)
    assert np.array_equal(payload.cpu().numpy().view(np.uint32), fr.sorted_payload)
    assert np.array_equal(rast.debug_tile_ranges().cpu().numpy().view(np.uint32), fr.ranges)
    got = img.cpu().numpy()
    both_finite = np.isfinite(got) & np.isfinite(fr.image)
    assert np.array_equal(np.isfinite(got), np.isfinite(fr.image))
    err = np.abs(got[both_finite] - fr.image[both_finite]).max() if both_finite.any() else 0.0
    assert err <= tol, err
    return fr


@pytest.mark.parametrize("seed", list(range(16)))
def test_random_scene(rast, seed):
    rng = np.random.default_rng(1000 + seed)
    w, h = int(rng.integers(17, 400)), int(rng.integers(17, 300))
    lo = float(rng.uniform(-8, -3))
    hi = lo + float(rng.uniform(0.5, 5.0))
    # camera somewhere around the cloud, looking roughly at it (or not: off-screen / behind cases matter too)
    spec = SceneSpec(f"rnd{seed}", int(rng.integers(50, 3000)), w, h, box=float(rng.uniform(0.5, 8)),
                     log_scale_range=(lo, hi), focal_frac=float(rng.uniform(0.3, 2.5)), qvec=_rand_quat(rng),
                     tvec=(float(rng.normal() * 0.5), float(rng.normal() * 0.5), float(rng.uniform(-1, 6))), seed=seed)
    sc, images, _ = scene_and_images(spec)
    arrs = [sc.xyz, sc.scales, sc.quats, (sc.rgb255 / 256).float(), sc.opacity_logit]
    prm = _lib.default_params(full_cover=int(rng.integers(0, 2)), sort_mode=int(rng.integers(1, 3)))
    _compare(rast, images[1].pack(), prm, arrs)


def test_degenerate_inputs(rast):
    spec = SceneSpec("deg", 600, 96, 80, log_scale_range=(-6.0, -2.0))
    sc, images, _ = scene_and_images(spec)
    cam = images[1].pack()
    xyz, scales, quats, col, op = sc.xyz.clone(), sc.scales.clone(), sc.quats.clone(), (sc.rgb255 / 256).float(), sc.opacity_logit.clone()
    # z_view exactly at / just around the 0.2 cull plane: place points along the view ray at chosen depths
    from intro_to_gaussian_splatting_b200.image import GaussianImage  # noqa: F401
    w2v = np.array(list(cam.world2view), np.float64).reshape(4, 4)
    R, t = w2v[:3, :3], w2v[3, :3]  # row-vector convention: v = p @ R + t
    for k, z in enumerate([0.2, np.nextafter(np.float32(0.2), np.float32(0)), np.nextafter(np.float32(0.2), np.float32(1)), 0.19999, 0.20001, 1e-3, -1.0]):
        v = np.array([0.01 * k, -0.02 * k, float(z)])
        xyz[k] = torch.tensor((v - t) @ np.linalg.inv(R), dtype=torch.float32)
    scales[10:40] = 1e-7                     # det clamp 1e-3, lambda floor 0.1
    scales[40:50] = torch.tensor([5.0, 1e-4, 1e-4])   # needle much longer than the screen
    scales[50:55] = 50.0                     # covers everything
    op[60:70] = 40.0                         # sigmoid saturates
    op[70:80] = -40.0
    quats[80:85] = 0.0                       # zero quaternion: NaN rotation in the reference, never rendered
    quats[85:90] = torch.tensor([1e-20, 0, 0, 0])
    xyz[90:100] = xyz[90]                    # ten coincident Gaussians: equal depth, tie order by index
    col[100:110] = 0.0
    arrs = [xyz, scales, quats, col, op]
    for sm in (_lib.GSB_SORT_FULL, _lib.GSB_SORT_SPLIT):
        for fc in (0, 1):
            _compare(rast, cam, _lib.default_params(full_cover=fc, sort_mode=sm), arrs)
    # PreprocessedScene of the same set, against the oracle, bit for bit (NaN patterns included)
    rast.upload(*[a.cuda() for a in arrs])
    pp, src = rast.preprocess(cam, with_source_index=True)
    mine = orc.preprocess(to_oracle_camera(cam), orc.default_params(), *arrs)
    assert np.array_equal(src.cpu().numpy(), mine["source_index"])
    for k in ("points", "covariance_2d", "depths", "inverse_covariance_2d", "radius", "min_x", "min_y", "max_x", "max_y"):
        a, b = bits(getattr(pp, k)), bits(mine[k])
        nan_a = np.isnan(a.view(np.float32))
        assert np.array_equal(nan_a, np.isnan(b.view(np.float32))), k
        assert np.array_equal(a[~nan_a], b[~nan_a]), k
