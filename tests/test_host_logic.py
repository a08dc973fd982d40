"""Host-side logic that needs no GPU: COLMAP readers (text + binary), PLY round trip, synthetic scenes."""

import os
import struct
import tempfile

import numpy as np
import pytest
import torch

from intro_to_gaussian_splatting_b200 import Gaussians
from intro_to_gaussian_splatting_b200.colmap_io import (read_camera_file, read_cameras_binary, read_image_file,
                                                        read_images_binary)
from intro_to_gaussian_splatting_b200.ply_io import fetchPly, storePly
from intro_to_gaussian_splatting_b200.synth import make_scene, write_colmap_text


def test_colmap_text_and_binary_agree():
    sc = make_scene("small", n_views=3)
    d = tempfile.mkdtemp()
    write_colmap_text(sc, d)
    cams, imgs = read_camera_file(d), read_image_file(d)
    assert list(cams) == [1] and cams[1].model == "PINHOLE" and (cams[1].width, cams[1].height) == (160, 96)
    assert sorted(imgs) == [1, 2, 3] and imgs[2].name == "v1.jpg" and imgs[2].camera_id == 1
    # the same model in COLMAP's binary layout (written here by hand) must parse to the same values,
    # and the *.bin files win over *.txt like in splat/utils.py:269-290
    with open(os.path.join(d, "cameras.bin"), "wb") as f:
        f.write(struct.pack("<Q", 1))
        f.write(struct.pack("<iiQQ", 1, 1, 160, 96))
        f.write(struct.pack("<dddd", *cams[1].params))
    with open(os.path.join(d, "images.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(imgs)))
        for i in sorted(imgs):
            im = imgs[i]
            f.write(struct.pack("<idddddddi", i, *im.qvec, *im.tvec, im.camera_id))
            f.write(im.name.encode() + b"\x00")
            f.write(struct.pack("<Q", 2))
            f.write(struct.pack("<ddq", 1.0, 2.0, 7) + struct.pack("<ddq", 3.0, 4.0, -1))
    cb, ib = read_cameras_binary(os.path.join(d, "cameras.bin")), read_images_binary(os.path.join(d, "images.bin"))
    assert np.array_equal(cb[1].params, cams[1].params) and cb[1].model == "PINHOLE"
    for i in imgs:
        assert np.array_equal(ib[i].qvec, imgs[i].qvec) and np.array_equal(ib[i].tvec, imgs[i].tvec)
        assert ib[i].name == imgs[i].name
    assert read_camera_file(d)[1].width == 160 and read_image_file(d)[3].name == "v2.jpg"
    try:
        read_camera_file(tempfile.mkdtemp())
        assert False, "missing model must raise like the reference (ValueError)"
    except ValueError:
        pass


def test_ply_round_trip_and_constructor_side_effect():
    sc = make_scene("tiny")
    d = tempfile.mkdtemp()
    path = os.path.join(d, "pc.ply")
    storePly(path, sc.xyz.numpy(), sc.rgb255.numpy())
    pc = fetchPly(path)
    assert np.array_equal(pc.points, sc.xyz.numpy())
    assert np.array_equal(np.rint(pc.colors * 255).astype(np.uint8), sc.rgb255.numpy().astype(np.uint8))
    assert pc.normals.shape == (300, 3) and not pc.normals.any()
    g = Gaussians(sc.xyz, sc.rgb255, model_path=d, write_ply=True)
    assert os.path.exists(g.point_cloud_path) and fetchPly(g.point_cloud_path).points.shape == (300, 3)
    # container defaults of the reference (splat/gaussians.py:20-33)
    assert torch.allclose(g.colors.cpu(), sc.rgb255 / 256) and float(g.scales[0, 0]) == float(np.float32(0.001))
    assert torch.equal(g.quaternions.cpu()[0], torch.tensor([1.0, 0, 0, 0]))
    assert abs(float(g.opacity[0]) - float(np.log(0.9999 / (1 - 0.9999)))) < 1e-2
    c3 = g.get_3d_covariance_matrix()
    assert c3.shape == (300, 3, 3) and torch.allclose(c3[0].cpu(), torch.eye(3) * 1e-6, atol=1e-9)


def test_full_record_ply_round_trips_in_both_layouts():
    """scales / rotations / opacity persist (the reference's PLY only holds xyz + rgb)."""
    from intro_to_gaussian_splatting_b200.ply_io import load_gaussians, save_gaussians

    sc = make_scene("small")
    d = tempfile.mkdtemp()
    cols = (sc.rgb255 / 256).float()
    arrs = dict(points=sc.xyz, scales=sc.scales, quaternions=sc.quats, colors=cols, opacity=sc.opacity_logit)
    # native: bit-exact
    p1 = os.path.join(d, "native.ply")
    save_gaussians(p1, *arrs.values())
    got = load_gaussians(p1)
    for k, t in arrs.items():
        assert got[k].dtype == np.float32 and np.array_equal(got[k], t.numpy().reshape(got[k].shape)), k
    # 3DGS property names: log scales and f_dc colours round-trip to float rounding
    p2 = os.path.join(d, "point_cloud.ply")
    save_gaussians(p2, *arrs.values(), layout="3dgs")
    raw = open(p2, "rb").read(600).decode("ascii", "replace")
    for prop in ("f_dc_0", "scale_2", "rot_3", "opacity"):
        assert f"property float {prop}" in raw
    got = load_gaussians(p2)
    assert np.array_equal(got["points"], sc.xyz.numpy()) and np.array_equal(got["quaternions"], sc.quats.numpy())
    assert np.array_equal(got["opacity"], sc.opacity_logit.numpy())
    assert np.allclose(got["scales"], sc.scales.numpy(), rtol=2e-6) and np.allclose(got["colors"], cols.numpy(), atol=2e-7)
    # a trained 3DGS file also carries higher SH degrees and normals: extra properties are ignored
    from intro_to_gaussian_splatting_b200.ply_io import _read_vertices, _write
    v, _ = _read_vertices(p2)
    wide = np.zeros(v.shape[0], dtype=np.dtype(v.dtype.descr + [(f"f_rest_{i}", "<f4") for i in range(45)]))
    for k in v.dtype.names:
        wide[k] = v[k]
    _write(p2, wide)
    assert np.array_equal(load_gaussians(p2)["points"], sc.xyz.numpy())
    # the container: from_ply / save_ply
    g = Gaussians.from_ply(p1, model_path=d)
    assert torch.equal(g.scales.cpu(), sc.scales) and torch.equal(g.quaternions.cpu(), sc.quats)
    assert torch.equal(g.opacity.cpu(), sc.opacity_logit) and torch.allclose(g.colors.cpu(), cols, atol=1e-7)
    g.save_ply(os.path.join(d, "again.ply"))
    assert np.array_equal(load_gaussians(os.path.join(d, "again.ply"))["scales"], sc.scales.numpy())
    storePly(os.path.join(d, "xyz_rgb_only.ply"), sc.xyz.numpy(), sc.rgb255.numpy())
    with pytest.raises(ValueError):  # the reference's own layout has no scales / rotations / opacity to load
        load_gaussians(os.path.join(d, "xyz_rgb_only.ply"))
    with pytest.raises(ValueError):
        save_gaussians(p1, sc.xyz, sc.scales[:-1], sc.quats, cols, sc.opacity_logit)


def test_synth_is_deterministic_and_matches_its_spec():
    a, b = make_scene("small"), make_scene("small")
    for k in ("xyz", "rgb255", "scales", "quats", "opacity_logit"):
        assert torch.equal(getattr(a, k), getattr(b, k))
    assert a.xyz.abs().max() <= 3.0 and 0 <= a.rgb255.min() and a.rgb255.max() < 255
    assert a.scales.min() >= np.exp(-6.0) * 0.999 and a.scales.max() <= np.exp(-3.0) * 1.001
    c1 = make_scene("cfg1", n_override=10)
    assert torch.all(c1.scales == 0.01) and torch.all(c1.quats[:, 0] == 1) and len(c1.views) == 1
    assert len(make_scene("cfg4", n_override=4).views) == 256


def test_fit_validates_its_arguments_before_touching_the_gpu():
    """train.fit argument errors are ValueErrors raised before a rasterizer (GPU) is needed."""
    from types import SimpleNamespace

    import pytest
    import torch

    from intro_to_gaussian_splatting_b200 import _lib, fit

    cam = _lib.GsbCamera()
    cam.width, cam.height = 32, 16
    gs = SimpleNamespace(points=torch.zeros(1, 3), scales=torch.ones(1, 3), quaternions=torch.ones(1, 4),
                         colors=torch.zeros(1, 3), opacity=torch.zeros(1, 1))
    ok = torch.zeros(16, 32, 3)
    with pytest.raises(ValueError, match="one target"):
        fit(gs, [cam], [], steps=1)
    with pytest.raises(ValueError, match="unknown attributes"):
        fit(gs, [cam], [ok], steps=1, trainable=("points", "colour"))
    with pytest.raises(ValueError, match="learning rates"):
        fit(gs, [cam], [ok], steps=1, lr={"colour": 1.0})
    with pytest.raises(ValueError, match="target of shape"):
        fit(gs, [cam], [torch.zeros(32, 16, 3)], steps=1)


def test_committed_bench_lines_follow_the_contract():
    """The bench lines kept under profiles/ (outputs of bench.py on the B200 pool) carry every key of the bench
    contract; the reference arm's line carries its own."""
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ours = json.loads(open(os.path.join(root, "profiles", "r2_bench_cfg3_1gpu.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in ours, k
    assert "workload" in ours["config"] and "model" not in ours["config"]
    assert ours["n_gpus"] == 1 and ours["scaling"] == "weak" and ours["higher_is_better"] is True
    assert ours["vs_baseline"] is None and ours["data"] == "synthetic" and ours["warmup"] >= 3
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in ours["e2e"], k
    assert ours["e2e"]["d2h_bytes_per_step"] >= 1920 * 1080 * 3 * 4 and ours["e2e"]["value"] != ours["value"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in ours["roofline"], k
    # compositing is bound by issue slots: its roofline is reported in warp instructions per second
    assert ours["roofline"]["bound"] in ("hbm", "tensor", "issue")
    if ours["roofline"]["bound"] == "issue":
        assert ours["roofline"]["unit"] == "Gwarp-inst/s" and 0.0 < ours["roofline"]["frac"] <= 1.0
        assert os.path.exists(os.path.join(root, ours["roofline"]["calibration"]["file"]))
    assert any(st["stage"] == "project" and "frac_hbm" in st for st in ours["roofline"]["stages"])
    assert ours["e2e"]["d2h_only"]["gb_per_s_all_ranks"] > 0 and ours["config"]["repeats"] >= 5
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in ours["cpu_baseline"], k
    assert ours["cpu_baseline"]["kind"] in ("port", "reference")
    assert ours["gpu_launches"] > 0 and not ours["clocks"]["reasons"]
    ref = json.loads(open(os.path.join(root, "profiles", "r2_bench_reference_arm.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0
    assert ref["cpu_baseline"]["value"] == ref["value"]
