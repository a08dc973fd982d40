"""CPU tests of oracle/backward_oracle.py, the float64 autograd restatement that checks gsb_render_backward.

The reference has no backward pass, so the restatement is pinned where something of the reference exists:
its forward image against the C oracle (pinned to the reference by tests/test_oracle_golden.py), and the projection
half of its gradient against the reference's OWN autograd through GaussianScene.preprocess (fixtures made by
tests/golden/make_golden.py).  A finite-difference check covers the compositing half.
"""

import numpy as np
import pytest
import torch

import helpers
from intro_to_gaussian_splatting_b200 import _lib
from intro_to_gaussian_splatting_b200.synth import SceneSpec
from oracle import backward_oracle as bo
from oracle import oracle as orc

DENSE = SceneSpec("grad_dense", 260, 64, 48, box=2.5, log_scale_range=(-3.6, -1.6))
THICK = SceneSpec("grad_thick", 300, 48, 48, box=0.8, log_scale_range=(-2.5, -1.0))  # early termination happens


def _setup(spec, full_cover=1):
    sc, images, _ = helpers.scene_and_images(spec)
    cam = images[sorted(images)[0]].pack()
    prm = _lib.default_params(full_cover=full_cover)
    ocam, oprm = helpers.to_oracle_camera(cam), helpers.to_oracle_params(prm)
    arrs = [a.numpy() for a in helpers.scene_arrays(sc)]
    return ocam, oprm, arrs, orc.render(ocam, oprm, *arrs)


@pytest.mark.parametrize("spec", ["tiny", DENSE, THICK], ids=["tiny", "dense", "thick"])
def test_forward_matches_c_oracle(spec):
    ocam, oprm, arrs, fr = _setup(spec)
    ts = [torch.tensor(a.astype(np.float64)) for a in arrs]
    with torch.no_grad():
        img = bo.render(ocam, oprm, *ts, fr.ranges, fr.sorted_payload, fr.ntx, fr.nty).numpy()
    assert fr.image.max() > 0.05
    # float64 against the fp32 restatement: the fp32 inverse covariance carries ~1e-5 relative error
    assert np.abs(img - fr.image).max() <= 5e-5


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_projection_gradient_matches_reference_autograd(name):
    fx = helpers.golden(f"preprocess_grad_{name}.npz")
    sc, images, _ = helpers.scene_and_images(name)
    cam = images[int(fx["view"])].pack()
    prm = _lib.default_params()
    ocam, oprm = helpers.to_oracle_camera(cam), helpers.to_oracle_params(prm)
    arrs = [a.numpy() for a in helpers.scene_arrays(sc)]
    src = orc.preprocess(ocam, oprm, *arrs)["source_index"].astype(np.int64)  # depth-sorted row -> Gaussian
    ts = [torch.tensor(a.astype(np.float64), requires_grad=True) for a in arrs]
    sel = torch.as_tensor(src)
    pr = bo.project(ocam, oprm, ts[0][sel], ts[1][sel], ts[2][sel], ts[4].reshape(-1)[sel])
    w = {k[2:]: torch.tensor(fx[k].astype(np.float64)) for k in fx.files if k.startswith("w_")}
    loss = (torch.stack([pr["px"], pr["py"]], 1) * w["points"]).sum() \
        + (pr["inv"].reshape(-1, 2, 2) * w["inverse_covariance_2d"]).sum() \
        + (pr["op1"][:, None] * w["sigmoid_opacity"]).sum() + (ts[3][sel] * w["colors"]).sum()
    loss.backward()
    for k, t in zip(("points", "scales", "quaternions", "colors", "opacity"), ts):
        ref = fx[f"g_{k}"].astype(np.float64)
        got = t.grad.numpy()
        assert got.shape == ref.shape
        # the reference differentiates in fp32: allow its own rounding (ill-conditioned 2x2 inverses amplify it)
        err = np.abs(got - ref)
        assert (err <= 2e-3 * np.abs(ref) + 1e-4 * np.abs(ref).max()).all(), (k, err.max(), np.abs(ref).max())
        assert np.linalg.norm(got - ref) <= 1e-4 * np.linalg.norm(ref), k


def test_autograd_matches_finite_differences():
    ocam, oprm, arrs, fr = _setup(DENSE)
    rng = np.random.default_rng(3)
    gi = rng.standard_normal(fr.image.shape)
    _, grads = bo.gradients(ocam, oprm, arrs, gi, fr.ranges, fr.sorted_payload, fr.ntx, fr.nty)
    used = np.unique(fr.sorted_payload)
    names = ("points", "scales", "quaternions", "colors", "opacity")
    gt = torch.tensor(gi)

    def loss_at(k, row, col, delta):
        ts = [torch.tensor(a.astype(np.float64)) for a in arrs]
        ts[k][row, col] += delta
        with torch.no_grad():
            return float((bo.render(ocam, oprm, *ts, fr.ranges, fr.sorted_payload, fr.ntx, fr.nty) * gt).sum())

    checked = 0
    for k, name in enumerate(names):
        g = grads[name]
        # the two rows with the largest gradient of this attribute: a sign or factor error cannot hide there
        for row in used[np.argsort(-np.abs(g[used]).max(axis=1))[:2]]:
            col = int(np.argmax(np.abs(g[row])))
            h = 1e-6 * max(1.0, abs(float(arrs[k][row, col])))
            fd = (loss_at(k, row, col, h) - loss_at(k, row, col, -h)) / (2 * h)
            assert abs(fd - g[row, col]) <= 1e-4 * abs(g[row, col]) + 1e-7, (name, row, col, fd, g[row, col])
            checked += 1
    assert checked == 10
