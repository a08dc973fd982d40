"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container (needs /root/reference; CPU only):

    CUDA_VISIBLE_DEVICES="" python tests/golden/make_golden.py [--with-cfg1]

The reference has no golden vectors of its own; these files are outputs of
its own code (GaussianImage, GaussianScene.preprocess, GaussianScene.render_image) on the seeded
synthetic scenes of intro_to_gaussian_splatting_b200/synth.py, driven through oracle/ref_harness.py
(stub plyfile, COLMAP text model, torch.argsort forced stable).  torch version is recorded.

  camera_<scene>.npz      GaussianImage tensors for every view        (splat/image.py:19-70)
  preprocess_<scene>.npz  the 12 PreprocessedScene fields, depth-sorted (splat/gaussian_scene.py:70-144)
  tiles_<scene>.npz       what render_image handed to render_tile: per tile (x_min, y_min, count) and
                          the concatenated row indices (into the depth-sorted arrays)  (:208-237)
  render_<scene>.npz      GaussianScene.render_image output, (W,H,3) fp32  (:200-238)
  preprocess_grad_<scene>.npz  the REFERENCE'S OWN autograd through GaussianScene.preprocess: weights w_* of a
                          random linear loss over (points, inverse_covariance_2d, sigmoid_opacity, colors) and
                          its gradients g_* wrt points / scales / quaternions / colors / opacity
                          (oracle/ref_harness.py: reference_preprocess_gradients); pins the projection half of
                          oracle/backward_oracle.py
  hashes.json             sha256 of the raw bytes of each PreprocessedScene field for the big
                          configs (cfg2, cfg3 at full size), where storing the arrays is too large
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from intro_to_gaussian_splatting_b200.synth import make_scene  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402


def camera_arrays(im):
    return dict(
        world2view=im.world2view.numpy(), full_proj=im.full_proj_transform.numpy(),
        projection_matrix=im.projection_matrix.numpy(),
        f_x=im.f_x.numpy(), f_y=im.f_y.numpy(), tan_fovX=im.tan_fovX.numpy(), tan_fovY=im.tan_fovY.numpy(),
        fovX=im.fovX.numpy(), fovY=im.fovY.numpy(), width=im.width.numpy(), height=im.height.numpy(),
    )


def save_cameras(name, sc, rs):
    out = {}
    for idx, im in rs.images.items():
        for k, v in camera_arrays(im).items():
            out[f"v{idx}_{k}"] = v
    out["qvecs"] = np.array([q for q, _ in sc.views], np.float64)
    out["tvecs"] = np.array([t for _, t in sc.views], np.float64)
    np.savez_compressed(os.path.join(HERE, f"camera_{name}.npz"), **out)


def preprocess_arrays(pp):
    return {k: v.numpy() for k, v in pp._asdict().items()}


def render_with_tile_capture(rs, idx):
    """Run the reference render_image while recording the arguments of every render_tile call."""
    pp = rh.reference_preprocess(rs, idx)
    pts = pp.points.numpy()
    tiles, rows = [], []
    orig = rs.render_tile

    def spy(x_min, y_min, points_in_tile_mean, colors, opacities, inverse_covariance, tile_size=16):
        m = points_in_tile_mean.numpy()
        # recover the row indices: rows of the depth-sorted arrays, in order, whose means match
        # (means can repeat only if two Gaussians project identically; resolve by bbox masks instead)
        mask = ((pp.min_x <= x_min + tile_size) & (pp.max_x >= x_min)
                & (pp.min_y <= y_min + tile_size) & (pp.max_y >= y_min)).numpy()
        idxs = np.nonzero(mask)[0]
        assert idxs.shape[0] == m.shape[0] and np.array_equal(pts[idxs], m), "tile capture mismatch"
        tiles.append((x_min, y_min, idxs.shape[0]))
        rows.append(idxs.astype(np.int32))
        return orig(x_min=x_min, y_min=y_min, points_in_tile_mean=points_in_tile_mean, colors=colors,
                    opacities=opacities, inverse_covariance=inverse_covariance, tile_size=tile_size)

    rs.render_tile = spy
    try:
        t0 = time.time()
        img = rh.reference_render_image(rs, idx)
        dt = time.time() - t0
    finally:
        rs.render_tile = orig
    tiles = np.array(tiles, np.int32).reshape(-1, 3)
    rows = np.concatenate(rows) if rows else np.zeros(0, np.int32)
    return img.numpy(), tiles, rows, dt


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--with-cfg1", action="store_true", help="also run the ~7 min config-1 reference render")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    assert rh.reference_available(), "needs /root/reference"
    meta = {"torch": torch.__version__, "numpy": np.__version__, "threads": torch.get_num_threads(), "timings_s": {}}
    meta_path = os.path.join(HERE, "meta.json")
    hashes_path = os.path.join(HERE, "hashes.json")
    if os.path.exists(meta_path):
        meta["timings_s"] = json.load(open(meta_path)).get("timings_s", {})
    hashes = json.load(open(hashes_path)) if os.path.exists(hashes_path) else {}

    small_cases = [("tiny", 1, 1), ("small", 1, 1), ("orbit", 4, 3)]
    if args.with_cfg1:
        small_cases.append(("cfg1", 1, 1))
    for name, n_views, idx in small_cases:
        if args.only and name != args.only:
            continue
        sc = make_scene("small" if name == "orbit" else name, n_views=n_views)
        rs = rh.build_reference_scene(sc)
        save_cameras(name, sc, rs)
        pp = rh.reference_preprocess(rs, idx)
        np.savez_compressed(os.path.join(HERE, f"preprocess_{name}.npz"), view=idx, **preprocess_arrays(pp))
        img, tiles, rows, dt = render_with_tile_capture(rs, idx)
        np.savez_compressed(os.path.join(HERE, f"tiles_{name}.npz"), view=idx, tiles=tiles, rows=rows)
        np.savez_compressed(os.path.join(HERE, f"render_{name}.npz"), view=idx, image_wh3=img)
        meta["timings_s"][f"render_image_{name}"] = round(dt, 2)
        if name in ("tiny", "small"):
            w, g = rh.reference_preprocess_gradients(rh.build_reference_scene(sc), idx, seed=7)
            np.savez_compressed(os.path.join(HERE, f"preprocess_grad_{name}.npz"), view=idx, seed=7,
                                **{f"w_{k}": v for k, v in w.items()}, **{f"g_{k}": v for k, v in g.items()})
        print(name, "render_image", f"{dt:.1f}s", img.shape, float(img.max()), flush=True)

    for name in ("cfg2", "cfg3"):
        if args.only and name != args.only:
            continue
        sc = make_scene(name)
        rs = rh.build_reference_scene(sc)
        save_cameras(name, sc, rs)
        t0 = time.time()
        pp = rh.reference_preprocess(rs, 1)
        dt = time.time() - t0
        meta["timings_s"][f"preprocess_{name}"] = round(dt, 3)
        h = {k: sha(v) for k, v in preprocess_arrays(pp).items() if k != "sigmoid_opacity"}
        h["M"] = int(pp.depths.shape[0])
        hashes[name] = h
        print(name, "preprocess", f"{dt:.2f}s", "M", h["M"], flush=True)

    json.dump(hashes, open(hashes_path, "w"), indent=1, sort_keys=True)
    json.dump(meta, open(meta_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
