"""Large-configuration check on the GPU: BASELINE configs 4 and 5 (3 M / 6 M Gaussians), size-independent
properties of the sorted stream (the full comparison with the oracle is tests/test_gpu_large.py)."""
import argparse
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from intro_to_gaussian_splatting_b200 import Rasterizer, _lib  # noqa: E402
from intro_to_gaussian_splatting_b200.colmap_io import read_camera_file, read_image_file  # noqa: E402
from intro_to_gaussian_splatting_b200.image import GaussianImage  # noqa: E402
from intro_to_gaussian_splatting_b200.synth import CONFIGS, make_scene, write_colmap_text  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg5")
ap.add_argument("--full-cover", type=int, default=1)
ap.add_argument("--frames", type=int, default=3)
a = ap.parse_args()

sc = make_scene(CONFIGS[a.config], n_views=1)
d = tempfile.mkdtemp()
write_colmap_text(sc, d)
cams, imgs = read_camera_file(d), read_image_file(d)
cam = GaussianImage(cams[1], imgs[1]).pack()
r = Rasterizer(0)
arrs = (sc.xyz, sc.scales, sc.quats, (sc.rgb255 / 256).float(), sc.opacity_logit)
r.upload(*[t.cuda() for t in arrs])
prm = _lib.default_params(full_cover=a.full_cover, collect_stage_times=1)
img = torch.empty((sc.spec.height, sc.spec.width, 3), device="cuda")
for f in range(a.frames):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r.render(cam, prm, out=img)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    info = r.frame_info()
    print(f"{a.config} frame {f}: N={info.n} M={info.m_in_view} K={info.k_instances} tiles={info.tiles_x}x{info.tiles_y} "
          f"wall={dt * 1e3:.2f} ms  " + " ".join(f"{k}={v * 1e3:.0f}us" for k, v in r.stage_times().items()), flush=True)
print("GPU memory in use: %.2f GB" % (torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9))

# size-independent properties
keys, payload = r.debug_sorted_keys()
k = keys.cpu().numpy().view(np.uint64)
rng = r.debug_tile_ranges().cpu().numpy().view(np.uint32).astype(np.int64)
L = rng[:, 1] - rng[:, 0]
assert np.all(k[:-1] <= k[1:]), "keys not sorted"
assert L.sum() == k.shape[0], "ranges do not cover the key array"
nz = rng[L > 0]
assert np.all(nz[1:, 0] == nz[:-1, 1]), "ranges have gaps"
tiles_of_keys = (k >> np.uint64(32)).astype(np.int64)
assert np.array_equal(np.repeat(np.arange(rng.shape[0]), L), tiles_of_keys), "range/tile mismatch"
same = k[:-1] == k[1:]
p = payload.cpu().numpy().view(np.uint32)
assert np.all(p[:-1][same] < p[1:][same]), "ties not in Gaussian-index order"
dbg = r.debug_projection()
cnt = dbg["tile_count"].cpu().numpy().view(np.uint32)
assert np.array_equal(np.bincount(p, minlength=info.n).astype(np.uint32), cnt), "instances per Gaussian != tile count"
print(f"properties ok: list mean {L.mean():.0f} max {L.max()}  image max {float(img.max()):.4f} finite {bool(torch.isfinite(img).all())}")

# the full comparison with the oracle at these sizes lives in tests/test_gpu_large.py
