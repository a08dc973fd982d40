"""Render a few frames of one config (for ncu / compute-sanitizer runs; not a benchmark)."""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from intro_to_gaussian_splatting_b200 import Rasterizer, _lib  # noqa: E402
from intro_to_gaussian_splatting_b200.colmap_io import read_camera_file, read_image_file  # noqa: E402
from intro_to_gaussian_splatting_b200.image import GaussianImage  # noqa: E402
from intro_to_gaussian_splatting_b200.synth import CONFIGS, make_scene, write_colmap_text  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg3")
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--sort-mode", default="auto")
ap.add_argument("--full-cover", type=int, default=1)
ap.add_argument("--stage-times", type=int, default=0)
ap.add_argument("--cull-alpha", type=float, default=None)
ap.add_argument("--checksum", type=int, default=0, help="print a digest of every frame (to compare kernel variants)")
ap.add_argument("--views", default="", help="comma-separated orbit view numbers to render instead of 0..frames-1")
a = ap.parse_args()
views = [int(v) for v in a.views.split(",") if v] or list(range(max(a.frames, 1)))
sc = make_scene(CONFIGS[a.config], n_views=max(views) + 1)
d = tempfile.mkdtemp()
write_colmap_text(sc, d)
cams, imgs = read_camera_file(d), read_image_file(d)
packed = [GaussianImage(cams[imgs[i].camera_id], imgs[i]).pack() for i in sorted(imgs)]
r = Rasterizer(0)
r.upload(sc.xyz.cuda(), sc.scales.cuda(), sc.quats.cuda(), (sc.rgb255 / 256).float().cuda(), sc.opacity_logit.cuda())
mode = {"auto": 0, "full": 1, "split": 2}[a.sort_mode]
prm = _lib.default_params(full_cover=a.full_cover, sort_mode=mode, collect_stage_times=a.stage_times)
if a.cull_alpha is not None:
    prm.cull_alpha = a.cull_alpha
img = torch.empty((sc.spec.height, sc.spec.width, 3), device="cuda")
for f in views:
    r.render(packed[f], prm, out=img)
    torch.cuda.synchronize()
    info = r.frame_info()
    msg = (f"frame {f}: M={info.m_in_view} V={info.v_with_tiles} K={info.k_instances} Ks={info.k_sorted} "
           f"key_bits={info.key_bits} launches={info.kernel_launches}")
    if a.stage_times:
        msg += " " + " ".join(f"{k}={v * 1e3:.1f}us" for k, v in r.stage_times().items())
    if a.checksum:
        import hashlib
        msg += " sha256=" + hashlib.sha256(img.cpu().numpy().tobytes()).hexdigest()[:16]
    print(msg, flush=True)
