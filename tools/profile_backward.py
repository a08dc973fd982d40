"""Time the training step (forward with save_for_backward + gsb_render_backward) on one config (not the bench)."""
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
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--full-cover", type=int, default=1)
a = ap.parse_args()
sc = make_scene(CONFIGS[a.config], n_views=max(a.frames, 1))
d = tempfile.mkdtemp()
write_colmap_text(sc, d)
cams, imgs = read_camera_file(d), read_image_file(d)
packed = [GaussianImage(cams[imgs[i].camera_id], imgs[i]).pack() for i in sorted(imgs)]
r = Rasterizer(0)
r.upload(sc.xyz.cuda(), sc.scales.cuda(), sc.quats.cuda(), (sc.rgb255 / 256).float().cuda(), sc.opacity_logit.cuda())
plain = _lib.default_params(full_cover=a.full_cover)
saved = _lib.default_params(full_cover=a.full_cover, save_for_backward=1)
H, W = sc.spec.height, sc.spec.width
img = torch.empty((H, W, 3), device="cuda")
gi = torch.randn((H, W, 3), device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for f in range(a.frames):
    cam = packed[f]
    ev[0].record()
    r.render(cam, plain, out=img)
    ev[1].record()
    r.render(cam, saved, out=img)
    ev[2].record()
    g = r.render_backward(cam, saved, gi)
    ev[3].record()
    torch.cuda.synchronize()
    info = r.frame_info()
    print(f"frame {f}: K={info.k_instances} forward={ev[0].elapsed_time(ev[1]):.3f} ms "
          f"forward+aux={ev[1].elapsed_time(ev[2]):.3f} ms backward={ev[2].elapsed_time(ev[3]):.3f} ms "
          f"|g_points|={g['points'].norm().item():.4g}", flush=True)
