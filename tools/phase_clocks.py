"""Phase clocks of the radix pass (not a benchmark): builds libgsb_b200 with -DGSB_PHASE_CLOCKS into a scratch
directory, renders a few config-3 frames with it and prints, per frame, the mean cycles thread 0 of a tile spends
between the marks of csrc/onesweep.cu (the tile pass; PHASE_MODE=0: the four depth-sort passes).  The shipped library is not touched.

    python tools/phase_clocks.py            # on a B200 box
"""
import ctypes as C
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "intro_to_gaussian_splatting_b200", "csrc")

# PHASE_MODE=0 clocks the depth sort's (key, payload) passes; default: the tile pass (keys only, entry out)
MODE = ["-DGSB_PHASE_MODE=" + os.environ["PHASE_MODE"]] if os.environ.get("PHASE_MODE") else []
work = tempfile.mkdtemp(prefix="gsb_phase_")
objs = []
for name in ("project", "binning", "onesweep", "composite", "backward", "gsb_api"):
    extra = ["--fmad=false"] if name in ("project", "composite") else []
    if name == "project":
        extra += ["-prec-div=true", "-prec-sqrt=true"]
    obj = os.path.join(work, name + ".o")
    subprocess.run(["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin",
                    "/usr/bin/g++", "-Xcompiler", "-fPIC,-O2", "-DGSB_PHASE_CLOCKS", *MODE, *extra, "-c",
                    os.path.join(CSRC, name + ".cu"), "-o", obj], check=True)
    objs.append(obj)
lib_path = os.path.join(work, "libgsb_b200.so")
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-shared", "-o",
                lib_path, *objs, "-cudart", "static"], check=True)

from intro_to_gaussian_splatting_b200 import _lib  # noqa: E402

_lib.LIB_PATH = lib_path  # load the instrumented build instead of the shipped one

import torch  # noqa: E402

from intro_to_gaussian_splatting_b200 import Rasterizer  # noqa: E402
from intro_to_gaussian_splatting_b200.colmap_io import read_camera_file, read_image_file  # noqa: E402
from intro_to_gaussian_splatting_b200.image import GaussianImage  # noqa: E402
from intro_to_gaussian_splatting_b200.synth import CONFIGS, make_scene, write_colmap_text  # noqa: E402

sc = make_scene(CONFIGS["cfg3"], n_views=6)
d = tempfile.mkdtemp()
write_colmap_text(sc, d)
cams, imgs = read_camera_file(d), read_image_file(d)
packed = [GaussianImage(cams[imgs[i].camera_id], imgs[i]).pack() for i in sorted(imgs)]
r = Rasterizer(0)
r.upload(sc.xyz.cuda(), sc.scales.cuda(), sc.quats.cuda(), (sc.rgb255 / 256).float().cuda(), sc.opacity_logit.cuda())
prm = _lib.default_params(full_cover=1)
img = torch.empty((sc.spec.height, sc.spec.width, 3), device="cuda")
lib = _lib.load()
out = (C.c_ulonglong * 16)()
names = ["load+rank", "barrier (slowest warp)", "digit scan+publish", "smem scatter", "look-back digit 0",
         "barrier (slowest digit)", "write"]
for f in range(6):
    r.render(packed[f], prm, out=img)
    torch.cuda.synchronize()
    lib.gsb_debug_phase_clocks(out)
    n = out[15]
    if f >= 3 and n:
        print(f"frame {f}: {n} tiles: " + "; ".join(f"{nm} {out[k] / n:.0f}" for k, nm in enumerate(names))
              + f"; total {sum(out[k] for k in range(7)) / n:.0f} cycles; look-back words per walk {out[8] / n:.1f}, re-reads while waiting {out[9] / n:.1f}", flush=True)
shutil.rmtree(work, ignore_errors=True)
