"""Small driver for compute-sanitizer runs (memcheck / racecheck / synccheck) over every kernel of the library:

    compute-sanitizer --tool racecheck python tools/sanitize_driver.py

Both sort modes, both tile grids, forward with and without saved state, backward, launch by launch and as a graph, the parity getters (which run
expand_kernel and the key rebuild) and gsb_preprocess, on the `small` and `tiny` scenes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
from helpers import scene_and_images, scene_arrays  # noqa: E402

from intro_to_gaussian_splatting_b200 import Rasterizer, _lib  # noqa: E402

for name, fc, mode in (("small", 1, _lib.GSB_SORT_SPLIT), ("small", 0, _lib.GSB_SORT_FULL), ("tiny", 1, _lib.GSB_SORT_SPLIT)):
    sc, images, _ = scene_and_images(name)
    cam = images[1].pack()
    r = Rasterizer(0)
    r.upload(*[a.cuda() for a in scene_arrays(sc)])
    for save in (0, 1):
        prm = _lib.default_params(full_cover=fc, sort_mode=mode, save_for_backward=save)
        img = r.render(cam, prm)
        if save:
            r.render_backward(cam, prm, torch.ones_like(img))
    side = torch.cuda.Stream()  # any stream but the default one: the frame goes out as one CUDA graph launch
    with torch.cuda.stream(side):
        for save in (0, 1):
            prm = _lib.default_params(full_cover=fc, sort_mode=mode, save_for_backward=save)
            img = r.render(cam, prm)
            if save:
                r.render_backward(cam, prm, torch.ones_like(img))
    side.synchronize()
    r.debug_sorted_keys()
    r.debug_tile_ranges()
    pp = r.preprocess(cam)
    r.render_preprocessed(cam.height, cam.width, 16, pp.points, pp.colors, pp.inverse_covariance_2d, pp.min_x, pp.max_x,
                          pp.min_y, pp.max_y, pp.sigmoid_opacity)  # REF_CU: expand_kernel in the stream
    torch.cuda.synchronize()
    r.close()
print("driver done")
