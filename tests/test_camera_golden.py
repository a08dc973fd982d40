"""Host camera conventions (utils.py / image.py) against tensors produced by the reference's own
GaussianImage (splat/image.py:19-70) -- bit-exact, every view of every golden scene."""

import numpy as np
import pytest

from helpers import golden, scene_and_images


@pytest.mark.parametrize("name,base,nv", [("tiny", "tiny", 1), ("small", "small", 1), ("orbit", "small", 4),
                                          ("cfg1", "cfg1", 1), ("cfg2", "cfg2", 1), ("cfg3", "cfg3", 1)])
def test_gaussian_image_matches_reference(name, base, nv):
    sc, images, _ = scene_and_images(base, n_views=nv, n_override=8)  # cameras do not depend on N
    g = golden(f"camera_{name}.npz")
    assert len(images) == nv
    for idx, im in images.items():
        pre = f"v{idx}_"
        for key, attr in [("world2view", im.world2view), ("full_proj", im.full_proj_transform),
                          ("projection_matrix", im.projection_matrix), ("f_x", im.f_x), ("f_y", im.f_y),
                          ("tan_fovX", im.tan_fovX), ("tan_fovY", im.tan_fovY), ("fovX", im.fovX), ("fovY", im.fovY),
                          ("width", im.width), ("height", im.height)]:
            a = attr.cpu().numpy()
            assert a.dtype == np.float32
            assert np.array_equal(a.view(np.uint32), g[pre + key].view(np.uint32)), f"{name} view {idx}: {key}"
        cam = im.pack()
        assert list(cam.world2view) == [float(v) for v in g[pre + "world2view"].reshape(-1)]
        assert (cam.width, cam.height) == (int(g[pre + "width"][0]), int(g[pre + "height"][0]))
