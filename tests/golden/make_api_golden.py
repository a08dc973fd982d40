"""Golden outputs of the reference's remaining public methods (run in the build container, CPU only):

    CUDA_VISIBLE_DEVICES="" python tests/golden/make_api_golden.py

  api_small.npz   from the UNMODIFIED reference on the `small` synthetic scene, view 1:
      GaussianImage.intrinsic_matrix / extrinsic_matrix / projection                 (splat/image.py:32-40,:68-70)
      GaussianScene.render_points_image                                             (splat/gaussian_scene.py:44-51)
      GaussianScene.get_2d_covariance on the in-view points                         (:53-68)
      GaussianScene.render_tile on the list of one tile, aligned and unaligned      (:173-198)
      GaussianScene.render_pixel on the same list                                   (:146-171)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from intro_to_gaussian_splatting_b200.synth import make_scene  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402


def main():
    assert rh.reference_available(), "needs /root/reference"
    sc = make_scene("small")
    rs = rh.build_reference_scene(sc)
    im = rs.images[1]
    out = dict(intrinsic_matrix=im.intrinsic_matrix.numpy(), extrinsic_matrix=im.extrinsic_matrix.numpy(),
               projection=im.projection.numpy())
    with torch.no_grad():
        pts, cols = rs.render_points_image(1)
        out["points_image"], out["points_image_colors"] = pts.numpy(), cols.numpy()
        g = rs.gaussians
        keep = rh.load_reference().utils.in_view_frustum(points=g.points, view_matrix=im.world2view)
        cov3 = g.get_3d_covariance_matrix()[keep]
        out["cov2d"] = rs.get_2d_covariance(1, g.points[keep], cov3).numpy()
        pp = rh.reference_preprocess(rs, 1)
        # the list render_image would hand to render_tile for the tile at (64, 32)
        x_min, y_min, T = 64, 32, 16
        mask = ((pp.min_x <= x_min + T) & (pp.max_x >= x_min) & (pp.min_y <= y_min + T) & (pp.max_y >= y_min))
        rows = torch.nonzero(mask)[:, 0][:40]  # a prefix keeps the Python loop short; still a valid depth-ordered list
        args = dict(points_in_tile_mean=pp.points[rows], colors=pp.colors[rows], opacities=pp.sigmoid_opacity[rows],
                    inverse_covariance=pp.inverse_covariance_2d[rows])
        out["tile_rows"] = rows.numpy()
        out["tile_aligned"] = rs.render_tile(x_min=x_min, y_min=y_min, tile_size=T, **args).numpy()
        out["tile_unaligned"] = rs.render_tile(x_min=x_min + 5, y_min=y_min + 3, tile_size=T, **args).numpy()
        out["tile_size8"] = rs.render_tile(x_min=x_min, y_min=y_min, tile_size=8, **args).numpy()
        out["pixel_70_40"] = rs.render_pixel(pixel_coords=torch.Tensor([70, 40]).view(1, 2), **args).numpy()
        out["pixel_70_40_minw"] = rs.render_pixel(pixel_coords=torch.Tensor([70, 40]).view(1, 2), min_weight=0.5, **args).numpy()
    np.savez_compressed(os.path.join(HERE, "api_small.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
