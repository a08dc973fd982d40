"""View-sharded data-parallel training check (run under torchrun, one rank per GPU):
   every rank starts from the same perturbed Gaussians, differentiates its own views, gradients are averaged with
   one packed NCCL all-reduce per step.  Prints one line per rank; exits non-zero on failure."""
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import helpers  # noqa: E402
from intro_to_gaussian_splatting_b200 import Rasterizer, _lib, fit  # noqa: E402
from intro_to_gaussian_splatting_b200.sharding import ViewShard  # noqa: E402
from intro_to_gaussian_splatting_b200.synth import SceneSpec  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
spec = SceneSpec("grad_dense", 260, 64, 48, box=2.5, log_scale_range=(-3.6, -1.6))
sc, images, _ = helpers.scene_and_images(spec, n_views=4)
cams = [images[i].pack() for i in sorted(images)]
prm = _lib.default_params(full_cover=1)
pts, scl, qts, col, opa = helpers.scene_arrays(sc)
r = Rasterizer(local)
r.upload(pts, scl, qts, col, opa)
targets = [r.render(c, prm).clone() for c in cams]
g = torch.Generator().manual_seed(1)  # same perturbation on every rank
gs = SimpleNamespace(points=pts.clone(), scales=scl.clone(), quaternions=qts.clone(),
                     colors=(col + 0.2 * torch.randn(col.shape, generator=g)).clamp(0, 1),
                     opacity=opa + 0.5 * torch.randn(opa.shape, generator=g))
hist = fit(gs, cams, targets, steps=40, params=prm, rasterizer=r, lr={"colors": 2e-2, "opacity": 5e-2},
           trainable=("colors", "opacity"), shard=ViewShard(world, rank, len(cams)))
# all ranks must hold identical parameters after training
ref = gs.colors.detach().clone()
dist.broadcast(ref, src=0)
same = bool(torch.equal(ref, gs.colors.detach()))
first, last = sum(hist[:2]) / 2, sum(hist[-2:]) / 2
print(f"rank {rank}/{world}: loss {first:.3e} -> {last:.3e}, parameters identical across ranks: {same}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (same and last < 0.5 * first) else 1)
