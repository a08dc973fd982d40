"""View sharding on CPU: world_size-2 gloo processes (the N>1 host logic; the kernels need a GPU)."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from intro_to_gaussian_splatting_b200.sharding import (ViewShard, allreduce_gradients, broadcast_gaussians,
                                                      gather_frames)
from intro_to_gaussian_splatting_b200.synth import make_scene


def test_view_assignment_partitions_the_orbit():
    for world in (1, 2, 3, 4, 8):
        seen = []
        for r in range(world):
            sh = ViewShard(world, r, 256)
            mine = sh.my_views()
            assert all(sh.owner_of(v) == r for v in mine)
            assert [sh.view_of_step(s) for s in range(len(mine))] == mine
            seen += mine
        assert sorted(seen) == list(range(256))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1000
        arrays = None
        if rank == 0:
            sc = make_scene("small", n_override=n)
            arrays = [sc.xyz, sc.scales, sc.quats, (sc.rgb255 / 256).float(), sc.opacity_logit]
        got = broadcast_gaussians(arrays, n, torch.device("cpu"), world, rank)
        ref = make_scene("small", n_override=n)
        want = [ref.xyz, ref.scales, ref.quats, (ref.rgb255 / 256).float(), ref.opacity_logit]
        ok = all(torch.equal(a, b) for a, b in zip(got, want)) and [tuple(a.shape) for a in got] == [
            (n, 3), (n, 3), (n, 4), (n, 3), (n, 1)]
        sh = ViewShard(world, rank, 7)
        frames = [torch.full((2, 2, 3), float(v)) for v in sh.my_views()]
        allf = gather_frames(frames, sh)
        if rank == 0:
            ok = ok and [float(f[0, 0, 0]) for f in allf] == [float(v) for v in range(7)]
        # the training exchange: five gradient arrays (one of them absent), one packed all-reduce, mean over ranks
        shapes = [(n, 3), (n, 3), (n, 4), (n, 3), (n, 1)]
        grads = [torch.full(sh_, float(rank + 1) * (i + 1)) for i, sh_ in enumerate(shapes)]
        grads[1] = None
        allreduce_gradients(grads, world)
        for i, g in enumerate(grads):
            if g is not None:
                ok = ok and tuple(g.shape) == shapes[i] and bool((g == 1.5 * (i + 1)).all())
        summed = [torch.full((4, 2), float(rank + 1))]
        allreduce_gradients(summed, world, average=False)
        ok = ok and bool((summed[0] == 3.0).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}
