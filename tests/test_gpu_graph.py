"""GPU tests of the frame submitted as ONE CUDA graph (csrc/gsb_api.cu: FrameCapture).

On any stream but the legacy default one `gsb_render` captures its own launch sequence, updates the instantiated
graph in place and launches it once.  Everything the launch-by-launch path guarantees must hold: lists bit-exact
against the oracle, pixels bit-identical to the launch-by-launch frame, the tail re-queued when a count outgrows the
capacities, topology changes (image size, pass count, saved state) between frames, the backward pass, the
asynchronous egress; and the default stream / GSB_GRAPH=0 must keep queueing launch by launch.
"""

import os

import numpy as np
import pytest
import torch

from helpers import scene_and_images, scene_arrays, to_oracle_camera, to_oracle_params, u64
from intro_to_gaussian_splatting_b200 import Rasterizer, _lib
from intro_to_gaussian_splatting_b200.synth import SceneSpec
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
PIXEL_TOL = 1e-4


def _check_lists(rast, fr):
    info = rast.frame_info()
    assert info.m_in_view == fr.proj.m and info.k_instances == fr.keys.shape[0]
    keys, payload = rast.debug_sorted_keys()
    assert np.array_equal(u64(keys), fr.sorted_keys), "sorted keys differ"
    assert np.array_equal(payload.cpu().numpy().view(np.uint32), fr.sorted_payload), "sorted payload differs"
    assert np.array_equal(rast.debug_tile_ranges().cpu().numpy().view(np.uint32), fr.ranges), "tile ranges differ"


@pytest.mark.parametrize("name,full_cover", [("cfg2", 1), ("small", 0), ("cfg1", 0)])
def test_graph_frames_match_oracle_and_launch_by_launch_frames(name, full_cover):
    sc, images, _ = scene_and_images(name, n_views=2)
    arrs = scene_arrays(sc)
    prm = _lib.default_params(full_cover=full_cover)
    r = Rasterizer(0)
    r.upload(*[a.cuda() for a in arrs])
    side = torch.cuda.Stream()
    for idx in (1, 2, 1):
        cam = images[idx].pack()
        direct = r.render(cam, prm).clone()  # default stream: launch by launch
        assert r.frame_info().graph_launch == 0
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            img = r.render(cam, prm)
            info = r.frame_info()
            assert info.graph_launch == 1 and info.kernel_launches >= 9
        side.synchronize()
        assert torch.equal(img, direct), "the graph frame differs from the launch-by-launch frame"
        fr = orc.render(to_oracle_camera(cam), to_oracle_params(prm), *arrs)
        with torch.cuda.stream(side):
            _check_lists(r, fr)
        assert np.abs(img.cpu().numpy() - fr.image).max() <= PIXEL_TOL
    r.close()


def test_graph_frame_whose_counts_outgrow_the_capacities():
    """First frame of a fresh context on a side stream: the graph runs with the abort flag set, the host grows the
    buffers and queues the tail again (launch by launch); the second frame is a plain graph frame."""
    spec = SceneSpec("huge", 300, 640, 400, box=1.0, log_scale_range=(-1.5, 0.0))
    sc, images, _ = scene_and_images(spec)
    cam = images[1].pack()
    arrs = scene_arrays(sc)
    prm = _lib.default_params(full_cover=1)
    fr = orc.render(to_oracle_camera(cam), to_oracle_params(prm), *arrs)
    r = Rasterizer(0)
    r.upload(*[a.cuda() for a in arrs])
    side = torch.cuda.Stream()
    requeued = []
    with torch.cuda.stream(side):
        for _ in range(3):
            img = r.render(cam, prm)
            info = r.frame_info()
            requeued.append(info.tail_requeued)
            assert info.graph_launch == 1
            side.synchronize()
            _check_lists(r, fr)
            assert np.abs(img.cpu().numpy() - fr.image).max() <= PIXEL_TOL
    assert requeued[0] == 1 and requeued[-1] == 0
    r.close()


def test_graph_survives_changes_of_topology_between_frames():
    """Image size (another tile grid, another number of radix passes over the super-tile ids), full_cover (a memset
    node more), save_for_backward (another compositing kernel), number of Gaussians: the instantiated graph is updated
    when it can be and rebuilt when it cannot; every frame is right."""
    side = torch.cuda.Stream()
    r = Rasterizer(0)
    wide = SceneSpec("wide2k", 20_000, 2000, 1200, log_scale_range=(-5.5, -2.5))
    seq = [("small", 1, 0), (wide, 1, 0), ("small", 0, 0), ("small", 0, 1), ("cfg2", 1, 1), ("small", 1, 0), (wide, 0, 0)]
    for name, fc, save in seq:
        sc, images, _ = scene_and_images(name)
        arrs = scene_arrays(sc)
        cam = images[1].pack()
        prm = _lib.default_params(full_cover=fc, save_for_backward=save)
        fr = orc.render(to_oracle_camera(cam), to_oracle_params(_lib.default_params(full_cover=fc)), *arrs)
        with torch.cuda.stream(side):
            r.upload(*[a.cuda() for a in arrs])
            for _ in range(2):
                img = r.render(cam, prm)
                assert r.frame_info().graph_launch == 1
            side.synchronize()
            _check_lists(r, fr)
            assert np.abs(img.cpu().numpy() - fr.image).max() <= PIXEL_TOL
    r.close()


def test_graph_frame_feeds_the_backward_pass():
    sc, images, _ = scene_and_images("small")
    arrs = [a.cuda() for a in scene_arrays(sc)]
    cam = images[1].pack()
    prm = _lib.default_params(full_cover=1, save_for_backward=1)
    r = Rasterizer(0)
    r.upload(*arrs)
    g_img = torch.rand((cam.height, cam.width, 3), device="cuda")
    r.render(cam, prm)
    ref = r.render_backward(cam, prm, g_img, frame_id=r.last_frame_id)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        r.render(cam, prm)
        assert r.frame_info().graph_launch == 1
        got = r.render_backward(cam, prm, g_img, frame_id=r.last_frame_id)
    side.synchronize()
    for k in ref:
        # atomics accumulate in another order from run to run: same tolerance as tests/test_gpu_backward.py's rerun check
        assert torch.allclose(got[k], ref[k], rtol=1e-4, atol=1e-6), k
    r.close()


def test_graph_frames_with_asynchronous_egress():
    sc, images, _ = scene_and_images("cfg2", n_views=4)
    arrs = scene_arrays(sc)
    r = Rasterizer(0)
    r.upload(*[a.cuda() for a in arrs])
    prm = _lib.default_params(full_cover=1)
    prm_a = _lib.default_params(full_cover=1, async_host_copy=1)
    cams = [images[i].pack() for i in (1, 2, 3, 4)]
    want = [r.render(c, prm).cpu() for c in cams]
    hosts = [torch.zeros_like(w).pin_memory() for w in want]
    hosts_u8 = [torch.zeros(w.shape, dtype=torch.uint8).pin_memory() for w in want]
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for c, h in zip(cams, hosts):
            r.render(c, prm_a, out=h)
            assert r.frame_info().graph_launch == 1
        for c, h in zip(cams, hosts_u8):
            r.render(c, prm_a, out=h, layout="u8")
        r.join_host_copies()
    side.synchronize()
    for w, h, h8 in zip(want, hosts, hosts_u8):
        assert torch.equal(w, h)
        assert torch.equal((w.clamp(0, 1) * 255).to(torch.uint8), h8) or \
            (h8.int() - (w.clamp(0, 1) * 255).int()).abs().max() <= 1
    r.close()


def test_graph_can_be_switched_off_and_timing_frames_stay_launch_by_launch():
    sc, images, _ = scene_and_images("small")
    arrs = scene_arrays(sc)
    cam = images[1].pack()
    side = torch.cuda.Stream()
    old = os.environ.get("GSB_GRAPH")
    os.environ["GSB_GRAPH"] = "0"
    try:
        r0 = Rasterizer(0)
    finally:
        if old is None:
            os.environ.pop("GSB_GRAPH", None)
        else:
            os.environ["GSB_GRAPH"] = old
    r1 = Rasterizer(0)
    for r in (r0, r1):
        r.upload(*[a.cuda() for a in arrs])
    with torch.cuda.stream(side):
        a = r0.render(cam, _lib.default_params(full_cover=1))
        assert r0.frame_info().graph_launch == 0
        b = r1.render(cam, _lib.default_params(full_cover=1))
        assert r1.frame_info().graph_launch == 1
        t = r1.render(cam, _lib.default_params(full_cover=1, collect_stage_times=1)).clone()
        assert r1.frame_info().graph_launch == 0 and r1.stage_times()["composite"] > 0
        f = r1.render(cam, _lib.default_params(full_cover=1, sort_mode=_lib.GSB_SORT_FULL)).clone()
        assert r1.frame_info().graph_launch == 0
        b2 = r1.render(cam, _lib.default_params(full_cover=1))
        assert r1.frame_info().graph_launch == 1
    side.synchronize()
    assert torch.equal(a, b) and torch.equal(b, t) and torch.equal(b, f) and torch.equal(b, b2)
    r0.close()
    r1.close()


def test_render_inside_a_callers_capture_is_refused_not_hung():
    """A caller's own stream capture cannot contain gsb_render (the host waits for the frame's counts): the call must
    come back -- launch by launch it would spin on a mailbox nobody writes."""
    sc, images, _ = scene_and_images("small")
    arrs = scene_arrays(sc)
    cam = images[1].pack()
    r = Rasterizer(0)
    r.upload(*[a.cuda() for a in arrs])
    r.render(cam, _lib.default_params(full_cover=1))
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    out = torch.empty((cam.height, cam.width, 3), device="cuda")
    with pytest.raises(RuntimeError):
        with torch.cuda.graph(g, stream=side):
            r.render(cam, _lib.default_params(full_cover=1), out=out)
    torch.cuda.synchronize()
    img = r.render(cam, _lib.default_params(full_cover=1))  # the context is still usable
    torch.cuda.synchronize()
    assert torch.isfinite(img).all()
    r.close()
