#!/usr/bin/env python
"""Where does the end-to-end figure lose against `value` and against the copy ceiling?  One GPU, config 3, the
pipelined loop of bench.py with the egress done different ways (the library's own, or re-built here out of
Rasterizer.render into device staging images + torch copies), so that each ingredient can be switched separately.

    python tools/e2e_probe.py [--config cfg3] [--steps 60] [--repeats 5] [--inflight 3]
"""
import argparse
import json
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--repeats", type=int, default=5)
    ap.add_argument("--inflight", type=int, default=3)
    ap.add_argument("--modes", default="serial,indep,value,lib,value")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    spec = CONFIGS[a.config]
    K, F, R = a.steps, a.inflight, a.repeats
    sc = make_scene(spec, n_views=256)
    d = tempfile.mkdtemp()
    write_colmap_text(sc, d)
    cf, imf = read_camera_file(d), read_image_file(d)
    cams = [GaussianImage(cf[imf[i].camera_id], imf[i]).pack() for i in sorted(imf)]
    arrays = [t.to(dev) for t in (sc.xyz, sc.scales, sc.quats, (sc.rgb255 / 256).float(), sc.opacity_logit)]
    rasts = [Rasterizer(0) for _ in range(F)]
    for r in rasts:
        r.upload(*arrays)
    streams = [torch.cuda.Stream(device=dev) for _ in range(F)]
    copy_streams = [torch.cuda.Stream(device=dev) for _ in range(F)]
    shared_copy = torch.cuda.Stream(device=dev)
    H, W = spec.height, spec.width
    D = 3  # device staging images per context (modes use 1, 2 or 3 of them)
    stage = [[torch.empty((H, W, 3), dtype=torch.float32, device=dev) for _ in range(D)] for _ in range(F)]
    hosts = [[torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(D)] for _ in range(F)]
    prm = _lib.default_params(full_cover=1)
    prm_a = _lib.default_params(full_cover=1, async_host_copy=1)
    views = list(range(K))
    for v in views[:8]:
        for f in range(F):
            rasts[f].render(cams[v], prm, out=stage[f][0])
            rasts[f].render(cams[v], prm_a, out=hosts[f][0])
            rasts[f].join_host_copies()
    torch.cuda.synchronize()

    def loop(mode):
        main_s = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        depth = {"value": 1, "alt": 2, "lib": 2, "own": 2, "shared": 2, "own3": 3, "shared3": 3}[mode]
        copied = [[None] * depth for _ in range(F)]
        torch.cuda.synchronize()
        e0.record(main_s)
        for st in streams + copy_streams + [shared_copy]:
            st.wait_event(e0)
        for s, v in enumerate(views):
            f, b = s % F, (s // F) % depth
            with torch.cuda.stream(streams[f]):
                if mode == "lib":
                    rasts[f].render(cams[v], prm_a, out=hosts[f][b])
                    continue
                if copied[f][b] is not None:
                    streams[f].wait_event(copied[f][b])
                rasts[f].render(cams[v], prm, out=stage[f][b])
                if mode in ("value", "alt"):
                    continue
                rendered = torch.cuda.Event()
                rendered.record(streams[f])
            cs = shared_copy if mode.startswith("shared") else copy_streams[f]
            cs.wait_event(rendered)
            with torch.cuda.stream(cs):
                hosts[f][b].copy_(stage[f][b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            copied[f][b] = ev
        for f in range(F):
            with torch.cuda.stream(streams[f]):
                if mode == "lib":
                    rasts[f].join_host_copies()
                done = torch.cuda.Event()
                done.record(streams[f])
            main_s.wait_event(done)
        for cs in copy_streams + [shared_copy]:
            done = torch.cuda.Event()
            done.record(cs)
            main_s.wait_event(done)
        e1.record(main_s)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def independent(kind="d2h"):
        """K renders and K copies with NO dependency between them (the copies read a resident image): does either
        side slow the other down?  Returns (ms until the renders are done, ms until the copies are done)."""
        main_s = torch.cuda.current_stream(dev)
        e0 = torch.cuda.Event(enable_timing=True)
        er = [torch.cuda.Event(enable_timing=True) for _ in range(F)]
        ec = [torch.cuda.Event(enable_timing=True) for _ in range(F)]
        torch.cuda.synchronize()
        e0.record(main_s)
        for st in streams + copy_streams:
            st.wait_event(e0)
        for s, v in enumerate(views):
            f = s % F
            with torch.cuda.stream(streams[f]):
                rasts[f].render(cams[v], prm, out=stage[f][0])
            with torch.cuda.stream(copy_streams[f]):
                if kind == "d2h":
                    hosts[f][s // F % 2].copy_(stage[f][2], non_blocking=True)
                elif kind == "h2d":
                    stage[f][2].copy_(hosts[f][s // F % 2], non_blocking=True)
                else:
                    stage[f][1].copy_(stage[f][2], non_blocking=True)
        for f in range(F):
            er[f].record(streams[f])
            ec[f].record(copy_streams[f])
        torch.cuda.synchronize()
        return max(e0.elapsed_time(e) for e in er), max(e0.elapsed_time(e) for e in ec)

    def serial_stages(background):
        """Serial frames with stage times while `background` (None, 'd2h', 'h2d', 'd2d') copies run on another stream."""
        prm_t = _lib.default_params(full_cover=1, collect_stage_times=1)
        acc, n = {}, 0
        torch.cuda.synchronize()
        for v in views[:24]:
            if background:
                with torch.cuda.stream(copy_streams[0]):
                    for _ in range(4):  # ~1.8 ms of copies queued ahead of a 0.5 ms frame
                        if background == "d2h":
                            hosts[0][0].copy_(stage[0][2], non_blocking=True)
                        elif background == "h2d":
                            stage[0][2].copy_(hosts[0][0], non_blocking=True)
                        else:
                            stage[0][1].copy_(stage[0][2], non_blocking=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rasts[1].render(cams[v], prm_t, out=stage[1][0])
            e1.record()
            torch.cuda.synchronize()
            st = rasts[1].stage_times()
            st["frame"] = e0.elapsed_time(e1)
            for k_, x in st.items():
                acc[k_] = acc.get(k_, 0.0) + x
            n += 1
        return {k_: round(1e3 * x / n, 1) for k_, x in acc.items() if x > 0}

    out = {}
    if "serial" in a.modes.split(","):
        for bg in (None, "d2h", "h2d", "d2d", None):
            r_ = serial_stages(bg)
            print(f"serial frames, background copies {bg}: {r_} us", flush=True)
            out.setdefault("serial_" + str(bg), []).append(r_)
    for kind in ("d2h", "h2d", "d2d"):
        if "indep" not in a.modes.split(","):
            break
        independent(kind)
        rs = [independent(kind) for _ in range(R)]
        r_ms = sorted(x[0] for x in rs)[R // 2]
        c_ms = sorted(x[1] for x in rs)[R // 2]
        print(f"indep {kind}: renders {1e3 * K / r_ms:.1f} frames/s, copies {1e3 * K / c_ms:.1f} frames/s "
              f"({K * H * W * 12 / c_ms / 1e6:.1f} GB/s) when both run at once, no dependency", flush=True)
        out["indep_" + kind] = [round(1e3 * K / r_ms, 1), round(1e3 * K / c_ms, 1)]
    for mode in [m for m in a.modes.split(",") if m not in ("indep", "serial")]:
        loop(mode)
        ms = sorted(loop(mode) for _ in range(R))
        fps = 1e3 * K / ms[len(ms) // 2]
        out.setdefault(mode, []).append(round(fps, 1))
        print(f"{mode:8s} {fps:8.1f} frames/s  (min {1e3 * K / ms[-1]:.1f} max {1e3 * K / ms[0]:.1f})", flush=True)
    print(json.dumps({"e2e_probe": out, "config": a.config, "steps": K, "inflight": F}))


if __name__ == "__main__":
    main()
