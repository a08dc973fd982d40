#!/usr/bin/env python
"""bench.py -- frames/s of the forward Gaussian-splat render path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one frame: projection -> tile binning -> radix sort -> tile ranges -> compositing of the
whole Gaussian set for one camera of the 256-view orbit.  Workload at N=1 is
BASELINE.json configs[2] ("cfg3": 1 M synthetic Gaussians, 1920x1080), the configuration the metric is
quoted on.  Multi-GPU is view-sharded: the Gaussian set is broadcast once with NCCL, rank r renders
views r, r+R, ... with no per-frame collective (weak scaling: K frames per rank).

Prints ONE JSON line (rank 0).  Timing rules followed: W >= 3 warm-up frames (+ one untimed sweep over
the timed views so no scratch buffer grows inside the timed region); L2 flushed (256 MiB memset)
between timed frames; per-frame CUDA events on the launch stream; max over ranks; clocks sampled from
nvidia-smi during the timed region.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "frames/s at 1080p, 1M Gaussians (view-sharded)"
UNIT = "frames/s"
ORBIT = 256


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--full-cover", type=int, default=1)
    ap.add_argument("--sort-mode", default="auto", choices=["auto", "full", "split"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=3,
                    help="frames in flight per GPU (one rasterizer context + one CUDA stream each); 1 = strictly serial frames")
    ap.add_argument("--ref-budget-s", type=float, default=240.0)
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def workload_name(spec):
    return f"{spec.name}: {spec.n} synthetic Gaussians (synth-v2 seed {spec.seed}), {spec.width}x{spec.height}, {ORBIT}-view orbit"


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="gsb_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------------
def oracle_frame(orc, ocam, oprm, arrays):
    t0 = time.perf_counter()
    fr = orc.render(ocam, oprm, *arrays)
    return time.perf_counter() - t0, fr


def make_oracle_inputs(spec_name, full_cover, n_views):
    from intro_to_gaussian_splatting_b200.colmap_io import read_camera_file, read_image_file
    from intro_to_gaussian_splatting_b200.image import GaussianImage
    from intro_to_gaussian_splatting_b200.synth import make_scene, write_colmap_text
    from oracle import oracle as orc

    sc = make_scene(spec_name, n_views=n_views)
    d = tempfile.mkdtemp(prefix="gsb_bench_")
    write_colmap_text(sc, d)
    cams, imgs = read_camera_file(d), read_image_file(d)
    ocams = []
    for i in sorted(imgs):
        cam = GaussianImage(cams[imgs[i].camera_id], imgs[i]).pack()
        o = orc.Camera()
        C.memmove(C.byref(o), C.byref(cam), C.sizeof(cam))
        ocams.append(o)
    arrays = (sc.xyz, sc.scales, sc.quats, (sc.rgb255 / 256).float(), sc.opacity_logit)
    arrays = tuple(a.numpy() for a in arrays)
    # all host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its workers)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    orc.set_num_threads(cores)
    return sc, orc, ocams, orc.default_params(full_cover=full_cover), arrays


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port, OpenMP over all host cores) on the same
    workload.  The reference's own Python loop needs ~74 us per (pixel, Gaussian) step (428 s for
    10 k Gaussians at 256x256; ~45 h for this frame), so the C port is a far FASTER stand-in for it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_views = max(args.steps + args.warmup, 1)
    sc, orc, ocams, oprm, arrays = make_oracle_inputs(args.config, args.full_cover, min(n_views, ORBIT))
    cores = orc.num_threads()
    t_w = []
    for w in range(max(args.warmup, 1)):
        dt, _ = oracle_frame(orc, ocams[w % len(ocams)], oprm, arrays)
        t_w.append(dt)
    est = min(t_w)
    steps_measured = max(1, min(args.steps, int(args.ref_budget_s / max(est, 1e-6))))
    total = 0.0
    for s in range(steps_measured):
        dt, fr = oracle_frame(orc, ocams[(args.warmup + s) % len(ocams)], oprm, arrays)
        total += dt
    ms = 1e3 * total / steps_measured
    val = 1e3 / ms
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "steps_measured": steps_measured, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(sc.spec), "full_cover": args.full_cover, "tile_size": 16,
                   "semantics": "ref_cpu", "sample": "whole frames (projection+binning+sort+compositing), one orbit view per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps_measured} whole frames of the workload, C port of the reference CPU path "
                                   f"(oracle/gs_oracle.c, OpenMP x{cores})"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def stage_bytes(info, n, width, height, split):
    """Algorithmic HBM bytes per frame and stage.  N Gaussians, M in view, V of them with
    tiles (not in GsbFrameInfo: M is used, an upper bound), K tile instances, P pixels.  SPLIT: the tile passes
    move one key per instance -- 4 bytes (tile << rank bits | position) when it fits, else 8 (tile << 32 | index)
    -- and the last pass writes only the 4-byte Gaussian index."""
    N, M, K = n, int(info.m_in_view), int(info.k_instances)
    P = width * height
    tiles = info.tiles_x * info.tiles_y
    cells = (info.tiles_x + 1) * (info.tiles_y + 1)
    kb = 4 if int(getattr(info, "key_bits", 64)) == 32 else 8
    if split:
        sort = (info.sort_passes - 1) * 2 * kb * K + (kb + 4) * K   # key read + written per pass; last pass key in, index out
        emit = 24 * M + kb * K
    else:
        sort = info.sort_passes * 24 * K                     # (8+4 read, 8+4 written) per pass
        emit = 24 * M + 12 * K
    b = {
        "project": 56 * N + 4 * N + 4 * N + 56 * M,         # planes in; depth key + count for all, record + rect in view
        "depth_sort": info.depth_passes * 16 * N - 4 * N if split else 0,  # first pass reads keys only
        "scan": (12 if split else 8) * N,
        "emit": emit,
        "sort": sort,
        "ranges": 4 * cells + 8 * tiles,                     # tile_stats: difference grid in, ranges out
        "composite": 52 * K + 12 * P,                        # upper bound: lists are cut short by early termination
    }
    return b


def bind_to_gpu_numa_node(device_index):
    """Multi-rank runs: pin this process to the CPUs NVML reports as local to its GPU, so that the pinned host
    images of the e2e leg are first-touched on the NUMA node behind the GPU's PCIe root (torchrun does not bind
    ranks).  Returns the number of CPUs bound to, or None when NVML / affinity is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from intro_to_gaussian_splatting_b200 import Rasterizer, _lib
    from intro_to_gaussian_splatting_b200.colmap_io import read_camera_file, read_image_file
    from intro_to_gaussian_splatting_b200.image import GaussianImage
    from intro_to_gaussian_splatting_b200.sharding import ViewShard, broadcast_gaussians
    from intro_to_gaussian_splatting_b200.synth import CONFIGS, make_scene, write_colmap_text

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1:
        numa = bind_to_gpu_numa_node(local)  # before any pinned allocation: the e2e images are copied to host memory
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    spec = CONFIGS[args.config]

    # cameras: every rank derives the same orbit; Gaussians: made on rank 0, broadcast once over NCCL
    sc_cams = make_scene(spec, n_views=ORBIT, n_override=1)
    d = tempfile.mkdtemp(prefix="gsb_bench_")
    write_colmap_text(sc_cams, d)
    cams_f, imgs_f = read_camera_file(d), read_image_file(d)
    cams = [GaussianImage(cams_f[imgs_f[i].camera_id], imgs_f[i]).pack() for i in sorted(imgs_f)]
    if rank == 0:
        sc = make_scene(spec, n_views=1)
        arrays = [sc.xyz, sc.scales, sc.quats, (sc.rgb255 / 256).float(), sc.opacity_logit]
    else:
        arrays = None
    t0 = time.perf_counter()
    arrays = broadcast_gaussians(arrays, spec.n, dev, world, rank)
    torch.cuda.synchronize()
    bcast_s = time.perf_counter() - t0

    F = max(1, args.inflight)
    rasts = [Rasterizer(local) for _ in range(F)]  # independent contexts: own scratch, own aux/copy streams
    for r_ in rasts:
        r_.upload(*arrays)
    rast = rasts[0]
    streams = [torch.cuda.Stream(device=dev) for _ in range(F)]
    sort_mode = {"auto": _lib.GSB_SORT_AUTO, "full": _lib.GSB_SORT_FULL, "split": _lib.GSB_SORT_SPLIT}[args.sort_mode]
    prm = _lib.default_params(full_cover=args.full_cover, sort_mode=sort_mode)
    prm_t = _lib.default_params(full_cover=args.full_cover, sort_mode=sort_mode, collect_stage_times=1)
    prm_a = _lib.default_params(full_cover=args.full_cover, sort_mode=sort_mode, async_host_copy=1)
    shard = ViewShard(world, rank, ORBIT)
    views = [shard.view_of_step(s) for s in range(K)]
    H, Wd = spec.height, spec.width
    img = torch.empty((H, Wd, 3), dtype=torch.float32, device=dev)
    imgs = [img] + [torch.empty_like(img) for _ in range(F - 1)]
    hosts = [torch.empty((H, Wd, 3), dtype=torch.float32).pin_memory() for _ in range(2 * F)]
    flushes = [torch.empty(256 << 20, dtype=torch.uint8, device=dev) for _ in range(F)]
    flush = flushes[0]

    # warm-up (>= 3 frames) + one untimed sweep over the timed views: scratch reaches its final size in every context
    for s in range(W):
        for f in range(F):
            rasts[f].render(cams[shard.view_of_step(K + s)], prm, out=imgs[f])
    for s, v in enumerate(views):
        for f in range(F):
            rasts[f].render(cams[v], prm, out=imgs[f])
            rasts[f].render(cams[v], prm_a, out=hosts[f])
            rasts[f].join_host_copies()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def pipelined_loop(outs, params, join, do_flush=False):
        """K frames, F in flight: frame s runs in context s % F on stream s % F (each context holds its OWN copy of the
        Gaussian set, so with F >= 3 the inputs cycled through are larger than L2: 3 x 56 MB > 126 MB); one event pair
        on the main stream brackets the whole loop (the side streams fork from the start event and are joined before
        the end event).  do_flush additionally writes 256 MiB before every frame INSIDE the timed region.
        Returns total ms, kernel launches, frame infos."""
        main = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        done = [torch.cuda.Event() for _ in range(F)]
        launches = 0
        infos = []
        barrier()
        e0.record(main)
        for st_ in streams:
            st_.wait_event(e0)
        for s, v in enumerate(views):
            f = s % F
            with torch.cuda.stream(streams[f]):
                if do_flush:
                    flushes[f].zero_()
                rasts[f].render(cams[v], params, out=outs[s % len(outs)])
            info = rasts[f].frame_info()
            launches += info.kernel_launches
            infos.append(info)
        for f in range(F):
            with torch.cuda.stream(streams[f]):
                if join:
                    rasts[f].join_host_copies()
                done[f].record(streams[f])
            main.wait_event(done[f])
        e1.record(main)
        barrier()
        return e0.elapsed_time(e1), launches, infos

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    explicit_flush = F < 3  # fewer than 3 scene copies do not exceed L2: fall back to flushing inside the region
    pipelined_loop(imgs, prm, join=False)                             # untimed pass: clocks in steady state
    ms_dev, launches, infos = pipelined_loop(imgs, prm, join=False, do_flush=explicit_flush)
    ms_e2e, _, _ = pipelined_loop(hosts, prm_a, join=True, do_flush=explicit_flush)
    pipelined_loop(imgs, prm, join=False, do_flush=True)                       # untimed pass of the flush variant
    ms_dev_flush, _, _ = pipelined_loop(imgs, prm, join=False, do_flush=True)  # same loop, 256 MiB written per frame
    clocks = sampler.stop() if rank == 0 else None

    # single-frame latency (serial frames, flush outside the events), for reference next to the throughput
    lat_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for s, v in enumerate(views):
        flush.zero_()
        lat_ev[s][0].record()
        rast.render(cams[v], prm, out=img)
        lat_ev[s][1].record()
    torch.cuda.synchronize()
    frame_latency_ms = sum(a.elapsed_time(b) for a, b in lat_ev) / K

    # the training step, outside the headline's timed region: forward with
    # save_for_backward + gsb_render_backward for a random dL/d image, serial frames, L2 flushed before each
    prm_b = _lib.default_params(full_cover=args.full_cover, sort_mode=prm.sort_mode, save_for_backward=1)
    gimg = torch.randn_like(img)
    nb_ = min(K, 20)
    tr_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(nb_)]
    for s in range(-2, nb_):  # two untimed iterations size the gradient scratch
        v = views[s % len(views)]
        flush.zero_()
        e = tr_ev[max(s, 0)]
        e[0].record()
        rast.render(cams[v], prm_b, out=img)
        e[1].record()
        rast.render_backward(cams[v], prm_b, gimg)
        e[2].record()
    torch.cuda.synchronize()
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in tr_ev) / nb_
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in tr_ev) / nb_

    # per-stage times for the roofline (separate loop: the event pairs add a little overhead)
    stage_ms = {k: 0.0 for k in _lib.STAGE_NAMES}
    stage_ms_first = None
    for v in views:
        flush.zero_()
        rast.render(cams[v], prm_t, out=img)
        st_v = rast.stage_times()
        if stage_ms_first is None:
            stage_ms_first = st_v
        for k, t in st_v.items():
            stage_ms[k] += t
    stage_ms = {k: t / K for k, t in stage_ms.items()}

    t = torch.tensor([ms_dev, ms_e2e, ms_dev_flush], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev_max, ms_e2e_max, ms_flush_max = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, sm_max_mhz, peak_src = load_peaks()
    split = infos[0].depth_passes > 0
    # stage accounting averaged over this rank's timed views
    sb = {k: 0.0 for k in _lib.STAGE_NAMES}
    for info in infos:
        for k, v in stage_bytes(info, spec.n, Wd, H, split).items():
            sb[k] += v / len(infos)
    stages = []
    for k in _lib.STAGE_NAMES:
        if stage_ms[k] <= 0:
            continue
        gbs = sb[k] / (stage_ms[k] * 1e-3) / 1e9
        stages.append({"stage": k, "ms": round(stage_ms[k], 4), "alg_mb": round(sb[k] / 1e6, 2),
                       "achieved_gbs": round(gbs, 1), "frac_hbm": round(gbs / hbm_peak, 4)})
    dom = max(stages, key=lambda s: s["ms"])
    kernel_of = {"project": "project_kernel", "depth_sort": "onesweep_kernel<u32> x4", "scan": "scan_kernel",
                 "emit": "emit_kernel (+ host read-back of K)", "sort": "onesweep_kernel<u64>", "ranges": "tile_stats_kernel",
                 "composite": "composite_fast_kernel"}
    traffic = None
    try:  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture of this config
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_roofline_traffic.json")))
        if args.full_cover == 1 and split:
            traffic = tj.get(args.config, {}).get(dom["stage"])
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": kernel_of[dom["stage"]], "achieved": dom["achieved_gbs"], "peak": hbm_peak,
                "unit": "GB/s", "frac": dom["frac_hbm"], "traffic": traffic, "peak_source": peak_src,
                "launch_ms": dom["ms"], "stages": stages}
    if dom["stage"] == "composite":
        roofline["note"] = ("compositing is issue-slot bound (fp32 + MUFU.EX2), not HBM bound; the HBM fraction is "
                            "reported for the schema, the issue-slot figure is in `issue`")

    mean_k = sum(int(i.k_instances) for i in infos) / len(infos)
    mean_m = sum(int(i.m_in_view) for i in infos) / len(infos)
    line = {
        "metric": METRIC, "value": K * world / (ms_dev_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(spec), "full_cover": args.full_cover, "tile_size": 16, "semantics": "ref_cpu",
                   "sort_mode": "split" if split else "full", "views_per_rank": K, "parallelism": f"view-sharded x{world}",
                   "frames_in_flight": F,
                   "l2": ("256 MiB written before every frame, inside the timed region" if explicit_flush else
                          f"inputs larger than L2: {F} contexts, each with its own {56 * spec.n / 1e6:.0f} MB copy of the "
                          f"Gaussian set, used round-robin ({F * 56 * spec.n / 1e6:.0f} MB > 126 MB L2); each frame also "
                          "streams ~0.4 GB of intermediates; no explicit flush"),
                   "value_with_flush_inside_timed_region": round(K * world / (ms_flush_max * 1e-3), 1),
                   "frame_latency_ms_serial": round(frame_latency_ms, 4),
                   "mean_in_view": round(mean_m), "mean_tile_instances": round(mean_k),
                   "gaussian_broadcast_s": round(bcast_s, 4), "cpus_bound_per_rank": numa},
        "e2e": {"value": K * world / (ms_e2e_max * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_max / K,
                "h2d_bytes_per_step": C.sizeof(_lib.GsbCamera) + C.sizeof(_lib.GsbParams),
                "d2h_bytes_per_step": H * Wd * 3 * 4 + 8,
                "note": "gsb_render with host structs in and a pinned host image out (async egress on the copy stream, "
                        "joined before the end event); same pipelined loop as `value`; Gaussians stay resident like "
                        "model weights"},
        "gpu_launches": launches,
        "launches_per_step": launches / K,
        "roofline": roofline,
        "clocks": clocks,
        "train_step": {"forward_ms": round(fwd_ms, 4), "backward_ms": round(bwd_ms, 4),
                       "steps_per_s": round(1e3 / (fwd_ms + bwd_ms), 1), "frames": nb_,
                       "note": "gsb_render(save_for_backward) + gsb_render_backward, serial frames, rank 0, not part "
                               "of `value`; gradients wrt all five attribute tensors"},
    }

    if world == 1 and not args.no_cpu_baseline:
        # the oracle port, timed on this box's host cores on ONE frame of the same workload (view 0)
        try:
            sc0, orc, ocams, oprm, oarrays = make_oracle_inputs(args.config, args.full_cover, 1)
            dt, fr = oracle_frame(orc, ocams[0], oprm, oarrays)
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                                    "sample": "1 whole frame (orbit view 0) of the workload through oracle/gs_oracle.c",
                                    "steps_executed": int(fr.steps)}
            comp_ms = stage_ms_first.get("composite", 0.0)  # views[0] is orbit view 0 on rank 0
            if comp_ms > 0 and clocks and clocks.get("sm_mhz"):
                # issue-slot roofline of compositing.  0.580 warp-instructions per executed (pixel, Gaussian) step
                # is the ncu count for this kernel on this view (smsp__inst_executed.sum / oracle step count,
                # profiles/r1_summary.md); the peak is 4 schedulers x 148 SMs x the SM clock sampled during the run.
                warp_inst = 0.580 * fr.steps
                peak = 148 * 4 * clocks["sm_mhz"] * 1e6
                roofline["issue"] = {"kernel": "composite_fast_kernel", "steps_view0": int(fr.steps),
                                     "composite_ms_view0": comp_ms,
                                     "warp_inst_per_s": warp_inst / (comp_ms * 1e-3),
                                     "peak_warp_inst_per_s": peak, "frac": warp_inst / (comp_ms * 1e-3) / peak,
                                     "note": "warp instructions = 0.580 x executed pixel-steps (ncu-calibrated, profiles/r1_summary.md)"}
        except Exception as e:  # the baseline must never take the bench line down
            line["cpu_baseline"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
