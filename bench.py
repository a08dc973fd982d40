#!/usr/bin/env python
"""bench.py -- frames/s of the forward Gaussian-splat render path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one frame: projection -> depth sort -> super-tile binning -> compositing of the whole Gaussian set
for one camera of the 256-view orbit.  Workload at N=1 is BASELINE.json configs[2]
("cfg3": 1 M synthetic Gaussians, 1920x1080), the configuration the metric is quoted on.  Multi-GPU is
view-sharded: the Gaussian set is broadcast once with NCCL, rank r renders views r, r+R, ... with no per-frame
collective (weak scaling: K frames per rank).

Prints ONE JSON line (rank 0).  Timing rules followed: W >= 3 warm-up frames (+ one untimed sweep over the timed
views so no scratch buffer grows inside a timed region); the K-frame loop is timed `--repeats` times, each
repetition bracketed by barrier + synchronize and one CUDA-event pair on the launch stream, max over ranks per
repetition; `value` is the MEDIAN repetition, p10/p90 are in `config`; inputs cycled through exceed L2 (see
config.l2); clocks sampled from nvidia-smi during the timed region.  Nothing in a timed region is precomputed:
every frame is projected, sorted, binned and composited from the resident Gaussian set.
"""

from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "frames/s at 1080p, 1M Gaussians (view-sharded)"
UNIT = "frames/s"
ORBIT = 256
CALIBRATION = os.path.join(ROOT, "profiles", "r2_roofline_calibration.json")


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
    ap.add_argument("--repeats", type=int, default=10, help="how many times the K-step loop is timed (median reported)")
    ap.add_argument("--orbit", action="store_true",
                    help="strong-scaling orbit run: the 256 views of the orbit are split over the ranks (256/N frames per "
                         "rank, --steps ignored); used for the cfg4 record in profiles/, not by the driver")
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


def pct(xs, q):
    xs = sorted(xs)
    if not xs:
        return None
    i = min(len(xs) - 1, max(0, int(round(q * (len(xs) - 1)))))
    return xs[i]


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
def stage_bytes(info, n, width, height):
    """Algorithmic HBM bytes per frame and stage.  N Gaussians, M in view, V of them with tiles,
    Ks keys moved by the tile-level radix passes (super-tile instances in SPLIT mode, tile instances in FULL), K tile
    instances, P pixels."""
    N, M, K = n, int(info.m_in_view), int(info.k_instances)
    V = int(info.v_with_tiles) or M
    Ks = int(info.k_sorted)
    P = width * height
    tiles = info.tiles_x * info.tiles_y
    cells = (info.tiles_x + 1) * (info.tiles_y + 1)
    kb = 4 if int(info.key_bits) == 32 else 8
    split = info.depth_passes > 0
    if split:
        two_level = info.super_w * info.super_h > 1
        out_b = 8 if two_level else 4                        # last pass: list entry {index, tile mask} or bare index
        sort = (info.sort_passes - 1) * 2 * kb * Ks + (kb + out_b + (12 if two_level else 4)) * Ks  # + order / rect gathers
        emit = 16 * V + kb * Ks                              # offsets + order + rect per Gaussian, one key per instance
        scan = 12 * V + 4 * V                                # order + rect in, offsets out
        # compositing reads list entries until the tile saturates: upper bound = every entry of every super-tile list
        # for each of its tiles, plus one record per tile instance, plus the image
        comp = 8 * Ks * info.super_w * info.super_h + 48 * K + 12 * P if two_level else 52 * K + 12 * P
    else:
        sort = info.sort_passes * 24 * K                     # (8+4 read, 8+4 written) per pass
        emit = 24 * M + 12 * K
        scan = 8 * N
        comp = 52 * K + 12 * P
    return {
        "project": 56 * N + 4 * N + 4 * N + 8 * N + 48 * V,  # planes in; depth key + count + rect for all, record if drawn
        "depth_sort": info.depth_passes * 16 * N - 4 * N if split else 0,  # first pass reads keys only
        "scan": scan,
        "emit": emit,
        "sort": sort,
        "ranges": 4 * cells + 8 * tiles,                     # tile_stats: difference grid in, ranges out
        "composite": comp,                                   # upper bound: lists are cut short by early termination
        "expand": 0,
    }


def host_topology():
    """NUMA nodes the box exposes (the e2e leg writes images into host memory: where that memory lives matters)."""
    out = {"numa_nodes": None, "cpus": os.cpu_count()}
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        out["numa_nodes"] = len(nodes)
    except Exception:
        pass
    return out


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
    from intro_to_gaussian_splatting_b200.sharding import ViewShard, broadcast_gaussians, gather_frames
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
    W = max(args.warmup, 3)
    K = ORBIT // world if args.orbit else args.steps
    R = max(1, args.repeats)
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
    bcast_first_s = time.perf_counter() - t0  # includes NCCL communicator bring-up
    bcast_s = None
    if world > 1:  # the transfer itself: broadcast the same 56 N bytes again on the warm communicator
        t0 = time.perf_counter()
        broadcast_gaussians(arrays if rank == 0 else None, spec.n, dev, world, rank)
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
    imgs = [torch.empty((H, Wd, 3), dtype=torch.float32, device=dev) for _ in range(F)]
    img = imgs[0]
    # host side of the e2e legs: TWO pinned images per context (the copy of frame i overlaps the render of frame i+1;
    # round 1 cycled through six, 150 MB per rank, for no benefit)
    hosts = [torch.empty((H, Wd, 3), dtype=torch.float32).pin_memory() for _ in range(2 * F)]
    hosts_u8 = [torch.empty((H, Wd, 3), dtype=torch.uint8).pin_memory() for _ in range(2 * F)]
    dev_u8 = torch.zeros((H, Wd, 3), dtype=torch.uint8, device=dev)
    flushes = [torch.empty(256 << 20, dtype=torch.uint8, device=dev) for _ in range(F)]
    flush = flushes[0]

    def host_slot(pool, s):  # frame s runs in context s % F and uses that context's two host images alternately
        return pool[2 * (s % F) + ((s // F) & 1)]

    # warm-up (>= 3 frames) + one untimed sweep over the timed views: scratch reaches its final size in every context
    for s in range(W):
        for f in range(F):
            rasts[f].render(cams[shard.view_of_step(K + s)], prm, out=imgs[f])
    for s, v in enumerate(views):
        for f in range(F):
            rasts[f].render(cams[v], prm, out=imgs[f])
    for f in range(F):
        rasts[f].render(cams[views[0]], prm_a, out=hosts[2 * f])
        rasts[f].render(cams[views[0]], prm_a, out=hosts_u8[2 * f], layout="u8")
        rasts[f].join_host_copies()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def pipelined_loop(pool, params, join, layout="hwc", do_flush=False):
        """K frames, F in flight: frame s runs in context s % F on stream s % F (each context holds its OWN copy of the
        Gaussian set, so with F >= 3 the inputs cycled through are larger than L2: 3 x 56 MB > 126 MB); one event pair
        on the main stream brackets the whole loop (the side streams fork from the start event and are joined before
        the end event).  pool None: device images.  Returns total ms, kernel launches, frame infos."""
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
                out = imgs[f] if pool is None else host_slot(pool, s)
                rasts[f].render(cams[v], params, out=out, layout=layout)
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

    def d2h_only_loop(pool):
        """The egress of the e2e leg on its own: K copies of a resident device image into the same pinned host images,
        on the same F streams, nothing rendered.  If this saturates where the e2e figure does, the limiter is the
        host side of the copies, not the renderer."""
        main = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        done = [torch.cuda.Event() for _ in range(F)]
        src = imgs[0] if pool is hosts else dev_u8
        barrier()
        e0.record(main)
        for st_ in streams:
            st_.wait_event(e0)
        for s in range(K):
            f = s % F
            with torch.cuda.stream(streams[f]):
                host_slot(pool, s).copy_(src, non_blocking=True)
        for f in range(F):
            with torch.cuda.stream(streams[f]):
                done[f].record(streams[f])
            main.wait_event(done[f])
        e1.record(main)
        barrier()
        return e0.elapsed_time(e1)

    def max_over_ranks(xs):
        t = torch.tensor(xs, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    explicit_flush = F < 3  # fewer than 3 scene copies do not exceed L2: fall back to flushing inside the region
    pipelined_loop(None, prm, join=False)                                # untimed pass: clocks in steady state
    ms_dev_rep, launches, infos = [], 0, []
    for _ in range(R):
        ms, launches, infos = pipelined_loop(None, prm, join=False, do_flush=explicit_flush)
        ms_dev_rep.append(ms)
    Re = max(1, min(R, 5))
    ms_e2e_rep = [pipelined_loop(hosts, prm_a, join=True, do_flush=explicit_flush)[0] for _ in range(Re)]
    ms_u8_rep = [pipelined_loop(hosts_u8, prm_a, join=True, layout="u8", do_flush=explicit_flush)[0] for _ in range(Re)]
    ms_d2h_rep = [d2h_only_loop(hosts) for _ in range(Re)]
    ms_d2h_u8_rep = [d2h_only_loop(hosts_u8) for _ in range(Re)]
    pipelined_loop(None, prm, join=False, do_flush=True)                 # untimed pass of the flush variant
    ms_flush = pipelined_loop(None, prm, join=False, do_flush=True)[0]   # same loop, 256 MiB written per frame
    clocks = sampler.stop() if rank == 0 else None

    # single-frame latency (serial frames on ONE context, flush outside the events), next to the throughput
    lat_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for s, v in enumerate(views):
        flush.zero_()
        lat_ev[s][0].record()
        rast.render(cams[v], prm, out=img)
        lat_ev[s][1].record()
    torch.cuda.synchronize()
    lat = [a.elapsed_time(b) for a, b in lat_ev]
    # ... and the throughput of ONE context on one stream, no flush: the host queues frame i+1 while frame i runs
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for v in views:
        rast.render(cams[v], prm, out=img)
    e1.record()
    torch.cuda.synchronize()
    ms_single_ctx = e0.elapsed_time(e1)

    # frames rendered by this rank must equal the frames rank 0 renders alone for the same view ids
    frames_identical = None
    if world > 1:
        check_views = views[: min(4, len(views))]
        mine = []
        for v in check_views:
            rast.render(cams[v], prm, out=img)
            torch.cuda.synchronize()
            mine.append((v, hashlib.sha256(img.cpu().numpy().tobytes()).hexdigest()))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            frames_identical = True
            for per_rank in gathered:
                for v, digest in per_rank:
                    rast.render(cams[v], prm, out=img)
                    torch.cuda.synchronize()
                    if hashlib.sha256(img.cpu().numpy().tobytes()).hexdigest() != digest:
                        frames_identical = False

    # optional egress: every rank's first frame gathered on rank 0 as device tensors over
    # NCCL (sharding.gather_frames), compared there with rank 0's own render of the same view; outside the timed region
    gather_check = None
    if world > 1:
        one_each = ViewShard(world, rank, world)  # "view" r of this shard = rank r's first timed view
        rast.render(cams[views[0]], prm, out=img)
        torch.cuda.synchronize()
        first_views = [None] * world
        dist.all_gather_object(first_views, views[0])
        t0 = time.perf_counter()
        got = gather_frames([img], one_each)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            same = True
            for r_, v in enumerate(first_views):
                rast.render(cams[v], prm, out=img)
                torch.cuda.synchronize()
                same = same and bool(torch.equal(img, got[r_]))
            gather_check = {"frames_equal_rank0_render": same, "seconds": round(dt, 4),
                            "mb": round(world * img.numel() * 4 / 1e6, 1)}

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

    ms_dev_rep = max_over_ranks(ms_dev_rep)
    ms_e2e_rep = max_over_ranks(ms_e2e_rep)
    ms_u8_rep = max_over_ranks(ms_u8_rep)
    ms_d2h_rep = max_over_ranks(ms_d2h_rep)
    ms_d2h_u8_rep = max_over_ranks(ms_d2h_u8_rep)
    ms_flush_max, ms_single_max = max_over_ranks([ms_flush, ms_single_ctx])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fps = lambda ms: K * world / (ms * 1e-3)  # noqa: E731
    ms_dev = pct(ms_dev_rep, 0.5)
    ms_e2e = pct(ms_e2e_rep, 0.5)
    ms_u8 = pct(ms_u8_rep, 0.5)
    ms_d2h = pct(ms_d2h_rep, 0.5)
    ms_d2h_u8 = pct(ms_d2h_u8_rep, 0.5)
    img_bytes = H * Wd * 3 * 4

    hbm_peak, sm_max_mhz, peak_src = load_peaks()
    split = infos[0].depth_passes > 0
    # stage accounting averaged over this rank's timed views
    sb = {k: 0.0 for k in _lib.STAGE_NAMES}
    for info in infos:
        for k, v in stage_bytes(info, spec.n, Wd, H).items():
            sb[k] += v / len(infos)
    calib = {}
    try:  # per-kernel instruction counts and DRAM bytes of ONE frame (orbit view 0), from an ncu capture of this build:
        # written by tools/calibrate_roofline.py, never typed in by hand
        cj = json.load(open(CALIBRATION))
        key = f"{args.config}/full_cover={args.full_cover}/sort={'split' if split else 'full'}"
        calib = cj.get("frames", {}).get(key, {})
    except Exception:
        calib = {}
    cal_stage = calib.get("stages", {})
    stages = []
    for k in _lib.STAGE_NAMES:
        if stage_ms[k] <= 0:
            continue
        gbs = sb[k] / (stage_ms[k] * 1e-3) / 1e9
        st = {"stage": k, "ms": round(stage_ms[k], 4), "alg_mb": round(sb[k] / 1e6, 2),
              "achieved_gbs": round(gbs, 1), "frac_hbm": round(gbs / hbm_peak, 4)}
        if k in cal_stage:
            st["dram_mb_ncu_view0"] = round(cal_stage[k]["dram_bytes"] / 1e6, 2)
        stages.append(st)
    dom = max(stages, key=lambda s: s["ms"])
    kernel_of = {"project": "project_kernel", "depth_sort": "onesweep_kernel<u32,pairs> x4", "scan": "scan_kernel",
                 "emit": "emit_kernel", "sort": "onesweep_kernel<keys-only, entry-out>", "ranges": "tile_stats_kernel",
                 "composite": "composite_fast_kernel", "expand": "expand_kernel"}
    # Compositing is bound by issue slots (fp32 + MUFU.EX2 from shared memory), every other stage by HBM
    # (SURVEY.md section 8d).  The headline roofline is the dominant kernel's, in ITS unit; the per-stage HBM figures
    # of the streaming kernels are in `stages`.
    roofline = None
    comp_ms0 = (stage_ms_first or {}).get("composite", 0.0)  # views[0] is orbit view 0 on rank 0
    sm_mhz = (clocks or {}).get("sm_mhz") or sm_max_mhz
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    if dom["stage"] == "composite" and "composite" in cal_stage and comp_ms0 > 0 and views[0] == 0:
        winst = float(cal_stage["composite"]["warp_inst"])
        peak = n_sm * 4 * sm_mhz * 1e6
        ach = winst / (comp_ms0 * 1e-3)
        roofline = {"bound": "issue", "kernel": kernel_of["composite"], "achieved": round(ach / 1e9, 2),
                    "peak": round(peak / 1e9, 2), "unit": "Gwarp-inst/s", "frac": round(ach / peak, 4),
                    "traffic": cal_stage["composite"]["dram_bytes"],
                    "launch_ms": round(comp_ms0, 4),
                    "warp_inst_per_launch": int(winst),
                    "warp_inst_per_pixel_step": None,  # filled in below when the cpu_baseline leg counted the steps
                    "peak_source": f"{n_sm} SMs x 4 schedulers x SM clock sampled during the run ({sm_mhz:.0f} MHz)",
                    "calibration": {"file": os.path.relpath(CALIBRATION, ROOT), "view": 0,
                                    "note": "warp instructions and DRAM bytes of this kernel for orbit view 0, from an ncu "
                                            "capture of this build (tools/calibrate_roofline.py); the launch time is this "
                                            "run's own CUDA-event time of the same view"}}
    if roofline is None:
        roofline = {"bound": "hbm", "kernel": kernel_of[dom["stage"]], "achieved": dom["achieved_gbs"], "peak": hbm_peak,
                    "unit": "GB/s", "frac": dom["frac_hbm"],
                    "traffic": (cal_stage.get(dom["stage"], {}) or {}).get("dram_bytes"),
                    "launch_ms": dom["ms"], "peak_source": peak_src}
        if dom["stage"] == "composite":
            # 52 B per (tile, Gaussian) instance is an UPPER bound of what the kernel reads (early termination, records
            # served from L2): it is not a fraction of anything
            roofline.update({"bound": "issue", "achieved": None, "peak": None, "unit": "Gwarp-inst/s", "frac": None,
                             "hbm_upper_bound_gbs": dom["achieved_gbs"],
                             "note": "compositing is issue-slot bound; no ncu calibration for this config / build "
                                     "(tools/calibrate_roofline.py --config ...), so no fraction is reported"})
    roofline["stages"] = stages
    roofline["hbm_peak_gbs"] = hbm_peak
    roofline["hbm_peak_source"] = peak_src

    mean_k = sum(int(i.k_instances) for i in infos) / len(infos)
    mean_ks = sum(int(i.k_sorted) for i in infos) / len(infos)
    mean_m = sum(int(i.m_in_view) for i in infos) / len(infos)
    line = {
        "metric": METRIC, "value": fps(ms_dev), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "strong" if args.orbit else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(spec), "full_cover": args.full_cover, "tile_size": 16, "semantics": "ref_cpu",
                   "sort_mode": "split" if split else "full", "views_per_rank": K, "parallelism": f"view-sharded x{world}",
                   "frames_in_flight": F, "repeats": R,
                   "submission": (f"{sum(int(i.graph_launch) for i in infos)} of {len(infos)} timed frames went out as ONE "
                                  "CUDA graph launch each (captured from the library's own launch sequence, arguments "
                                  "updated in place); gpu_launches counts the kernels inside them"),
                   "ms_per_step_p10_p50_p90": [round(pct(ms_dev_rep, q) / K, 5) for q in (0.1, 0.5, 0.9)],
                   "l2": ("256 MiB written before every frame, inside the timed region" if explicit_flush else
                          f"inputs larger than L2: {F} contexts, each with its own {56 * spec.n / 1e6:.0f} MB copy of the "
                          f"Gaussian set, used round-robin ({F * 56 * spec.n / 1e6:.0f} MB > 126 MB L2); each frame also "
                          "streams ~0.1 GB of intermediates; no explicit flush"),
                   "value_with_flush_inside_timed_region": round(fps(ms_flush_max), 1),
                   "value_one_context_one_stream": round(fps(ms_single_max), 1),
                   "frame_latency_ms_serial": round(sum(lat) / len(lat), 4),
                   "frame_latency_ms_p10_p50_p90": [round(pct(lat, q), 4) for q in (0.1, 0.5, 0.9)],
                   "mean_in_view": round(mean_m), "mean_tile_instances": round(mean_k),
                   "mean_super_tile_instances": round(mean_ks),
                   "cull_alpha": prm.cull_alpha,
                   "gaussian_broadcast_s": None if bcast_s is None else round(bcast_s, 4),
                   "gaussian_broadcast_first_call_s": round(bcast_first_s, 4),
                   "gaussian_broadcast_note": "first call includes NCCL communicator bring-up; the second one is the "
                                              f"{56 * spec.n / 1e6:.0f} MB transfer on the warm communicator",
                   "frames_identical": frames_identical,
                   "gather_frames_nccl": gather_check,
                   "host": dict(host_topology(), cpus_bound_per_rank=numa)},
        "e2e": {"value": fps(ms_e2e), "unit": UNIT, "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": C.sizeof(_lib.GsbCamera) + C.sizeof(_lib.GsbParams),
                "d2h_bytes_per_step": img_bytes + 8,
                "repeats": Re,
                "note": "gsb_render with host structs in and a pinned fp32 host image out (async egress on the copy "
                        "stream, joined before the end event); same pipelined loop as `value`; Gaussians stay resident "
                        "like model weights",
                # the egress alone: the same K copies into the same pinned images on the same streams, nothing rendered
                "d2h_only": {"frames_per_s_equivalent": round(fps(ms_d2h), 1),
                             "gb_per_s_all_ranks": round(K * world * img_bytes / (ms_d2h * 1e-3) / 1e9, 2),
                             "gb_per_s_per_gpu": round(K * img_bytes / (ms_d2h * 1e-3) / 1e9, 2),
                             "e2e_over_d2h_only": round(fps(ms_e2e) / fps(ms_d2h), 3),
                             "note": "ceiling of any fp32-image-per-frame figure on this box; e2e at or near it means the "
                                     "copies into host memory, not the renderer, set the number"},
                # a different output type, stated as such: 8-bit (H,W,3) through gsb_render_u8, a quarter of the bytes
                "u8": {"value": round(fps(ms_u8), 1), "unit": UNIT, "d2h_bytes_per_step": H * Wd * 3 + 8,
                       "d2h_only_frames_per_s_equivalent": round(fps(ms_d2h_u8), 1),
                       "note": "same loop through gsb_render_u8 (clamp(v,0,1)*255 on the device); NOT the headline e2e: "
                               "the reference op returns fp32"}},
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
        # the oracle port, timed on this box's host cores on whole frames of the same workload
        try:
            sc0, orc, ocams, oprm, oarrays = make_oracle_inputs(args.config, args.full_cover, 3)
            dts = []
            fr = None
            t_budget = time.perf_counter()
            for i in range(3):
                dt, fr_i = oracle_frame(orc, ocams[i], oprm, oarrays)
                dts.append(dt)
                fr = fr or fr_i
                if time.perf_counter() - t_budget > 30.0:
                    break
            dt = pct(dts, 0.5)
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                                    "sample": f"{len(dts)} whole frames (orbit views 0-{len(dts) - 1}) of the workload through "
                                              "oracle/gs_oracle.c, median",
                                    "steps_executed_view0": int(fr.steps)}
            if roofline.get("bound") == "issue":
                roofline["pixel_steps_view0_this_run"] = int(fr.steps)
                roofline["warp_inst_per_pixel_step"] = round(roofline["warp_inst_per_launch"] / int(fr.steps), 4)
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
