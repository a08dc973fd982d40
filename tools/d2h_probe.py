#!/usr/bin/env python
"""Device -> host copy ceiling of the box, per allocation strategy and rank count (not a benchmark of the renderer).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/d2h_probe.py

Every rank copies a 1920x1080x3 fp32 image (24.9 MB) from its GPU into host memory 200 times on two streams, all
ranks at once, for each kind of host buffer:
  pinned      cudaHostAlloc (what torch.Tensor.pin_memory() gives; what bench.py's e2e leg uses)
  registered  anonymous mmap + MADV_HUGEPAGE, touched, then cudaHostRegister
  ring1       pinned, ONE buffer reused by every copy (smallest host footprint)
Prints GB/s per rank (min / max) and the sum over ranks.  The e2e figure of bench.py cannot exceed
sum / 24.9 MB frames/s whatever the renderer does.
"""
import json
import mmap
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    H, W = 1080, 1920
    nbytes = H * W * 3 * 4
    src = torch.rand((H, W, 3), device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    reps = 200

    def run(bufs):
        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        for b in bufs:
            b.copy_(src, non_blocking=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main_s = torch.cuda.current_stream(dev)
        e0.record(main_s)
        for s in streams:
            s.wait_event(e0)
        for i in range(reps):
            with torch.cuda.stream(streams[i % 2]):
                bufs[i % len(bufs)].copy_(src, non_blocking=True)
        for s in streams:
            ev = torch.cuda.Event()
            ev.record(s)
            main_s.wait_event(ev)
        e1.record(main_s)
        barrier()
        ms = e0.elapsed_time(e1)
        gbs = reps * nbytes / (ms * 1e-3) / 1e9
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(allv, t)
        else:
            allv = [t]
        return [float(v) for v in allv]

    out = {"ranks": world}
    pinned = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(6)]
    out["pinned_ring6"] = run(pinned)
    out["pinned_ring2"] = run(pinned[:2])
    out["pinned_ring1"] = run(pinned[:1])
    try:
        size = (2 * nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20)
        mm = mmap.mmap(-1, size + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        if hasattr(mmap, "MADV_HUGEPAGE"):
            mm.madvise(mmap.MADV_HUGEPAGE)
        whole = torch.frombuffer(mm, dtype=torch.uint8)
        off = (-whole.data_ptr()) % (2 << 20)
        region = whole[off: off + size]
        region.zero_()  # touch: fault the (huge) pages in
        rc = torch.cuda.cudart().cudaHostRegister(region.data_ptr(), size, 0)
        out["registered_rc"] = int(rc) if not isinstance(rc, int) else rc
        bufs = [region[i * nbytes:(i + 1) * nbytes].view(torch.float32).view(H, W, 3) for i in range(2)]
        out["registered_is_pinned"] = bool(bufs[0].is_pinned())
        out["registered_thp_ring2"] = run(bufs)
        torch.cuda.cudart().cudaHostUnregister(region.data_ptr())
    except Exception as e:  # noqa: BLE001
        out["registered_error"] = repr(e)
    if rank == 0:
        info = {}
        for p in ("/sys/kernel/mm/transparent_hugepage/enabled", "/sys/kernel/mm/transparent_hugepage/defrag"):
            try:
                info[p] = open(p).read().strip()
            except Exception:
                pass
        try:
            info["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        except Exception:
            pass
        try:
            info["meminfo_huge"] = [l.strip() for l in open("/proc/meminfo") if "Huge" in l]
        except Exception:
            pass
        summary = {k: {"sum_gbs": round(sum(v), 1), "min": round(min(v), 1), "max": round(max(v), 1),
                       "frames_per_s_ceiling": round(sum(v) * 1e9 / nbytes)} for k, v in out.items() if isinstance(v, list)}
        print(json.dumps({"d2h_probe": summary, "host": info, "extra": {k: v for k, v in out.items() if not isinstance(v, list)}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
