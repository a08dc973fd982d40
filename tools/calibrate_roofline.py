#!/usr/bin/env python
"""Writes the per-kernel instruction counts and DRAM bytes bench.py's roofline uses, from an ncu capture of THIS build.

    python tools/calibrate_roofline.py [--config cfg3] [--out gpurun_out/r2_roofline_calibration.json]

Runs `tools/profile_frame.py --views 0,0,0` under ncu (metrics: smsp__inst_executed.sum, dram bytes read/written,
gpu__time_duration.sum; --clock-control none), keeps the kernels of the LAST frame, groups them into the stages of
gsb_stage_times.  bench.py divides the warp-instruction count of the compositing kernel by its own CUDA-event time of
orbit view 0 (and, when its cpu_baseline leg ran, by the pixel-steps that leg counted): no constant is typed in by
hand, and the file says which build it belongs to.  Copy the result to profiles/ after a GPU run.
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

METRICS = "smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"


def stage_of(name: str) -> str:
    if "project_kernel" in name:
        return "project"
    if "tile_stats_kernel" in name:
        return "ranges"
    if "scan_kernel" in name:
        return "scan"
    if "emit_kernel" in name:
        return "emit"
    if "expand_kernel" in name:
        return "expand"
    if "composite" in name:
        return "composite"
    if "onesweep_kernel" in name:
        # template arguments: <key type, items, mode, status word, look batch>; mode 0 = (key, payload) pairs
        args = name.split("onesweep_kernel<", 1)[1].split(">", 1)[0].split(",")
        mode = args[2].strip().lstrip("(int)")
        return "depth_sort" if mode == "0" else "sort"
    return "other"


def to_bytes(value: str, unit: str) -> float:
    v = float(value.replace(",", ""))
    u = unit.strip().lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(value: str, unit: str) -> float:
    v = float(value.replace(",", ""))
    u = unit.strip().lower()
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--full-cover", type=int, default=1)
    ap.add_argument("--sort-mode", default="auto")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_roofline_calibration.json"))
    a = ap.parse_args()

    fd, log = tempfile.mkstemp(suffix=".csv")
    os.close(fd)
    cmd = ["ncu", "--metrics", METRICS, "--clock-control", "none", "--csv", "--log-file", log, sys.executable,
           os.path.join(ROOT, "tools", "profile_frame.py"), "--config", a.config, "--views", "0,0,0", "--full-cover",
           str(a.full_cover), "--sort-mode", a.sort_mode]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    text = "".join(l for l in open(log) if not l.startswith("=="))
    rows = list(csv.DictReader(io.StringIO(text)))
    launches = {}
    order = []
    for r in rows:
        lid = int(r["ID"])
        if lid not in launches:
            launches[lid] = {"name": r["Kernel Name"]}
            order.append(lid)
        m, v, u = r["Metric Name"], r["Metric Value"], r["Metric Unit"]
        if m == "smsp__inst_executed.sum":
            launches[lid]["warp_inst"] = float(v.replace(",", ""))
        elif m.startswith("dram__bytes"):
            launches[lid]["dram_bytes"] = launches[lid].get("dram_bytes", 0.0) + to_bytes(v, u)
        elif m == "gpu__time_duration.sum":
            launches[lid]["us"] = to_us(v, u)
    seq = [launches[i] for i in order if stage_of(launches[i]["name"]) != "other"]
    # three identical frames, the first of which may have queued its tail twice (buffers still growing): keep what was
    # launched from the last projection on
    first = max(i for i, k in enumerate(seq) if stage_of(k["name"]) == "project")
    last = seq[first:]
    stages = {}
    for k in last:
        s = stages.setdefault(stage_of(k["name"]), {"warp_inst": 0.0, "dram_bytes": 0.0, "ncu_us": 0.0, "launches": 0})
        s["warp_inst"] += k.get("warp_inst", 0.0)
        s["dram_bytes"] += k.get("dram_bytes", 0.0)
        s["ncu_us"] += k.get("us", 0.0)
        s["launches"] += 1

    try:
        rev = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:
        rev = ""
    import hashlib
    lib = os.path.join(ROOT, "intro_to_gaussian_splatting_b200", "libgsb_b200.so")
    lib_sha = hashlib.sha256(open(lib, "rb").read()).hexdigest()[:16]
    split = "full" if a.sort_mode == "full" else "split"
    key = f"{a.config}/full_cover={a.full_cover}/sort={split}"
    out = {}
    if os.path.exists(a.out):
        try:
            out = json.load(open(a.out))
        except Exception:
            out = {}
    out.setdefault("frames", {})[key] = {
        "view": 0, "stages": stages,
        "kernels": [{"name": k["name"][:120], "warp_inst": k.get("warp_inst"), "dram_bytes": k.get("dram_bytes"), "us": k.get("us")}
                    for k in last],
    }
    out["how"] = ("ncu --metrics " + METRICS + " --clock-control none over tools/profile_frame.py --views 0,0,0 (last frame kept)")
    out["build"] = {"git": rev, "lib_sha256_16": lib_sha}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    c = stages.get("composite", {})
    print(f"{key}: composite {c.get('warp_inst', 0) / 1e6:.1f} M warp inst, {c.get('dram_bytes', 0) / 1e6:.1f} MB DRAM, "
          f"{c.get('ncu_us', 0):.1f} us under ncu")


if __name__ == "__main__":
    main()
