"""Point-cloud PLY I/O without `plyfile`.

The reference writes `<model_path>/point_cloud.ply` from the Gaussians constructor (splat/gaussians.py:17-18)
through plyfile and a Python tuple loop (splat/utils.py:102-125), and can read it back with fetchPly
(splat/utils.py:93-99).  This module writes/reads the same vertex layout -- x y z nx ny nz (float32) red green blue
(uint8), binary little endian -- with one structured numpy array, so a million points take milliseconds.  It is
host-side convenience next to the render path, not part of it."""

from __future__ import annotations

import numpy as np

from .schema import BasicPointCloud

_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                   ("red", "u1"), ("green", "u1"), ("blue", "u1")])


def storePly(path: str, xyz, rgb) -> None:
    """xyz (N,3) float, rgb (N,3) in 0..255 -- the arguments of the reference's storePly."""
    xyz = np.asarray(xyz, dtype=np.float32)
    rgb = np.asarray(rgb)
    n = xyz.shape[0]
    v = np.zeros(n, dtype=_DTYPE)
    v["x"], v["y"], v["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    v["red"], v["green"], v["blue"] = (np.clip(rgb[:, i], 0, 255).astype(np.uint8) for i in range(3))
    header = ("ply\nformat binary_little_endian 1.0\n" + f"element vertex {n}\n"
              + "".join(f"property float {k}\n" for k in ("x", "y", "z", "nx", "ny", "nz"))
              + "".join(f"property uchar {k}\n" for k in ("red", "green", "blue")) + "end_header\n")
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(v.tobytes())


def fetchPly(path: str) -> BasicPointCloud:
    """Read a PLY written by storePly (or any binary-little-endian PLY with exactly that vertex layout)."""
    with open(path, "rb") as f:
        n = None
        props = []
        fmt = None
        while True:
            line = f.readline()
            if not line:
                raise ValueError("PLY header not terminated")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                if tok[1] == "vertex":
                    n = int(tok[2])
            elif tok[0] == "property" and n is not None:
                props.append((tok[2], tok[1]))
            elif tok[0] == "end_header":
                break
        if fmt != "binary_little_endian" or n is None:
            raise ValueError("only binary_little_endian PLY point clouds are supported")
        want = [(k, "float") for k in ("x", "y", "z", "nx", "ny", "nz")] + [(k, "uchar") for k in ("red", "green", "blue")]
        if props[:9] != want:
            raise ValueError(f"unexpected vertex layout {props}")
        v = np.frombuffer(f.read(n * _DTYPE.itemsize), dtype=_DTYPE, count=n)
    positions = np.stack([v["x"], v["y"], v["z"]], axis=1)
    colors = np.stack([v["red"], v["green"], v["blue"]], axis=1) / 255.0
    normals = np.stack([v["nx"], v["ny"], v["nz"]], axis=1)
    return BasicPointCloud(points=positions, colors=colors, normals=normals)
