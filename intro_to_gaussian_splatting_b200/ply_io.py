"""PLY I/O without `plyfile`: the reference's point-cloud layout, and the FULL Gaussian record.

The reference writes `<model_path>/point_cloud.ply` from the Gaussians constructor (splat/gaussians.py:17-18)
through plyfile and a Python tuple loop (splat/utils.py:102-125), and can read it back with fetchPly
(splat/utils.py:93-99) -- positions and colours only; scale, rotation and opacity are hard-coded in the constructor
(splat/gaussians.py:23-33) and never persisted, so a trained scene cannot be loaded at all.  Here:

* storePly / fetchPly: the reference's vertex layout -- x y z nx ny nz (float32) red green blue (uint8), binary
  little endian -- as one structured numpy array (a million points take milliseconds);
* save_gaussians / load_gaussians: all five attribute tensors of `Gaussians`, either in this package's own layout
  (x y z sx sy sz qw qx qy qz r g b opacity_logit: exactly what the container holds, float32) or in the property
  names every 3D-Gaussian-splatting trainer writes (x y z nx ny nz f_dc_0..2 [f_rest_*] opacity scale_0..2 rot_0..3)
  with the documented conversions: scales = exp(scale_*), quaternion = rot_* (w x y z), opacity logit as is,
  colour = clamp(0.5 + 0.28209479 * f_dc, 0, 1) (degree-0 spherical harmonic; higher degrees are ignored: the
  reference has plain RGB).

Host-side convenience next to the render path, not part of it."""

from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from .schema import BasicPointCloud

_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                   ("red", "u1"), ("green", "u1"), ("blue", "u1")])
_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1",
              "char": "i1", "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
              "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}
_SH_C0 = 0.28209479177387814
NATIVE_FIELDS = ("x", "y", "z", "sx", "sy", "sz", "qw", "qx", "qy", "qz", "r", "g", "b", "opacity_logit")


def _write(path: str, v: np.ndarray, comments: Tuple[str, ...] = ()) -> None:
    names = {"<f4": "float", "u1": "uchar", "|u1": "uchar"}
    header = "ply\nformat binary_little_endian 1.0\n" + "".join(f"comment {c}\n" for c in comments)
    header += f"element vertex {v.shape[0]}\n"
    for k in v.dtype.names:
        header += f"property {names[v.dtype[k].str]} {k}\n"
    header += "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(v.tobytes())


def _read_vertices(path: str) -> Tuple[np.ndarray, List[str]]:
    """The vertex element of a binary-little-endian PLY as a structured array (+ the header's comment lines)."""
    with open(path, "rb") as f:
        n = None
        props: List[Tuple[str, str]] = []
        fmt = None
        comments: List[str] = []
        in_vertex = False
        seen_other_element_first = False
        while True:
            line = f.readline()
            if not line:
                raise ValueError("PLY header not terminated")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "comment":
                comments.append(" ".join(tok[1:]))
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n = int(tok[2])
                elif n is None:
                    seen_other_element_first = True
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError("list properties in the vertex element are not supported")
                if tok[1] not in _PLY_TYPES:
                    raise ValueError(f"unknown PLY property type {tok[1]!r}")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt != "binary_little_endian" or n is None:
            raise ValueError("only binary_little_endian PLY files with a vertex element are supported")
        if seen_other_element_first:
            raise ValueError("the vertex element must come first")
        dt = np.dtype(props)
        data = f.read(n * dt.itemsize)
        if len(data) < n * dt.itemsize:
            raise ValueError("PLY file is truncated")
        return np.frombuffer(data, dtype=dt, count=n), comments


def storePly(path: str, xyz, rgb) -> None:
    """xyz (N,3) float, rgb (N,3) in 0..255 -- the arguments of the reference's storePly."""
    xyz = np.asarray(xyz, dtype=np.float32)
    rgb = np.asarray(rgb)
    v = np.zeros(xyz.shape[0], dtype=_DTYPE)
    v["x"], v["y"], v["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    v["red"], v["green"], v["blue"] = (np.clip(rgb[:, i], 0, 255).astype(np.uint8) for i in range(3))
    _write(path, v)


def fetchPly(path: str) -> BasicPointCloud:
    """Read a PLY written by storePly (or any binary-little-endian PLY with x y z [nx ny nz] red green blue)."""
    v, _ = _read_vertices(path)
    names = v.dtype.names
    for k in ("x", "y", "z", "red", "green", "blue"):
        if k not in names:
            raise ValueError(f"unexpected vertex layout {names}")
    positions = np.stack([v["x"], v["y"], v["z"]], axis=1)
    colors = np.stack([v["red"], v["green"], v["blue"]], axis=1) / 255.0
    normals = (np.stack([v["nx"], v["ny"], v["nz"]], axis=1) if "nx" in names else np.zeros_like(positions))
    return BasicPointCloud(points=positions, colors=colors, normals=normals)


def _np(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


def save_gaussians(path: str, points, scales, quaternions, colors, opacity, layout: str = "native") -> None:
    """All five attribute tensors of `Gaussians` (reference layouts: (N,3) (N,3) linear (N,4) wxyz (N,3) = rgb/256
    (N,1) logit) -> one PLY.  layout "native": the values as they are; "3dgs": the property names and encodings of
    the 3D-Gaussian-splatting trainers (log scales, f_dc colours)."""
    p, s, q, c, o = (_np(t) for t in (points, scales, quaternions, colors, opacity))
    n = p.shape[0]
    o = o.reshape(n)
    if s.shape != (n, 3) or q.shape != (n, 4) or c.shape != (n, 3) or o.shape != (n,):
        raise ValueError("save_gaussians: attribute shapes do not match (N,3) (N,3) (N,4) (N,3) (N,1)")
    if layout == "native":
        v = np.zeros(n, dtype=np.dtype([(k, "<f4") for k in NATIVE_FIELDS]))
        for k, col in zip(NATIVE_FIELDS, (p[:, 0], p[:, 1], p[:, 2], s[:, 0], s[:, 1], s[:, 2], q[:, 0], q[:, 1], q[:, 2],
                                          q[:, 3], c[:, 0], c[:, 1], c[:, 2], o)):
            v[k] = col
        _write(path, v, ("intro_to_gaussian_splatting_b200 native Gaussian record: linear scales, wxyz quaternions, "
                         "rgb/256 colours, opacity logit",))
    elif layout == "3dgs":
        if (s <= 0).any():
            raise ValueError("save_gaussians(layout='3dgs'): scales must be positive (stored as logarithms)")
        names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2",
                 "rot_0", "rot_1", "rot_2", "rot_3"]
        v = np.zeros(n, dtype=np.dtype([(k, "<f4") for k in names]))
        v["x"], v["y"], v["z"] = p[:, 0], p[:, 1], p[:, 2]
        for i in range(3):
            v[f"f_dc_{i}"] = (c[:, i] - 0.5) / _SH_C0
            v[f"scale_{i}"] = np.log(s[:, i])
        for i in range(4):
            v[f"rot_{i}"] = q[:, i]
        v["opacity"] = o
        _write(path, v)
    else:
        raise ValueError(f"unknown layout {layout!r}")


def load_gaussians(path: str) -> Dict[str, np.ndarray]:
    """-> {"points" (N,3), "scales" (N,3) linear, "quaternions" (N,4) wxyz, "colors" (N,3) in [0,1], "opacity" (N,1)
    logit}, float32, from either layout save_gaussians writes (so also from any trained 3DGS point_cloud.ply)."""
    v, _ = _read_vertices(path)
    names = set(v.dtype.names)
    col = lambda *ks: np.stack([v[k].astype(np.float32) for k in ks], axis=1)  # noqa: E731
    if set(NATIVE_FIELDS) <= names:
        return {"points": col("x", "y", "z"), "scales": col("sx", "sy", "sz"), "quaternions": col("qw", "qx", "qy", "qz"),
                "colors": col("r", "g", "b"), "opacity": col("opacity_logit")}
    need = {"x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1",
            "rot_2", "rot_3"}
    if need <= names:
        colors = np.clip(0.5 + _SH_C0 * col("f_dc_0", "f_dc_1", "f_dc_2"), 0.0, 1.0).astype(np.float32)
        return {"points": col("x", "y", "z"), "scales": np.exp(col("scale_0", "scale_1", "scale_2")),
                "quaternions": col("rot_0", "rot_1", "rot_2", "rot_3"), "colors": colors, "opacity": col("opacity")}
    raise ValueError(f"{path}: neither the native nor the 3DGS Gaussian layout (properties: {sorted(names)})")
