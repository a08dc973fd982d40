"""Minimal COLMAP model readers (cameras / images; text and binary).

The reference vendors COLMAP's read_write_model.py (splat/read_colmap.py) and picks the file by
existence (splat/utils.py:269-290: *.bin first, then *.txt, else ValueError).  Only the fields the
render path consumes are parsed here: camera id/model/size/params and image id/qvec/tvec/camera/name.
2-D observations are skipped (they feed only the notebooks' scatter plots).
"""

from __future__ import annotations

import collections
import os
import struct
from typing import Dict

import numpy as np

Camera = collections.namedtuple("Camera", ["id", "model", "width", "height", "params"])
Image = collections.namedtuple("Image", ["id", "qvec", "tvec", "camera_id", "name", "xys", "point3D_ids"])

# model id -> (name, number of params), COLMAP's fixed table
_CAMERA_MODELS = {
    0: ("SIMPLE_PINHOLE", 3), 1: ("PINHOLE", 4), 2: ("SIMPLE_RADIAL", 4), 3: ("RADIAL", 5), 4: ("OPENCV", 8),
    5: ("OPENCV_FISHEYE", 8), 6: ("FULL_OPENCV", 12), 7: ("FOV", 5), 8: ("SIMPLE_RADIAL_FISHEYE", 4),
    9: ("RADIAL_FISHEYE", 5), 10: ("THIN_PRISM_FISHEYE", 12),
}


def _data_lines(path):
    with open(path, "r") as f:
        for raw in f:
            yield raw.strip()


def read_cameras_text(path: str) -> Dict[int, Camera]:
    cams = {}
    for line in _data_lines(path):
        if not line or line.startswith("#"):
            continue
        tok = line.split()
        cid = int(tok[0])
        cams[cid] = Camera(cid, tok[1], int(tok[2]), int(tok[3]), np.array([float(t) for t in tok[4:]]))
    return cams


def read_images_text(path: str) -> Dict[int, Image]:
    imgs = {}
    it = _data_lines(path)
    for line in it:
        if not line or line.startswith("#"):
            continue
        tok = line.split()
        iid = int(tok[0])
        q = np.array([float(t) for t in tok[1:5]])
        t = np.array([float(t) for t in tok[5:8]])
        imgs[iid] = Image(iid, q, t, int(tok[8]), tok[9], np.zeros((0, 2)), np.zeros((0,), np.int64))
        next(it, None)  # the 2-D observation line that follows every image line
    return imgs


def read_cameras_binary(path: str) -> Dict[int, Camera]:
    cams = {}
    with open(path, "rb") as f:
        (count,) = struct.unpack("<Q", f.read(8))
        for _ in range(count):
            cid, model_id, w, h = struct.unpack("<iiQQ", f.read(24))
            name, npar = _CAMERA_MODELS[model_id]
            params = struct.unpack("<" + "d" * npar, f.read(8 * npar))
            cams[cid] = Camera(cid, name, w, h, np.array(params))
    return cams


def read_images_binary(path: str) -> Dict[int, Image]:
    imgs = {}
    with open(path, "rb") as f:
        (count,) = struct.unpack("<Q", f.read(8))
        for _ in range(count):
            vals = struct.unpack("<idddddddi", f.read(64))
            iid, q, t, cam = vals[0], np.array(vals[1:5]), np.array(vals[5:8]), vals[8]
            name = b""
            while True:
                ch = f.read(1)
                if ch == b"\x00" or ch == b"":
                    break
                name += ch
            (n2d,) = struct.unpack("<Q", f.read(8))
            f.seek(24 * n2d, os.SEEK_CUR)  # (x, y, point3D_id) triples: not used on the render path
            imgs[iid] = Image(iid, q, t, cam, name.decode("utf-8"), np.zeros((0, 2)), np.zeros((0,), np.int64))
    return imgs


def read_camera_file(colmap_path: str) -> Dict[int, Camera]:
    b, t = os.path.join(colmap_path, "cameras.bin"), os.path.join(colmap_path, "cameras.txt")
    if os.path.exists(b):
        return read_cameras_binary(b)
    if os.path.exists(t):
        return read_cameras_text(t)
    raise ValueError(f"no cameras.bin / cameras.txt under {colmap_path}")


def read_image_file(colmap_path: str) -> Dict[int, Image]:
    b, t = os.path.join(colmap_path, "images.bin"), os.path.join(colmap_path, "images.txt")
    if os.path.exists(b):
        return read_images_binary(b)
    if os.path.exists(t):
        return read_images_text(t)
    raise ValueError(f"no images.bin / images.txt under {colmap_path}")
