"""Shared test plumbing: scenes, cameras, and the bridge between the product's structs and the oracle's."""

from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np
import torch

from intro_to_gaussian_splatting_b200 import _lib
from intro_to_gaussian_splatting_b200.colmap_io import read_camera_file, read_image_file
from intro_to_gaussian_splatting_b200.image import GaussianImage
from intro_to_gaussian_splatting_b200.synth import make_scene, write_colmap_text
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def scene_and_images(name, n_views=None, n_override=None):
    """SynthScene + {image id: GaussianImage} built the way a user would (COLMAP text model)."""
    sc = make_scene(name, n_views=n_views, n_override=n_override)
    d = tempfile.mkdtemp(prefix="gsb_t_")
    write_colmap_text(sc, d)
    cams, imgs = read_camera_file(d), read_image_file(d)
    images = {i: GaussianImage(cams[im.camera_id], im) for i, im in imgs.items()}
    return sc, images, d


def to_oracle_camera(cam: _lib.GsbCamera) -> orc.Camera:
    o = orc.Camera()
    assert C.sizeof(o) == C.sizeof(cam)
    C.memmove(C.byref(o), C.byref(cam), C.sizeof(cam))
    return o


def to_oracle_params(p: _lib.GsbParams) -> orc.Params:
    o = orc.Params()
    assert C.sizeof(o) == C.sizeof(p)
    C.memmove(C.byref(o), C.byref(p), C.sizeof(p))
    return o


def scene_arrays(sc):
    """The five Gaussian attribute arrays as the Gaussians container would hold them."""
    return sc.xyz, sc.scales, sc.quats, (sc.rgb255 / 256).float(), sc.opacity_logit


def bits(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def u64(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().view(np.uint64)
