"""The C-ABI library: loads on a CPU-only box, exports every symbol include/gsb.h declares, and
refuses to run (no fallback) without a CUDA device."""

import ctypes as C
import os
import re

import pytest
import torch

from intro_to_gaussian_splatting_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gsb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_library_agree():
    declared = _declared_symbols()
    assert declared == sorted(_lib.EXPORTED_SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in gsb.h but not exported"
    assert lib.gsb_version() == 1


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.GsbCamera) == 16 * 4 * 2 + 4 * 4 + 2 * 4
    assert C.sizeof(_lib.GsbParams) == 14 * 4
    assert C.sizeof(_lib.GsbFrameInfo) == 3 * 8 + 6 * 4


def test_default_params_are_the_reference_literals():
    p = _lib.default_params()
    assert p.tile_size == 16
    assert abs(p.minimum_z - 0.2) < 1e-7 and abs(p.fov_clamp - 1.3) < 1e-7
    assert abs(p.det_min - 1e-3) < 1e-9 and abs(p.lambda_floor - 0.1) < 1e-8
    assert p.sigma_extent == 3.0 and abs(p.min_weight - 1e-6) < 1e-12 and abs(p.alpha_max - 0.99) < 1e-7
    assert p.semantics == _lib.GSB_SEM_REF_CPU and p.full_cover == 0
    with pytest.raises(TypeError):
        _lib.default_params(not_a_field=1)


def test_error_strings():
    assert _lib.error_string(0) == "ok"
    assert "no CPU fallback" in _lib.error_string(-5)
    with pytest.raises(RuntimeError):
        _lib.check(-1, "x")


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_device_fails_loudly():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.gsb_create(C.byref(h), 0) == -5 and not h.value
    from intro_to_gaussian_splatting_b200 import Rasterizer

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Rasterizer()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "intro_to_gaussian_splatting_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert "gs_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
