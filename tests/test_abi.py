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
    assert lib.gsb_version() == _lib.GSB_API_VERSION == 2


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.GsbCamera) == 16 * 4 * 2 + 4 * 4 + 2 * 4
    assert C.sizeof(_lib.GsbParams) == 15 * 4
    assert C.sizeof(_lib.GsbFrameInfo) == 3 * 8 + 6 * 4 + 2 * 8 + 2 * 4 + 8 + 2 * 4


def test_struct_layouts_match_a_c_compiler(tmp_path):
    """Compile include/gsb.h with the host C compiler and compare sizeof / offsetof of every field with the ctypes
    mirrors (this package's and the oracle's): the header is the single source of truth of the ABI."""
    import shutil
    import subprocess

    from oracle import oracle as orc

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {"GsbCamera": (_lib.GsbCamera, orc.Camera), "GsbParams": (_lib.GsbParams, orc.Params),
               "GsbFrameInfo": (_lib.GsbFrameInfo, None)}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "gsb.h"', "int main(void) {"]
    for name, (ct, _) in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _t in ct._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([cc, "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, (ct, oc) in structs.items():
        assert int(got[name]) == C.sizeof(ct), name
        for field, _t in ct._fields_:
            assert int(got[f"{name}.{field}"]) == getattr(ct, field).offset, (name, field)
        if oc is not None:
            assert C.sizeof(oc) == C.sizeof(ct)
            assert [f for f, _ in oc._fields_] == [f for f, _ in ct._fields_], name
            for field, _t in oc._fields_:
                assert getattr(oc, field).offset == getattr(ct, field).offset, (name, field)


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
