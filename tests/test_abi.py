"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/threecrate_cuda.h declares; host-side validation mirrors the reference's error
behaviour; and without a CUDA device the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import threecrate_b200 as tc
from threecrate_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_builds_and_exports_every_declared_symbol():
    build.build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "threecrate_cuda.h")).read()
    declared = set(re.findall(r"\b(tc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.tc_version()


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.IcpResultC) == 48  # 7 f32 + f32 + u32 + i32 + (pad) + u64
    assert _lib.IcpResultC.n_correspondences.offset == 40
    assert C.sizeof(_lib.IndexInfoC) == 72
    assert C.sizeof(_lib.IcpScaleLevelC) == 12  # f32 + u32 + f32
    assert _lib.IcpScaleLevelC.max_correspondence_distance.offset == 8


def test_cpp_header_is_warning_free_and_self_contained(tmp_path):
    """include/threecrate_cuda.hpp compiles on its own with -Wall -Wextra -Werror -pedantic
    (syntax only: no CUDA toolkit or library needed by a host that just binds the ABI)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.cpp"
    src.write_text('#include "threecrate_cuda.hpp"\nint main() { return 0; }\n')
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-pedantic",
                        "-fsyntax-only", "-I", os.path.join(root, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_sass_is_sm100a_only():
    out = os.popen(f"cuobjdump -lelf {_lib.LIB_PATH} 2>/dev/null").read()
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(tc.GpuError):
        tc.Context(0)
    pts = np.random.default_rng(0).normal(size=(10, 3)).astype(np.float32)
    with pytest.raises(tc.GpuError):
        tc.estimate_normals(pts, 5)


def test_empty_cloud_is_ok_before_anything_else():
    # normals.rs:261-263 — empty -> Ok(empty) even with an invalid k, and without touching a GPU
    assert tc.estimate_normals(np.zeros((0, 3), np.float32), 2).shape == (0, 6)
    idx, dist, cnt = tc.k_nearest_neighbors(np.zeros((0, 3), np.float32), 4)
    assert idx.shape[0] == 0


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "threecrate_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "oracle/" not in src or f == "__init__.py", f
