"""In-tree build of libthreecrate_cuda.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m threecrate_b200.build [--force] [--verbose]

Objects go to threecrate_b200/csrc/_obj/, the library to threecrate_b200/lib/ (git-ignored, but
shipped to the GPU box by gpurun).  cudart is linked statically; libnccl is dlopen'ed at run time.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libthreecrate_cuda.so")
SOURCES = ["tc_api.cu", "tc_index.cu", "tc_search.cu", "tc_tile.cu", "tc_icp.cu", "tc_comm.cu",
           "tc_filter.cu"]
HEADERS = ["tc_internal.cuh", "tc_search.cuh", "tc_normal.cuh", os.path.join(ROOT, "include", "threecrate_cuda.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "-ccbin", "/usr/bin/g++", "-I", os.path.join(ROOT, "include")]


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _hdr_mtime():
    return max(_mtime(h if os.path.isabs(h) else os.path.join(CSRC, h)) for h in HEADERS)


EXTRA_DEFS: list = []  # e.g. ["-DTC_QCLOCK"] (tools/qclock.py)


def _compile(src: str, force: bool, verbose: bool):
    s = os.path.join(CSRC, src)
    o = os.path.join(OBJ, src.replace(".cu", ".o"))
    if not force and _mtime(o) > max(_mtime(s), _hdr_mtime()):
        return o, ""
    cmd = [NVCC, *ARCH, *CFLAGS, *EXTRA_DEFS, "-c", s, "-o", o]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return o, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(lambda s: _compile(s, force, verbose), SOURCES))
    objs = [o for o, _ in res]
    if verbose:
        for _, log in res:
            sys.stderr.write(log)
    if force or _mtime(LIB) < max(_mtime(o) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB, *objs, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    build_host_mirror_test(force)
    return LIB


HOST_TEST_SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
HOST_TEST_BIN = os.path.join(LIBDIR, "host_mirror_test")


def build_host_mirror_test(force: bool = False) -> str:
    """g++ build of the C++ host-side mirror's test program (include/threecrate_cuda.hpp over the
    C ABI); it links the shared library and finds it next to itself at run time."""
    hpp = os.path.join(ROOT, "include", "threecrate_cuda.hpp")
    if not os.path.exists(HOST_TEST_SRC):
        return ""
    newest = max(_mtime(HOST_TEST_SRC), _mtime(hpp), _mtime(LIB))
    if force or _mtime(HOST_TEST_BIN) < newest:
        cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-I",
               os.path.join(ROOT, "include"), HOST_TEST_SRC, "-L", LIBDIR, "-lthreecrate_cuda",
               "-Wl,-rpath,$ORIGIN", "-o", HOST_TEST_BIN]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"host mirror test failed to build:\n{r.stdout}\n{r.stderr}")
    return HOST_TEST_BIN


if __name__ == "__main__":
    if "--qclock" in sys.argv:  # per-query cycle counters compiled into the search kernels
        EXTRA_DEFS.append("-DTC_QCLOCK")
    print(build(force="--force" in sys.argv or "--qclock" in sys.argv, verbose="--verbose" in sys.argv))
