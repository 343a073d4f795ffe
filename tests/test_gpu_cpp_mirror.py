"""The C++ host-side mirror (include/threecrate_cuda.hpp) through its compiled test program
(tests/cpp/host_mirror_test.cpp): the reference's own test scenarios, from a compiled host."""
import os
import subprocess

import pytest

from threecrate_b200 import build

BIN = build.HOST_TEST_BIN


def test_cpp_mirror_builds_and_fails_loudly_without_gpu():
    build.build()
    assert os.path.exists(BIN)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=120)
    assert r.returncode == 77 and "no CUDA device" in r.stderr  # no CPU fallback


@pytest.mark.gpu
def test_cpp_mirror_reference_scenarios():
    build.build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stderr
    assert "0 failed" in r.stdout
