"""Multi-GPU paths (one process per GPU, NCCL).  Run under gpurun --gpus 2+; skipped on 1 GPU.
The same host logic is covered on CPU by tests/test_shard_logic.py (gloo, world_size 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.skipif("_ngpu() < 2")
def test_sharded_icp_and_normals_match_single_gpu():
    """2 ranks: source-sharded ICP with the 29-scalar all-reduce reproduces the 1-GPU transform to
    f32 rounding, all ranks agree bit-for-bit, and shard-wise normals tile the full result."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29531",
           os.path.join(ROOT, "tests", "multi_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "MULTI_OK" in r.stdout
