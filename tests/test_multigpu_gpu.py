"""GPU test of the strip-sharded frame across the GPUs of one box (needs >= 2 GPUs; skipped on a single-GPU box)."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
@pytest.mark.parametrize("size", [(640, 384), (1920, 1080)])
def test_strip_sharded_frame_equals_single_gpu_frame(size, transport):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29611",
           str(ROOT / "tests" / "mgpu_strip_check.py"), str(size[0]), str(size[1]), transport]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


def test_strip_sharded_frame_with_cost_aware_bounds():
    """Non-uniform strips (sharding.rebalance_bounds) reproduce the single-GPU frame bit for bit as well."""
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29612",
           str(ROOT / "tests" / "mgpu_strip_check.py"), "1920", "1080", "p2p", "uneven"]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
