"""Multi-GPU parity on hardware: the gallery-sharded compare (NCCL merge) equals the single-GPU pass.

Runs ``tests/multi_gpu_check.py`` under torchrun on 2 GPUs when the box has them (``gpurun --gpus 2``); skipped on a
one-GPU box, where ``tests/test_sharding_gloo.py`` (CPU, world_size 2) covers the collective plumbing."""

import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = Path(__file__).resolve().parents[1]


def test_two_gpu_sharded_compare_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(ROOT / "tests" / "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "multi-GPU check ok" in res.stdout
