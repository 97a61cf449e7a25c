"""The row-sharded step and item-sharded retrieval on real GPUs: one NCCL rank per visible device (at most 8), every
exchange route, against the reference goldens and - at tower shapes, TF32 and fp32 - against the oracle's un-sharded step
(scripts/check_sharded_gpu.py holds the per-rank body).  On a one-GPU box this runs with world size 1 (slot padding and the
static route are still live); the N = 2 / 4 / 8 logs of the same command are under profiles/."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_sharded_step_matches_reference_on_all_visible_gpus():
    world = max(1, min(torch.cuda.device_count(), 8))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + os.getpid() % 400), str(ROOT / "scripts" / "check_sharded_gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if "world=" in ln]
    assert len(lines) == 8 and all(" OK " in ln for ln in lines), r.stdout[-4000:]
