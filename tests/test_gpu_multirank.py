"""GPU (needs >= 2 devices): data-parallel correctness of the ENGINE's train step on the device - see tests/dist_worker.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_equal_one_rank_on_the_global_batch(engine):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("MULTIRANK ")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-4000:]
    out = json.loads(lines[-1][len("MULTIRANK "):])
    assert out["ok"], out
