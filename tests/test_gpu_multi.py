"""Multi-GPU parity (needs >= 2 GPUs on the box): the slab schedule over NVLink peer memory against one GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_schedule_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU on this box")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(REPO, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == world
