"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): slab-partitioned LLG steps
through fg_dist_* against the single-GPU path, launched like the bench (torchrun, one rank/GPU)."""
import os
import subprocess
import sys

import pytest

import cases

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_dist_matches_single_gpu(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    worker = os.path.join(cases.ROOT, "tests", "dist_gpu_worker.py")
    cmd = ["timeout", "600", sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1", "--master-port", "29541", worker]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=cases.ROOT,
                       env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert "DIST_GPU_OK" in r.stdout
