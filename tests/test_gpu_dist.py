"""Multi-GPU (row-sharded) parity on real GPUs: spawns torchrun over all visible GPUs.
Skipped when fewer than two GPUs are visible."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import logreg_b200
    return logreg_b200.device_count()


@pytest.mark.parametrize("kind", ["p2p", "nccl"])
def test_row_sharded_equals_single_gpu(kind):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(n, 8)
    port = 29500 + (os.getpid() % 500) + (0 if kind == "p2p" else 500)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py"), kind]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout
