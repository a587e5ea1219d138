"""GPU, N > 1: launches tests/dist_check.py under torchrun on as many GPUs as the box shows (2, 4, 8)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("n", [2, 4, 8])
def test_summa_parity_on_n_gpus(n):
    if ngpus() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + n), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "FAIL" not in r.stdout
