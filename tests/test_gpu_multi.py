"""Multi-GPU parity (needs >= 2 GPUs): launches tests/mgpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world,cells", [(2, 16), (2, 64), (4, 16), (8, 16)])
def test_distributed_parity(world, cells):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world + cells
    os.environ["MGPU_CELLS"] = str(cells)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "mgpu_check ok" in out.stdout and "mgpu_check agglomerated ok" in out.stdout
