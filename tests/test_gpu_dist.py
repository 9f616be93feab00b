"""Multi-GPU block-cyclic getrf: needs >= 2 GPUs on the box; on a 1-GPU box
the test is skipped (nothing to shard).  The N>1 HOST logic is covered on CPU by
tests/test_dist_host.py (gloo, world_size 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("n,nb", [(2048, 256), (1000, 64)])
def test_block_cyclic_matches_single_gpu(gpu_required, n, nb):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611",
                          os.path.join(ROOT, "tests", "dist_check.py"), str(n), str(nb)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "dist_check ok" in out.stdout
