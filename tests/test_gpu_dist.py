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


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n,nb,transport", [(2048, 256, "p2p"), (1000, 64, "p2p"), (1001, 128, "p2p"), (1536, 256, "nccl")])
def test_block_cyclic_matches_single_gpu(gpu_required, n, nb, transport):
    """one process per GPU under torchrun: peer-store transport (cudaIpc windows) and the NCCL fallback"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, B200LU_EXPECT_TRANSPORT=transport)
    if transport == "nccl":
        env["B200LU_DIST_MODE"] = "nccl"
    nproc = 3 if (_ngpu() >= 3 and n == 1001) else 2          # an odd rank count when the box has one
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                          os.path.join(ROOT, "tests", "dist_check.py"), str(n), str(nb)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "dist_check ok" in out.stdout
