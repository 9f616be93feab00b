import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import linearsolve_jl_b200 as ls
        return ls.useb200()
    except Exception:
        return False


@pytest.fixture(scope="session")
def ls():
    import linearsolve_jl_b200 as ls
    return ls


@pytest.fixture(scope="session")
def oracle():
    from oracle import lu_oracle
    lu_oracle.build()
    return lu_oracle


@pytest.fixture(scope="session")
def gpu_required():
    """GPU tests must FAIL (not skip) when the CUDA path is unusable on a GPU box."""
    import linearsolve_jl_b200 as ls
    assert os.path.exists(ls._capi.LIB_PATH), "libb200lu.so missing: run __graft_entry__.build()"
    assert ls.useb200(), "libb200lu.so could not create a handle on this machine (no sm_100 GPU?)"
    return True


def decisive_matrix(rng, n, dtype=np.float64):
    """rows of widely separated magnitude, shuffled: every pivot decision has an
    orders-of-magnitude margin (reference test/Core/blocked_lufact.jl:148-161)"""
    mags = 2.0 ** rng.permutation(np.arange(1, n + 1)).astype(np.float64)
    if dtype == np.float32:
        mags = 2.0 ** (rng.permutation(n) % 60).astype(np.float64) * (1.0 + rng.permutation(n) / n)
    A = np.diag(mags) @ (np.eye(n) + 0.01 * rng.standard_normal((n, n)))
    return np.asfortranarray(A[rng.permutation(n), :].astype(dtype))
