"""Trailing-update kernels against a plain PyTorch matmul of the same op (C -= A @ B), through
the C ABI test hook `b200lu_debug_gemm_sub`.  FP32: the tcgen05 (TMA + TMEM) 3xTF32 kernel must
reach FP32-level accuracy (tolerance stated below); FP64: the DMMA kernel against torch FP64.
Reference op: `_blocked_lu_schur!`, src/blocked_lufact.jl:186-620."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(ls, dtype_code, tdtype, M, N, K, sgemm_mode=None):
    import torch
    C = ls._capi
    dev = torch.device("cuda", 0)
    h = ls.Handle(dtype_code)
    if sgemm_mode is not None:
        h.set_option(C.OPT_SGEMM_MODE, sgemm_mode)
    g = torch.Generator(device=dev).manual_seed(1234 + M + 7 * N + 13 * K)
    lda = ((M + 15) // 16) * 16
    # column-major storage: torch tensors of shape (cols, ld) hold one column per row
    A = torch.zeros((K, lda), dtype=tdtype, device=dev)
    A[:, :M] = torch.rand((K, M), dtype=tdtype, device=dev, generator=g) - 0.5
    B = torch.rand((N, K), dtype=tdtype, device=dev, generator=g) - 0.5          # ldb = K
    Cm = torch.zeros((N, lda), dtype=tdtype, device=dev)
    Cm[:, :M] = torch.rand((N, M), dtype=tdtype, device=dev, generator=g) - 0.5
    C0 = Cm.clone()
    h.debug_gemm_sub(M, N, K, A.data_ptr(), lda, B.data_ptr(), K, Cm.data_ptr(), lda)
    torch.cuda.synchronize()
    # reference in FP64: C0 - A B  (math: C[m, n] -= sum_k A[m, k] B[k, n])
    Ad = A[:, :M].double().T      # M x K
    Bd = B.double().T             # K x N
    ref = C0[:, :M].double().T - Ad @ Bd
    got = Cm[:, :M].double().T
    # padding rows of C must be untouched
    assert torch.equal(Cm[:, M:], C0[:, M:])
    scale = (Ad.abs() @ Bd.abs()).max().item() + 1.0
    return (got - ref).abs().max().item() / scale


@pytest.mark.parametrize("shape", [(128, 256, 32), (1024, 1024, 256), (1000, 777, 256), (4096, 2048, 192),
                                   (640, 520, 40), (2049, 513, 256)])
def test_sgemm_tcgen05_3xtf32(gpu_required, ls, shape):
    """FP32 trailing update on tcgen05: error relative to |A||B| within 8 eps32 (an FP32 FFMA dot
    product of length 256 is allowed K * eps32; plain single-pass TF32 would be ~1e-3)."""
    M, N, K = shape
    import torch
    err = _run(ls, ls._capi.F32, torch.float32, M, N, K, sgemm_mode=2)   # 2 = tcgen05 whatever the size
    assert err < 8 * np.finfo(np.float32).eps, err


@pytest.mark.parametrize("shape", [(1024, 1024, 256), (1000, 777, 256)])
def test_sgemm_ffma(gpu_required, ls, shape):
    M, N, K = shape
    import torch
    err = _run(ls, ls._capi.F32, torch.float32, M, N, K, sgemm_mode=1)
    assert err < 8 * np.finfo(np.float32).eps, err


@pytest.mark.parametrize("shape", [(1024, 1024, 256), (1000, 777, 250), (130, 70, 34)])
def test_dgemm_dmma(gpu_required, ls, shape):
    M, N, K = shape
    import torch
    err = _run(ls, ls._capi.F64, torch.float64, M, N, K)
    assert err < 8 * np.finfo(np.float64).eps, err


def test_sgemm_tcgen05_more_than_65535_columns(gpu_required, ls):
    """The FP32 / MIXED trailing update of a factorization with n > 65535 + 2 nb has more columns than gridDim.y
    allows: the operand-split kernel strides over the columns (round-1 finding: 'invalid configuration argument')."""
    import torch
    err = _run(ls, ls._capi.F32, torch.float32, 256, 70000, 128, sgemm_mode=2)
    assert err < 8 * np.finfo(np.float32).eps, err
