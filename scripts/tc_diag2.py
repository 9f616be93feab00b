import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import linearsolve_jl_b200 as ls
C = ls._capi
dev = torch.device("cuda", 0)
h = ls.Handle(C.F32); h.set_option(C.OPT_SGEMM_MODE, 2)
M, N, K = 128, 256, 32
lda = 128
A = torch.zeros((M, K), device=dev); A[:, :] = (torch.arange(M, device=dev).float()[:, None] + 1) + 1000 * torch.arange(K, device=dev).float()[None, :]
B = torch.zeros((K, N), device=dev); B[:, :] = (torch.arange(K, device=dev).float()[:, None] + 1) + 100 * torch.arange(N, device=dev).float()[None, :]
At = A.T.contiguous(); Bt = B.T.contiguous(); Ct = torch.zeros((N, lda), device=dev)
h.debug_gemm_sub(M, N, K, At.data_ptr(), lda, Bt.data_ptr(), K, Ct.data_ptr(), lda)
torch.cuda.synchronize()
print("C[0,:4]", (-Ct[:4, 0]).tolist(), "expect", (A @ B)[0, :4].tolist())
