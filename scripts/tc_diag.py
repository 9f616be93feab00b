import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import linearsolve_jl_b200 as ls
C = ls._capi
dev = torch.device("cuda", 0)
h = ls.Handle(C.F32); h.set_option(C.OPT_SGEMM_MODE, 2)
def run(A, B, M, N, K):
    # A: (M,K) math, B: (K,N) math
    lda = ((M + 15)//16)*16
    At = torch.zeros((K, lda), dtype=torch.float32, device=dev); At[:, :M] = A.T
    Bt = B.T.contiguous()
    Ct = torch.zeros((N, lda), dtype=torch.float32, device=dev)
    h.debug_gemm_sub(M, N, K, At.data_ptr(), lda, Bt.data_ptr(), K, Ct.data_ptr(), lda)
    torch.cuda.synchronize()
    return -Ct[:, :M].T   # = A @ B expected
M, N, K = 128, 256, 32
A = torch.ones((M, K), device=dev); B = torch.ones((K, N), device=dev)
R = run(A, B, M, N, K); print("ones: min/max", R.min().item(), R.max().item(), "expect", K)
# A[m,k] = m, only k=0 ; B[0,n] = 1
A = torch.zeros((M, K), device=dev); A[:, 0] = torch.arange(M, device=dev).float(); B = torch.zeros((K, N), device=dev); B[0, :] = 1
R = run(A, B, M, N, K); print("A=m at k0: R[:,0][:40]", R[:40, 0].tolist()); print("   R[5,:8]", R[5, :8].tolist())
# which k is paired: A[m,k]=1 only for k=ka ; B[k,n] = k+1
for ka in (0, 1, 7, 8, 9, 31):
    A = torch.zeros((M, K), device=dev); A[:, ka] = 1; B = (torch.arange(K, device=dev).float() + 1)[:, None].repeat(1, N)
    R = run(A, B, M, N, K); print("ka", ka, "-> R[0,0], R[64,100] =", R[0, 0].item(), R[64, 100].item(), "expect", ka + 1)
# n mapping: B[0,n] = n, A[:,0]=1
A = torch.zeros((M, K), device=dev); A[:, 0] = 1; B = torch.zeros((K, N), device=dev); B[0, :] = torch.arange(N, device=dev).float()
R = run(A, B, M, N, K); print("n map R[3,:40]", R[3, :40].tolist())
