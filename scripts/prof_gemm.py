"""one trailing update of the headline factorization as a stand-alone launch (for ncu): the first full update of
n = 32768 with nb = 256 is C[32512 x 32256] -= L21[32512 x 256] U12[256 x 32256] inside the n x n matrix (ld = n).
usage: prof_gemm.py [n] [nb]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import linearsolve_jl_b200 as ls
C = ls._capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
h = ls.Handle(C.F64)
h.set_option(C.OPT_GEMM_CFG, 1 if n >= 12288 else 0)      # what the factorization of this size picks
A = torch.empty((n, n), dtype=torch.float64, device="cuda:0")     # column-major, ld = n
h.fill_uniform_device(A.data_ptr(), n, n, n, seed=3)
es = 8
M, N, K = n - nb, n - 2 * nb, nb
L21 = A.data_ptr() + es * nb                       # rows nb.., columns 0..nb
U12 = A.data_ptr() + es * (2 * nb * n)             # rows 0..nb, columns 2nb..
C22 = A.data_ptr() + es * (2 * nb * n + nb)        # rows nb.., columns 2nb..
for _ in range(3):
    h.debug_gemm_sub(M, N, K, L21, n, U12, n, C22, n)
torch.cuda.synchronize()
print("M N K", M, N, K, "algorithmic bytes", 2 * M * N * es + (M + N) * K * es)
