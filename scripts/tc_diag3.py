import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import linearsolve_jl_b200 as ls
C = ls._capi
dev = torch.device("cuda", 0)
h = ls.Handle(C.F32); h.set_option(C.OPT_SGEMM_MODE, 2)
M, N, K = 128, 256, 32
At = torch.ones((K, M), device=dev); Bt = torch.ones((N, K), device=dev); Ct = torch.zeros((N, M), device=dev)
h.debug_gemm_sub(M, N, K, At.data_ptr(), M, Bt.data_ptr(), K, Ct.data_ptr(), M)
torch.cuda.synchronize()
print("ones: C min/max", Ct.min().item(), Ct.max().item())
