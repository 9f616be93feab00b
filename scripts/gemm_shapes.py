"""trailing-update DGEMM on the shapes one rank of an 8-GPU run sees (M x M/8 x 128) and the single-GPU shapes, for
every tile configuration (B200LU_OPT_GEMM_CFG): wall clock over repeated synchronous calls, TFLOP/s"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import linearsolve_jl_b200 as ls
C = ls._capi
n = 32768
A = torch.empty((n, n // 4), dtype=torch.float64, device="cuda:0").fill_(0.001)   # panel, U block and C all fit
h = ls.Handle(C.F64)
es = 8
for (M, N, K) in ((32640, 4080, 128), (24576, 3072, 128), (16384, 2048, 128), (8192, 1024, 128), (32512, 4064, 256),
                  (16384, 2048, 256), (32512, 8064, 256)):
    L21 = A.data_ptr()
    U12 = A.data_ptr() + es * (256 * n)
    C22 = A.data_ptr() + es * (512 * n)
    out = []
    for cfg in (0, 1, 2):
        h.set_option(C.OPT_GEMM_CFG, cfg)
        for _ in range(3):
            h.debug_gemm_sub(M, N, K, L21, n, U12, n, C22, n)
        torch.cuda.synchronize()
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            h.debug_gemm_sub(M, N, K, L21, n, U12, n, C22, n)
        dt = (time.perf_counter() - t0) / reps
        out.append(f"cfg{cfg} {dt * 1e3:.3f} ms {2.0 * M * N * K / dt / 1e12:.1f} TF/s")
    print(f"M={M} N={N} K={K}: " + " | ".join(out), flush=True)
