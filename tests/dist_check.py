"""Multi-GPU check, run under torchrun (one rank per GPU):
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_check.py [n]
Block-cyclic distributed getrf must reproduce the single-GPU factorization BITWISE on
every rank's own column blocks (same kernels, same per-element FMA order), with the same
pivots and info; solve_dist must agree with the single-GPU solve."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linearsolve_jl_b200 as ls  # noqa: E402

C = ls._capi


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    # NCCL id: rank 0 creates, torch.distributed broadcasts the 128 bytes
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(C.Handle.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    hd = C.Handle(C.F64, device=local)
    hd.set_option(C.OPT_NB, nb)
    hd.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    nloc = hd.dist_local_cols(n)
    lda = n + (n % 2)          # 16-byte multiple
    Aloc = torch.empty((nloc, lda), dtype=torch.float64, device=dev)
    hd.fill_uniform_device(Aloc.data_ptr(), lda, n, nloc, seed=99, first_global_col=rank * nb,
                           col_block=nb, col_block_stride=world * nb)
    # single-GPU reference on the same matrix
    hs = C.Handle(C.F64, device=local)
    hs.set_option(C.OPT_NB, nb)
    Afull = torch.empty((n, n), dtype=torch.float64, device=dev)
    hs.fill_uniform_device(Afull.data_ptr(), n, n, n, seed=99)
    info_s = hs.factor_device(Afull.data_ptr(), n, n)
    LU = torch.from_numpy(hs.get_factors()).to(dev)           # [i, j]
    ipiv_s = hs.get_ipiv()
    info_d = hd.factor_dist(Aloc.data_ptr(), n, lda)
    torch.cuda.synchronize()
    assert info_s == info_d == 0, (info_s, info_d)
    assert np.array_equal(hd.get_ipiv(), ipiv_s)
    # compare my column blocks bitwise
    nblk = (n + nb - 1) // nb
    lc = 0
    maxdiff = 0.0
    for g in range(rank, nblk, world):
        jb = min(nb, n - g * nb)
        mine = Aloc[lc:lc + jb, :n]                            # [local col, row]
        ref = LU[:, g * nb:g * nb + jb].T                      # [col, row]
        d = (mine - ref).abs().max().item()
        maxdiff = max(maxdiff, d)
        lc += jb
    assert maxdiff == 0.0, f"rank {rank}: factors differ from single-GPU by {maxdiff}"
    # solve
    b = torch.empty((3, n), dtype=torch.float64, device=dev)
    hs.fill_uniform_device(b.data_ptr(), n, n, 3, seed=7)
    xs = torch.empty_like(b)
    xd = torch.empty_like(b)
    hs.solve_device(b.data_ptr(), n, xs.data_ptr(), n, 3)
    transport = hd.dist_transport()
    assert transport == os.environ.get("B200LU_EXPECT_TRANSPORT", transport), transport
    if transport == "p2p":
        # distributed getrs: the factors stay distributed, the right-hand side travels (not the same
        # summation order as the single-GPU sweeps: compared through the backward error)
        for _ in range(2):
            hd.solve_dist(b.data_ptr(), n, xd.data_ptr(), n, 3)
        torch.cuda.synchronize()
        assert (xs - xd).abs().max().item() <= 1e3 * n * 2.2e-16 * xs.abs().max().item()
        x1 = torch.empty(n, dtype=torch.float64, device=dev)       # one right-hand side: the fused step kernel
        for _ in range(2):
            hd.solve_dist(b[2].data_ptr(), n, x1.data_ptr(), n, 1)
        torch.cuda.synchronize()
        assert (xs[2] - x1).abs().max().item() <= 1e3 * n * 2.2e-16 * xs.abs().max().item()
    else:
        xd.copy_(xs)   # the NCCL fallback transport has no distributed getrs
    A = torch.empty((n, n), dtype=torch.float64, device=dev)
    hs.fill_uniform_device(A.data_ptr(), n, n, n, seed=99)
    r = (A.T @ xd[0] - b[0]).norm() / (A.norm() * xd[0].norm())
    assert r.item() <= 10 * n * np.finfo(np.float64).eps
    dist.barrier()
    if rank == 0:
        print(f"dist_check ok: n={n} nb={nb} ranks={world} transport={transport} factors bitwise equal, "
              f"backward error {r.item():.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
