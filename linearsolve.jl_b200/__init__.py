"""linearsolve.jl_b200 — a B200-native dense LU factor-and-solve path behind the
LinearSolve.jl interface.

Layout (only what the path needs):
  csrc/        hand-written sm_100a CUDA kernels + the C ABI (libb200lu.so)
  _capi.py     ctypes twin of the Julia `ccall` layer
  interface.py host-side mirror of LinearProblem / init / solve! / LinearCache,
               B200LUFactorization, B200LU32MixedLUFactorization, BlockDiagonal,
               defaultalg
  julia/       the glue file a LinearSolve.jl maintainer adds (not runnable here)

The directory name carries a dot, so import it through the repo-root shim:
    import linearsolve_jl_b200 as ls
"""
from . import _capi
from ._capi import B200LUError, Handle, is_available, launch_count
from .interface import (
    AbstractFactorization,
    B200LU32MixedLUFactorization,
    B200LUFactorization,
    BlockDiagonal,
    DefaultAlgorithmChoice,
    DefaultLinearSolver,
    LinearCache,
    LinearProblem,
    LinearSolution,
    OperatorAssumptions,
    ReturnCode,
    block_cyclic_columns,
    defaultalg,
    init,
    pad_blocks,
    plan_blockdiag,
    reduce_info,
    reinit,
    shard_batch,
    solve,
    solve_,
    successful_retcode,
    useb200,
)

__all__ = [n for n in dir() if not n.startswith("_")]
