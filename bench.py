#!/usr/bin/env python
"""bench.py — the dense-LU hot path on B200 (BASELINE.json metric: FP64 LU GFLOP/s counted as
2/3 n^3 at n = 8192 / 32768 on 1/2/4/8 B200; batched solves/sec).

Default workload (no flags beyond --gpus/--steps/--warmup): the headline size of the metric,
    ONE FP64 system, n = 32768:  getrf  +  getrs of one right-hand side  (= one `solve!` of a
    fresh LinearProblem), STRONG scaling over the GPUs:
      N = 1 : the single-GPU path (b200lu_factor_device / b200lu_solve_device);
      N > 1 : one rank per GPU under torchrun, 1-D block-cyclic columns, the factored panel handed
              over by peer stores into mapped windows (NCCL = plumbing / fallback), look-ahead,
              distributed getrs (b200lu_factor_dist / b200lu_solve_dist).  Before the timed region
              the same run checks the distributed factors BITWISE against the single-GPU
              factorization (n = 4096) and, after it, the backward error at the timed size.
  `value`   : 2/3 n^3 / (device time of getrf + getrs, CUDA events on the launching streams, max
              over ranks), inputs resident in HBM.
  `e2e`     : the same through the public API (init / cache.A = / solve!) with PINNED HOST buffers —
              H2D of A and b, D2H of x and the pivots inside the timed region; at N > 1 rank 0
              drives all N GPUs through ONE multi-GPU handle (`B200LUFactorization(devices = 0:N-1)`),
              the other ranks idle.
  sub-dicts : `config2_n8192` (BASELINE config 2: n = 8192, cache reuse over 100 right-hand sides;
              N = 1), `batched_65536x64` (config 4, sharded by batch index), `e2e_pageable`.
  `roofline`: the trailing-update DMMA GEMM bracketed in situ by CUDA events on its launching
              stream (B200LU_OPT_PROFILE) against the FP64 tensor peak measured on this GPU by the
              library's register-resident DMMA probe (MEASURED_PEAKS.json has no FP64 entry).
  `cpu_baseline` / `--impl reference`: LAPACK dgetrf + dgetrs (scipy OpenBLAS: the arithmetic of the
              reference's LUFactorization; Julia cannot run here) on ALL host cores, on a bounded
              sample of the workload (the leading n = 12288 block; GFLOP/s is a rate).
`--workload lu --size 8192` runs config 2 alone, `--workload mixed` config 3, `--workload batched`
config 4, `--workload dist` the block-cyclic path at any size (e.g. config 5: --size 65536).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HEADLINE_N = 32768
CPU_SAMPLE_N = 12288


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--n", "--size", dest="n", type=int, default=None)   # use --size under torchrun (its parser abbreviates --n)
    p.add_argument("--nrhs", type=int, default=None)
    p.add_argument("--workload", default="headline", choices=["headline", "lu", "batched", "mixed", "dist"])
    p.add_argument("--batch", type=int, default=65536)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-comparator", action="store_true")
    p.add_argument("--no-extras", action="store_true", help="skip the config2 / batched / pageable sub-dicts")
    p.add_argument("--nb", type=int, default=0)
    p.add_argument("--lookahead", type=int, default=-1)
    p.add_argument("--rpt", type=int, default=-1)
    p.add_argument("--gemm-cfg", type=int, default=-1)
    p.add_argument("--panel-mode", type=int, default=-1)
    p.add_argument("--sgemm-mode", type=int, default=-1)
    args = p.parse_args()
    if args.workload == "headline":
        args.n = args.n or HEADLINE_N
        args.nrhs = args.nrhs or 1
    elif args.workload == "dist":
        args.n = args.n or HEADLINE_N
        args.nrhs = args.nrhs or 1
    else:
        args.n = args.n or 8192
        if args.nrhs is None:
            args.nrhs = 1 if args.workload == "mixed" else 100
    return args


def lu_flops(n):
    return 2.0 / 3.0 * n ** 3


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region: NVML polled every 10 ms from
    a thread (the timed regions here are fractions of a second: `nvidia-smi -lms` would see 1-2 samples);
    falls back to an `nvidia-smi -lms 100` subprocess when pynvml is not importable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nv, self.stop_flag = index, [], None, None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.hdl = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.thr = threading.Thread(target=self._poll, daemon=True)
            self.thr.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.hdl, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.hdl, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.hdl) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.hdl)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.hdl)
                self.rows.append((sm, mx, pw, rs))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.thr.join(timeout=1)
            sm = [r[0] for r in self.rows]
            reasons = sorted({nm for r in self.rows for nm, bit in self.BITS.items() if r[3] & bit})
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[1] for r in self.rows) if sm else None,
                    "power_w_max": max(r[2] for r in self.rows) if sm else None, "samples": len(sm), "reasons": reasons,
                    "source": "NVML polled every 10 ms during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------- CPU arm ----
def all_host_threads():
    """Pin every BLAS thread pool to ALL host cores, whatever OMP_NUM_THREADS says (torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would silently turn the reference arm into a 1-core run).
    Returns (context manager, core count)."""
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0)) or cores
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=cores), cores
    except Exception:
        import contextlib
        return contextlib.nullcontext(), cores


def blas_threads_in_use():
    try:
        from threadpoolctl import threadpool_info
        return max([d.get("num_threads", 1) for d in threadpool_info() if d.get("user_api") == "blas"] + [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_lu_step(A, B, sequential=False):
    """the reference CPU path's arithmetic: LAPACK dgetrf + dgetrs (all BLAS threads).  sequential: the
    right-hand sides one dgetrs call each (what `cache.b = b_i; solve!(cache)` does in the reference)."""
    from scipy.linalg import lapack
    lu, piv, info = lapack.dgetrf(A, overwrite_a=False)
    if sequential:
        x = None
        for c in range(B.shape[1]):
            x, info2 = lapack.dgetrs(lu, piv, B[:, c])
        return x
    x, info2 = lapack.dgetrs(lu, piv, B)
    return x


def workload_name(args):
    n, nrhs = args.n, args.nrhs
    if args.workload in ("headline", "dist"):
        return f"f64 getrf n={n} + getrs {nrhs} rhs (one solve! of a fresh LinearProblem), one system over all GPUs"
    if args.workload == "mixed":
        return f"f32-factor + f64-refinement getrf n={n} + getrs {nrhs} rhs (mixed-precision LU)"
    return f"f64 getrf n={n} + getrs {nrhs} rhs (LinearCache reuse)"


METRIC = "FP64 LU GFLOP/s (2/3 n^3), getrf + getrs"


def run_reference(args, rank, world):
    """--impl reference: LAPACK on the host cores, same config/metric.  Rank 0 only."""
    if rank != 0:
        return
    n, nrhs = args.n, args.nrhs
    ns = min(n, CPU_SAMPLE_N)
    limiter, cores = all_host_threads()
    rng = np.random.default_rng(123)
    A = np.asfortranarray(rng.random((ns, ns)))
    B = np.asfortranarray(rng.random((ns, nrhs)))
    with limiter:
        used = blas_threads_in_use()
        for _ in range(args.warmup):
            cpu_lu_step(A, B)
        ts = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            cpu_lu_step(A, B)
            ts.append(time.perf_counter() - t0)
    t = float(np.mean(ts))
    val = lu_flops(ns) / t / 1e9
    sample = (f"per step: dgetrf + dgetrs({nrhs} rhs) of the leading n={ns} block of the n={n} workload "
              f"(GFLOP/s is a rate; the full size would take ~{lu_flops(n) / (val * 1e9):.0f} s per step); "
              "scipy OpenBLAS = the arithmetic of the reference's LUFactorization (Julia not runnable here)")
    line = {
        "impl": "reference", "metric": METRIC,
        "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.workload in ("headline", "dist") else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "n": n, "nrhs": nrhs},
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": used, "host_cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_entry(n, nrhs, sequential=False):
    ns = min(n, CPU_SAMPLE_N)
    limiter, cores = all_host_threads()
    rng = np.random.default_rng(123)
    A = np.asfortranarray(rng.random((ns, ns)))
    B = np.asfortranarray(rng.random((ns, nrhs)))
    with limiter:
        used = blas_threads_in_use()
        cpu_lu_step(A, B, sequential)   # warm-up (thread pool, page faults)
        t0 = time.perf_counter()
        cpu_lu_step(A, B, sequential)
        tc = time.perf_counter() - t0
    return {"value": lu_flops(ns) / tc / 1e9, "unit": "GFLOP/s", "cores": used, "host_cores": cores, "kind": "port",
            "sample": f"dgetrf + {'%d x dgetrs(1 rhs)' % nrhs if sequential else 'dgetrs(%d rhs)' % nrhs} once on the leading n={ns} block "
                      f"of the n={n} workload (LAPACK via scipy OpenBLAS), after one warm-up", "seconds": tc}


def read_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def newest_profile_traffic(pattern, key_read="dram__bytes_read.sum", key_write="dram__bytes_write.sum"):
    """dram bytes of ONE launch from the newest committed ncu capture matching `pattern` under profiles/
    (None when there is none): the roofline's `traffic` is read from the evidence, not typed in."""
    import glob
    import re
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", pattern))):
        try:
            txt = open(f).read()
        except Exception:
            continue
        tot, found = 0.0, 0
        for key in (key_read, key_write):
            m = re.search(re.escape(key) + r"\s+(\w+)\s+([0-9.,]+)", txt)
            if m:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(1), None)
                if scale:
                    tot += float(m.group(2).replace(",", "")) * scale
                    found += 1
        if found == 2:
            m = re.search(r"algorithmic_bytes\s+-\s+([0-9]+)", txt)
            s = re.search(r"shape\s+-\s+(\S+)", txt)
            best = {"bytes": tot, "file": os.path.relpath(f, ROOT), "algorithmic": float(m.group(1)) if m else None,
                    "shape": s.group(1) if s else None}
    return best


# ------------------------------------------------------------------------- single GPU ----
def bench_single(args, ls, torch, dev, local, n, nrhs, workload, steps, warmup, sample_clocks=True, with_roofline=True,
                 with_e2e=True, with_comparator=True, e2e_sequential=True):
    """One FP64 (or mixed) system on ONE GPU: device-timed step, in-situ roofline, e2e through the API."""
    C = ls._capi
    dtype_code = C.MIXED if workload == "mixed" else C.F64
    h = ls.Handle(dtype_code, device=local)
    for opt, v in ((C.OPT_NB, args.nb or None), (C.OPT_LOOKAHEAD, args.lookahead if args.lookahead >= 0 else None),
                   (C.OPT_PANEL_RPT, args.rpt if args.rpt >= 0 else None), (C.OPT_GEMM_CFG, args.gemm_cfg if args.gemm_cfg >= 0 else None),
                   (C.OPT_PANEL_MODE, args.panel_mode if args.panel_mode >= 0 else None),
                   (C.OPT_SGEMM_MODE, args.sgemm_mode if args.sgemm_mode >= 0 else None)):
        if v is not None:
            h.set_option(opt, v)
    shift = 5.0 if workload == "mixed" else 0.0
    A_dev = torch.empty((n, n), dtype=torch.float64, device=dev)      # column-major n x n (lda = n)
    B_dev = torch.empty((nrhs, n), dtype=torch.float64, device=dev)   # nrhs columns of length n
    X_dev = torch.empty_like(B_dev)
    h.fill_uniform_device(A_dev.data_ptr(), n, n, n, seed=123, diag_shift=shift)
    h.fill_uniform_device(B_dev.data_ptr(), n, n, nrhs, seed=977)

    def step_device():
        info = h.factor_device(A_dev.data_ptr(), n, n)
        t_f = h.timing(C.T_FACTOR) + h.timing(C.T_H2D)
        h.solve_device(B_dev.data_ptr(), n, X_dev.data_ptr(), n, nrhs)
        return info, t_f, h.timing(C.T_SOLVE)

    # roofline yardstick, probed BEFORE the load (cool GPU) and again after it: the larger one is the peak
    peak_dmma_before = h.probe_peak(C.PEAK_FP64_DMMA) if with_roofline else 0.0
    for _ in range(warmup):
        step_device()
    sampler = ClockSampler(local) if sample_clocks else None
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    l0 = ls.launch_count()
    t0 = time.perf_counter()
    tf = ts = 0.0
    for _ in range(steps):
        info, a, b = step_device()
        tf += a
        ts += b
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    launches = ls.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    assert info == 0
    dev_ms = (tf + ts) / steps
    value = lu_flops(n) / (dev_ms * 1e-3) / 1e9

    # residual check of the last step (not timed), on the device: the answer must be right
    x0, b0 = X_dev[0], B_dev[0]
    r = torch.mv(A_dev.t(), x0) - b0          # A_dev[j, i] = entry (i, j)
    berr = float(r.norm() / (A_dev.norm() * x0.norm()))
    assert berr <= 10 * n * np.finfo(np.float64).eps, f"backward error {berr}"

    peaks = read_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    out = {"value": value, "ms_per_step": dev_ms, "getrf_ms": tf / steps, "getrs_ms": ts / steps,
           "wall_ms_per_step": wall / steps * 1e3, "getrf_gflops": lu_flops(n) / ((tf / steps) * 1e-3) / 1e9,
           "backward_error": berr, "gpu_launches": int(launches), "clocks": clocks,
           "nb": h.get_option(C.OPT_NB), "lookahead": h.get_option(C.OPT_LOOKAHEAD)}

    if with_roofline:
        h.set_option(C.OPT_PROFILE, 1)
        g_ms = g_fl = g_n = 0.0
        nprof = 2 if n <= 16384 else 1
        for _ in range(nprof):
            h.factor_device(A_dev.data_ptr(), n, n)
            g_ms += h.timing(C.T_GEMM); g_fl += h.counter(C.C_GEMM_FLOPS); g_n += h.counter(C.C_GEMM_LAUNCHES)
            t_fact_prof = h.timing(C.T_FACTOR)
        h.set_option(C.OPT_PROFILE, 0)
        peak_dmma_after = h.probe_peak(C.PEAK_FP64_DMMA)
        peak_dmma = max(peak_dmma_before, peak_dmma_after)
        achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        if workload == "mixed":
            bf16 = peaks.get("bf16_tflops", 1590.0)
            tc_peak = bf16 / 2.0 / 3.0
            tr = newest_profile_traffic("*ncu_sgemm3x*details*.txt")
            roofline = {
                "bound": "tensor", "kernel": "sgemm3x_tc_kernel (tcgen05 kind::tf32, TMA-fed, TMEM accumulator, 3xTF32)",
                "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s", "frac": achieved / tc_peak,
                "traffic": tr["bytes"] if tr else None, "traffic_source": tr["file"] if tr else None,
                "peak_source": "measured bf16 burst %.0f TFLOP/s / 2 (tf32 rate) / 3 (MMAs per FP32-accurate product); "
                               "achieved counts 2MNK useful flops" % bf16,
                "launches_profiled": int(g_n), "gemm_share_of_getrf": (g_ms / nprof) / t_fact_prof if t_fact_prof else None,
            }
        else:
            tr = newest_profile_traffic("*ncu_dgemm*metrics*.txt")
            roofline = {
                "bound": "tensor", "kernel": "dgemm_sub_kernel (FP64 DMMA trailing update)",
                "achieved": achieved, "peak": peak_dmma, "unit": "TFLOP/s", "frac": achieved / peak_dmma if peak_dmma else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (the first full trailing update of this
                # size as a stand-alone launch, scripts/prof_gemm.py), parsed from the newest committed ncu capture
                # of this kernel under profiles/ together with that launch's shape and algorithmic bytes
                "traffic": tr["bytes"] if tr else None, "traffic_source": tr["file"] if tr else None,
                "traffic_algorithmic_of_that_launch": tr["algorithmic"] if tr else None,
                "traffic_launch_shape": tr["shape"] if tr else None,
                "peak_source": "library DMMA.8x8x4 register-resident probe on this GPU (no FP64 entry in MEASURED_PEAKS.json)",
                "launches_profiled": int(g_n), "gemm_share_of_getrf": (g_ms / nprof) / t_fact_prof if t_fact_prof else None,
                "dfma_probe_tflops": h.probe_peak(C.PEAK_FP64_DFMA), "hbm_copy_probe_gbs": h.probe_peak(C.PEAK_HBM_COPY),
                "peak_probe_before": peak_dmma_before, "peak_probe_after": peak_dmma_after,
                "getrf_frac_of_fp64_peak": (lu_flops(n) / ((tf / steps) * 1e-3) / 1e12) / peak_dmma if peak_dmma else None,
            }
        t_solve = (ts / steps) * 1e-3
        fbytes = (8.0 if workload != "mixed" else 4.0) * n * n
        if nrhs == 1:
            roofline["getrs"] = {"bound": "hbm", "achieved": fbytes / t_solve / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": fbytes / t_solve / 1e9 / hbm_peak, "form": "ONE right-hand side (a vector b)",
                                 "kernel": "trsv3_kernel (cluster chain over DSMEM)" if n >= 6144 else "trsv2_kernel (2-D work items)",
                                 "note": "every factor entry read once per right-hand side; MIXED adds refinement sweeps"}
        else:
            roofline["getrs"] = {"bound": "tensor", "achieved": 2.0 * n * n * nrhs / t_solve / 1e12, "unit": "TFLOP/s",
                                 "peak": peak_dmma if workload != "mixed" else None,
                                 "form": f"ONE n x {nrhs} matrix right-hand side (blocked TRSM); the sequential form is timed in e2e",
                                 "note": "diagonal 1024-blocks by the block-row kernel, off-diagonal updates (2 n^2 nrhs flops) "
                                         "on the trailing-update GEMM"}
        out["roofline"] = roofline

    if with_comparator and workload != "mixed":
        try:
            Am = A_dev.t().contiguous()          # row-major copy of the math matrix for torch
            torch.linalg.lu_factor(Am)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e30
            for _ in range(2):
                e0.record()
                torch.linalg.lu_factor(Am)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out["comparators"] = {"torch_linalg_lu_factor_ms": best,
                                  "torch_linalg_lu_factor_gflops": lu_flops(n) / (best * 1e-3) / 1e9,
                                  "note": "cuSOLVER getrf behind torch.linalg.lu_factor on the same matrix, device-resident, "
                                          "min of 2; library code, not part of the product path"}
            del Am
        except Exception as ex:   # comparator only
            out["comparators"] = {"error": str(ex)[:200]}

    if with_e2e:
        out.update(bench_e2e(ls, torch, A_dev, B_dev, n, nrhs, devices=None, device=local, steps=max(1, min(steps, 3)),
                             sequential=e2e_sequential, mixed=(workload == "mixed")))
    del A_dev, B_dev, X_dev
    h.close()
    torch.cuda.empty_cache()
    return out


def bench_e2e(ls, torch, A_dev, B_dev, n, nrhs, devices, device, steps, sequential=True, mixed=False, pageable=False,
              host_register=False):
    """End to end through the public API (init / cache.A = / cache.b = / solve!) with host buffers:
    H2D of A and every b, D2H of every x and of the pivots inside the timed region."""
    if pageable:
        A_store = np.empty((n, n), dtype=np.float64)                   # plain pageable memory: Julia's rand(n, n)
        A_store[...] = A_dev.cpu().numpy()
        A_host = A_store.T
        B_host = np.array(B_dev.cpu().numpy())
    else:
        A_pin = torch.empty((n, n), dtype=torch.float64).pin_memory()
        A_pin.copy_(A_dev.cpu())
        A_host = A_pin.numpy().T            # Fortran-ordered view: [i, j] = entry (i, j)
        B_pin = torch.empty((nrhs, n), dtype=torch.float64).pin_memory()
        B_pin.copy_(B_dev.cpu())
        B_host = B_pin.numpy()
    if mixed:
        alg = ls.B200LU32MixedLUFactorization(device=device)
    else:
        alg = ls.B200LUFactorization(device=device, devices=devices, host_register=host_register)
    cache = ls.init(ls.LinearProblem(A_host, B_host[0]), alg, alias_A=True, alias_b=True)

    def step_seq():
        cache.A = A_host                 # fresh matrix -> refactor (H2D of A inside)
        for r in range(nrhs):
            cache.b = B_host[r]          # cache reuse: getrs only (H2D b, D2H x inside)
            sol = ls.solve_(cache)
        return sol

    res = {}
    key = ("e2e_pageable_registered" if host_register else "e2e_pageable") if pageable else "e2e"
    t0 = time.perf_counter()
    step_seq()
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0      # with host_register: the one-time cudaHostRegister of A is in here
    t0 = time.perf_counter()
    for _ in range(steps):
        sol = step_seq()
    torch.cuda.synchronize()
    te = (time.perf_counter() - t0) / steps
    assert sol.retcode == ls.ReturnCode.Success
    res[key] = {"value": lu_flops(n) / te / 1e9, "unit": "GFLOP/s",
                "h2d_bytes_per_step": n * n * 8 + nrhs * n * 8,
                "d2h_bytes_per_step": nrhs * n * 8 + n * 8, "ms_per_step": te * 1e3, "steps": steps,
                "host_memory": ("pageable (numpy), page-locked once by the library (B200LU_OPT_HOST_REGISTER)" if host_register
                                else "pageable (numpy)") if pageable else "pinned",
                "first_call_ms": t_first * 1e3,
                "api": "init(LinearProblem) ; cache.A = A ; %s(cache.b = b_i ; solve!(cache))" % (f"{nrhs} x " if nrhs > 1 else "")}
    if devices:
        res[key]["api"] += f" with B200LUFactorization(devices = 0:{len(devices) - 1}) — one process drives all GPUs"
    # the same work with the right-hand sides as ONE n x nrhs matrix b (one solve!, getrs as a blocked TRSM)
    if nrhs > 1 and not pageable and sequential:
        Bm_host = B_host.T               # (n, nrhs) Fortran-ordered view of the pinned buffer
        cache_m = ls.init(ls.LinearProblem(A_host, Bm_host), alg, alias_A=True, alias_b=True)

        def step_m():
            cache_m.A = A_host
            return ls.solve_(cache_m)

        step_m()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            solm = step_m()
        torch.cuda.synchronize()
        tm = (time.perf_counter() - t0) / steps
        assert solm.retcode == ls.ReturnCode.Success
        res["e2e_matrix_rhs"] = {"value": lu_flops(n) / tm / 1e9, "unit": "GFLOP/s", "ms_per_step": tm * 1e3,
                                 "api": "init(LinearProblem(A, B::Matrix n x nrhs)) ; cache.A = A ; solve!(cache)"}
        del cache_m
    del cache
    return res


def bench_batched_quick(ls, torch, dev, local, batch, steps, warmup, seed0=5, fused=True):
    """BASELINE config 4 (shard of `batch` 64x64 FP64 systems on this GPU): device-timed factor + solve."""
    C = ls._capi
    n = 64
    h = ls.Handle(C.F64, device=local)
    A_dev = torch.empty((batch, n, n), dtype=torch.float64, device=dev)
    b_dev = torch.empty((batch, n), dtype=torch.float64, device=dev)
    x_dev = torch.empty_like(b_dev)
    h.fill_uniform_device(A_dev.data_ptr(), n, n, batch * n, seed=seed0, diag_shift=0.0)
    A_dev += 64.0 * torch.eye(n, device=dev, dtype=torch.float64)
    h.fill_uniform_device(b_dev.data_ptr(), n, n, batch, seed=seed0 + 1)

    def step():
        if fused:   # factor + the first solve in ONE kernel, factors written out and kept
            bad = h.factor_solve_batched_device(A_dev.data_ptr(), b_dev.data_ptr(), x_dev.data_ptr(), batch, n)
            return bad, h.timing(C.T_FACTOR), 0.0
        bad = h.factor_batched_device(A_dev.data_ptr(), batch, n)
        tf = h.timing(C.T_FACTOR)
        h.solve_batched_device(b_dev.data_ptr(), x_dev.data_ptr(), 1)
        return bad, tf, h.timing(C.T_SOLVE)

    for _ in range(warmup):
        step()
    l0 = ls.launch_count()
    tf = ts = 0.0
    for _ in range(steps):
        bad, a, b = step()
        tf += a; ts += b
    launches = ls.launch_count() - l0
    assert bad == 0
    r = torch.einsum("sji,sj->si", A_dev, x_dev) - b_dev
    assert float(r.abs().max()) < 1e-10
    # a later solve with the kept factors (LinearCache reuse) must still work and agree
    x2 = torch.empty_like(x_dev)
    h.solve_batched_device(b_dev.data_ptr(), x2.data_ptr(), 1)
    assert float((x2 - x_dev).abs().max()) < 1e-12
    del A_dev, b_dev, x_dev, x2
    h.close()
    torch.cuda.empty_cache()
    return tf / steps, ts / steps, launches


def batched_entry(ms_f, ms_s, per, world, hbm):
    n = 64
    bytes_sys = 2 * n * n * 8 + n * 4 + 2 * n * 8     # A in, LU out, ipiv (int32), b in, x out = 66,816 B
    ms = ms_f + ms_s
    ach = per * bytes_sys / (ms * 1e-3) / 1e9
    ach_f = per * (2 * n * n * 8 + n * 4 + 4) / (ms_f * 1e-3) / 1e9
    # dram bytes of ONE launch of the factor-only kernel on 16384 systems (the committed ncu capture)
    tr = newest_profile_traffic("*ncu_batched_warp*metrics*.txt")
    return {"value": world * per / (ms * 1e-3), "unit": "systems/s", "ms_per_step": ms, "getrf_ms": ms_f, "getrs_ms": ms_s,
            "variant": "factor + solve of one right-hand side in ONE kernel, factors written out and kept (LinearCache contract)"
                       if ms_s == 0.0 else "factor, then solve (two kernels), factors kept",
            "systems_per_gpu": per,
            "roofline": {"bound": "hbm", "kernel": "getrf_batched_warp_kernel<double,64,SOLVE> (warp per system, shared-memory resident)"
                         if ms_s == 0.0 else "getrf_batched_*_kernel + getrs_batched_kernel", "achieved": ach, "peak": hbm,
                         "unit": "GB/s", "frac": ach / hbm, "getrf_only_frac": ach_f / hbm,
                         "algorithmic_bytes_per_system": bytes_sys, "hbm_bound_systems_per_s_per_gpu": hbm * 1e9 / bytes_sys,
                         "traffic": tr["bytes"] if tr else None, "traffic_source": tr["file"] if tr else None,
                         "traffic_launch": "getrf_batched_warp_kernel<double,64> (factor only: A in, LU + pivots out) on 16384 "
                                           "systems: %.0f B per system" % (tr["bytes"] / 16384) if tr else None}}


# --------------------------------------------------------------------------------- main ----
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import linearsolve_jl_b200 as ls
    C = ls._capi

    if not os.path.exists(C.LIB_PATH):
        raise SystemExit("libb200lu.so missing — run `python __graft_entry__.py build` (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.workload == "batched":
        return bench_batched(args, ls, torch, dist, dev, rank, world, local, barrier, max_over_ranks)
    if world > 1:
        # one system over all GPUs: every multi-rank launch measures north_star's block-cyclic design
        return bench_dist(args, ls, torch, dist, dev, rank, world, local, barrier, max_over_ranks)
    if args.workload == "dist":
        return bench_dist(args, ls, torch, dist, dev, rank, world, local, barrier, max_over_ranks)

    n, nrhs = args.n, args.nrhs
    wl = "mixed" if args.workload == "mixed" else "lu"
    res = bench_single(args, ls, torch, dev, local, n, nrhs, wl, args.steps, args.warmup,
                       with_e2e=not args.no_e2e, with_comparator=not args.no_comparator)
    line = {
        "metric": METRIC, "value": res["value"], "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if args.workload == "headline" else "weak", "vs_baseline": None,
        "dtype": "f64" if wl == "lu" else "f32 factor + f64 refine", "data": "synthetic",
        "config": {"workload": workload_name(args), "n": n, "nrhs": nrhs, "multi_gpu": "single GPU",
                   "l2": "inputs (A = %.0f MiB) larger than L2" % (n * n * 8 / 2 ** 20),
                   "nb": res["nb"], "lookahead": res["lookahead"]},
    }
    for k in ("getrf_ms", "getrs_ms", "wall_ms_per_step", "getrf_gflops", "backward_error", "gpu_launches", "clocks",
              "roofline", "comparators", "e2e", "e2e_matrix_rhs"):
        if k in res:
            line[k] = res[k]

    if args.workload == "headline" and not args.no_extras:
        # BASELINE config 2 beside the headline: n = 8192, cache reuse over 100 right-hand sides
        c2 = bench_single(args, ls, torch, dev, local, 8192, 100, "lu", max(3, min(args.steps, 10)), 3, sample_clocks=False,
                          with_roofline=True, with_e2e=not args.no_e2e, with_comparator=not args.no_comparator)
        sub = {"workload": "f64 getrf n=8192 + getrs 100 rhs (LinearCache reuse), BASELINE config 2",
               "value": c2["value"], "unit": "GFLOP/s", "ms_per_step": c2["ms_per_step"],
               "value_form": "device-resident: getrf + ONE blocked TRSM over the 8192 x 100 matrix right-hand side",
               "getrf_ms": c2["getrf_ms"], "getrs_ms": c2["getrs_ms"], "getrf_gflops": c2["getrf_gflops"]}
        for k in ("e2e", "e2e_matrix_rhs", "comparators", "roofline", "backward_error"):
            if k in c2:
                sub[k] = c2[k]
        if "e2e" in sub:
            sub["e2e"]["form"] = "100 sequential solve! calls (the literal cache-reuse loop); e2e_matrix_rhs = one matrix right-hand side"
        if not args.no_cpu_baseline:
            sub["cpu_baseline_sequential"] = cpu_baseline_entry(8192, 100, sequential=True)
            sub["cpu_baseline_matrix_rhs"] = cpu_baseline_entry(8192, 100, sequential=False)
        line["config2_n8192"] = sub
        # BASELINE config 4 beside it
        ms_f, ms_s, _ = bench_batched_quick(ls, torch, dev, local, args.batch, 5, 3)
        line["batched_65536x64"] = batched_entry(ms_f, ms_s, args.batch, 1, read_peaks().get("hbm_gbs", 6650.0))
        # the headline once more from plain pageable host memory (what `rand(n, n)` in Julia is)
        if not args.no_e2e:
            A_dev = torch.empty((n, n), dtype=torch.float64, device=dev)
            B_dev = torch.empty((nrhs, n), dtype=torch.float64, device=dev)
            hh = ls.Handle(C.F64, device=local)
            hh.fill_uniform_device(A_dev.data_ptr(), n, n, n, seed=123)
            hh.fill_uniform_device(B_dev.data_ptr(), n, n, nrhs, seed=977)
            hh.close()
            line.update(bench_e2e(ls, torch, A_dev, B_dev, n, nrhs, None, local, steps=2, pageable=True))
            # ... and with the library page-locking that buffer once (a cache owns its copy of A for its lifetime)
            line.update(bench_e2e(ls, torch, A_dev, B_dev, n, nrhs, None, local, steps=2, pageable=True, host_register=True))
            del A_dev, B_dev

    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_entry(n, nrhs)
    print(json.dumps(line), flush=True)


def bench_dist(args, ls, torch, dist, dev, rank, world, local, barrier, max_over_ranks):
    """ONE FP64 system factored and solved by all ranks — 1-D block-cyclic columns (nb = 256), the owner
    factors the panel and stores it into its peers' windows, look-ahead, distributed getrs.  Strong
    scaling: `value` = 2/3 n^3 / (max over ranks of the device time of factor_dist + solve_dist)."""
    C = ls._capi
    # outer panel width: the library's own choice by GPU count (256; 128 from eight GPUs on, where the panel chain
    # outweighs the better GEMM shape) unless --nb is given
    n, nrhs = args.n, args.nrhs
    eps = np.finfo(np.float64).eps

    def make_handle():
        h = C.Handle(C.F64, device=local)
        if args.nb:
            h.set_option(C.OPT_NB, args.nb)
        if args.gemm_cfg >= 0:
            h.set_option(C.OPT_GEMM_CFG, args.gemm_cfg)
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if world > 1:
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(C.Handle.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
        h.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
        return h

    h = make_handle()
    nb = h.get_option(C.OPT_NB)
    transport = h.dist_transport() if world > 1 else "single rank"

    # ---- in-run self-check (the driver's test box has one GPU): distributed factors == single-GPU factors, bitwise
    def self_check(nc):
        nloc = h.dist_local_cols(nc)
        Aloc = torch.empty((max(nloc, 1), nc), dtype=torch.float64, device=dev)
        h.fill_uniform_device(Aloc.data_ptr(), nc, nc, max(nloc, 1), seed=99, first_global_col=rank * nb,
                              col_block=nb, col_block_stride=world * nb)
        hs = C.Handle(C.F64, device=local)
        hs.set_option(C.OPT_NB, nb)
        Af = torch.empty((nc, nc), dtype=torch.float64, device=dev)
        hs.fill_uniform_device(Af.data_ptr(), nc, nc, nc, seed=99)
        assert hs.factor_device(Af.data_ptr(), nc, nc) == 0
        LU = torch.from_numpy(hs.get_factors()).to(dev)
        assert h.factor_dist(Aloc.data_ptr(), nc, nc) == 0
        same = bool(np.array_equal(h.get_ipiv(), hs.get_ipiv()))
        lc = 0
        for g in range(rank, -(-nc // nb), world):
            jb = min(nb, nc - g * nb)
            same = same and bool(torch.equal(Aloc[lc:lc + jb, :nc], LU[:, g * nb:g * nb + jb].T))
            lc += jb
        hs.close()
        t = torch.tensor([1.0 if same else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() == 1.0)

    check_bitwise = self_check(4096)
    torch.cuda.empty_cache()
    peak_before = h.probe_peak(C.PEAK_FP64_DMMA)      # the yardstick on a cool GPU; probed again after the load

    nloc = h.dist_local_cols(n)
    Aloc = torch.empty((nloc, n), dtype=torch.float64, device=dev)
    B_dev = torch.empty((nrhs, n), dtype=torch.float64, device=dev)
    X_dev = torch.empty_like(B_dev)
    h.fill_uniform_device(B_dev.data_ptr(), n, n, nrhs, seed=977)

    def fill():
        h.fill_uniform_device(Aloc.data_ptr(), n, n, nloc, seed=123, first_global_col=rank * nb,
                              col_block=nb, col_block_stride=world * nb)

    can_solve = transport != "nccl"

    def step():
        fill()
        barrier()
        info = h.factor_dist(Aloc.data_ptr(), n, n)
        tf = h.timing(C.T_FACTOR)
        ts = 0.0
        if can_solve:
            h.solve_dist(B_dev.data_ptr(), n, X_dev.data_ptr(), n, nrhs)
            ts = h.timing(C.T_SOLVE)
        return info, tf, ts

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ls.launch_count()
    tf = ts = 0.0
    for _ in range(args.steps):
        info, a, b = step()
        tf += a
        ts += b
    clocks = sampler.stop()
    launches = ls.launch_count() - l0
    assert info == 0
    tf_ms = max_over_ranks(tf / args.steps)
    ts_ms = max_over_ranks(ts / args.steps)
    ms = max_over_ranks((tf + ts) / args.steps)

    # ---- backward error at the timed size: r = b - A x with A regenerated block by block on its owner
    berr = None
    if can_solve:
        fill()
        x = X_dev[0]
        cols = torch.from_numpy(ls.block_cyclic_columns(n, nb, rank, world)).to(dev)
        part = torch.mv(Aloc.t(), x[cols]) if nloc > 0 else torch.zeros(n, dtype=torch.float64, device=dev)
        fro2 = (Aloc * Aloc).sum().reshape(1)
        if world > 1:
            dist.all_reduce(part)
            dist.all_reduce(fro2)
        berr = float((part - B_dev[0]).norm() / (fro2.sqrt() * x.norm()))
        assert berr <= 10 * n * eps, f"backward error {berr}"

    # ---- in-situ roofline of the trailing update on this rank's share (one profiled factorization)
    h.set_option(C.OPT_PROFILE, 1)
    fill()
    barrier()
    h.factor_dist(Aloc.data_ptr(), n, n)
    g_ms, g_fl = h.timing(C.T_GEMM), h.counter(C.C_GEMM_FLOPS)
    t_prof = h.timing(C.T_FACTOR)
    chain = [h.timing(C.T_PANEL), h.timing(C.T_LOOKAHEAD), h.timing(C.T_PUSH)]
    if world > 1:
        ct = torch.tensor(chain, dtype=torch.float64, device=dev)
        dist.all_reduce(ct)                      # every panel has one owner: the sum over ranks is the whole chain
        chain = [float(x) for x in ct.tolist()]
    h.set_option(C.OPT_PROFILE, 0)
    peak = max(peak_before, h.probe_peak(C.PEAK_FP64_DMMA))
    ach = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    gemm_share = max_over_ranks(g_ms / t_prof if t_prof else 0.0)
    value = lu_flops(n) / (ms * 1e-3) / 1e9
    trd = newest_profile_traffic("*ncu_dgemm*metrics*.txt")
    line = {"metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "n": n, "nrhs": nrhs, "nb": nb,
                       "multi_gpu": ("1-D block-cyclic columns nb=%d over %d ranks (one process per GPU); panel hand-off: %s; "
                                     "look-ahead on the block each rank factors next; distributed getrs" %
                                     (nb, world, {"p2p": "peer stores into cudaIpc-mapped windows over NVLink",
                                                  "nccl": "ncclBroadcast (fallback transport)"}.get(transport, transport))),
                       "l2": "inputs (A = %.0f MiB per rank) larger than L2" % (nloc * n * 8 / 2 ** 20),
                       "timing": "CUDA events on each rank's main stream around factor_dist + solve_dist, device barrier at entry, max over ranks"},
            "getrf_ms": tf_ms, "getrs_ms": ts_ms, "getrf_gflops": lu_flops(n) / (tf_ms * 1e-3) / 1e9,
            "dist_check": "bitwise" if check_bitwise else "MISMATCH",
            "dist_check_detail": "n=4096: every rank's column blocks and the pivots equal the single-GPU factorization bit for bit"
                                 if check_bitwise else "n=4096: distributed factors differ from the single-GPU factors",
            "backward_error": berr, "transport": transport,
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "dgemm_sub_kernel (FP64 DMMA trailing update, this rank's column blocks)",
                         "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if peak else None,
                         # ncu runs one GPU at a time: the same kernel's stand-alone capture on one GPU
                         "traffic": trd["bytes"] if trd else None, "traffic_source": trd["file"] if trd else None,
                         "traffic_algorithmic_of_that_launch": trd["algorithmic"] if trd else None,
                         "traffic_launch_shape": trd["shape"] if trd else None,
                         "gemm_share_of_getrf_max_over_ranks": gemm_share,
                         "getrf_frac_of_aggregate_fp64_peak": (lu_flops(n) / (tf_ms * 1e-3) / 1e12) / (peak * world) if peak else None,
                         "chain_ms": {"panel_factorizations": chain[0], "lookahead_block_updates": chain[1], "panel_hand_off": chain[2],
                                      "sum": sum(chain), "profiled_getrf_ms": max_over_ranks(t_prof),
                                      "how": "CUDA events on the owners' panel streams in one extra profiled factorization, "
                                             "summed over all panels (every panel has exactly one owner)"},
                         "limiter": "the serial chain push -> look-ahead block update -> panel factorization -> push; "
                                    "the trailing GEMMs (1/N per rank) hide under it"}}
    assert check_bitwise, "distributed factors differ from the single-GPU factorization"
    del Aloc
    h.close()
    torch.cuda.empty_cache()

    # ---- config 4 beside it: the batch sharded by index, no communication
    if not args.no_extras:
        per = args.batch // world
        ms_f, ms_s, _ = bench_batched_quick(ls, torch, dev, local, per, 5, 3, seed0=5 + 2 * rank)
        barrier()
        ms_f, ms_s = max_over_ranks(ms_f), max_over_ranks(ms_s)
        line["batched_65536x64"] = batched_entry(ms_f, ms_s, per, world, read_peaks().get("hbm_gbs", 6650.0))

    # ---- e2e: rank 0 drives ALL the GPUs through one multi-GPU handle from pinned host memory; the other
    # ranks wait on a HOST-side (gloo) barrier, so that no NCCL kernel spins on their GPUs meanwhile
    barrier()
    if not args.no_e2e:
        cpu_group = dist.new_group(backend="gloo") if world > 1 else None
        if rank == 0:
            try:
                A_dev = torch.empty((n, n), dtype=torch.float64, device=dev)
                hh = C.Handle(C.F64, device=local)
                hh.fill_uniform_device(A_dev.data_ptr(), n, n, n, seed=123)
                hh.close()
                line.update(bench_e2e(ls, torch, A_dev, B_dev, n, nrhs, devices=tuple(range(world)), device=local,
                                      steps=max(1, min(args.steps, 3))))
                del A_dev
            except Exception as ex:   # the device-timed line stands on its own
                line["e2e_error"] = str(ex)[:300]
        if cpu_group is not None:
            dist.barrier(group=cpu_group)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_batched(args, ls, torch, dist, dev, rank, world, local, barrier, max_over_ranks):
    """BASELINE config 4: `batch` independent 64x64 FP64 systems sharded by batch index (no communication)."""
    per = args.batch // world
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ms_f, ms_s, launches = bench_batched_quick(ls, torch, dev, local, per, args.steps, args.warmup, seed0=5 + 2 * rank)
    barrier()
    clocks = sampler.stop()
    ms_f, ms_s = max_over_ranks(ms_f), max_over_ranks(ms_s)
    ent = batched_entry(ms_f, ms_s, per, world, read_peaks().get("hbm_gbs", 6650.0))
    line = {"metric": "batched independent 64x64 FP64 solves/s (factor + solve, factors kept)", "value": ent["value"],
            "unit": "systems/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ent["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": f"{args.batch} systems of 64x64 f64, sharded by batch index",
                                            "l2": "inputs (%.0f MiB) larger than L2" % (per * 64 * 64 * 8 / 2 ** 20)},
            "getrf_ms": ms_f, "getrs_ms": ms_s, "gpu_launches": int(launches), "clocks": clocks, "roofline": ent["roofline"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
