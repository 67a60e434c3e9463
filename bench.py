#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native harmonic 3D BEM hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                     (the reference algorithm's CPU arm on the host cores)

Metric (BASELINE.json): "harmonic BEM assembly Gentries/s + LU solves/s at N=30k DOF".  One *step* = one frequency of the
sweep = assemble the 30258 x 30258 complex influence matrix + LU-factorise + solve (src/multifebe.f90:107-124 loop body).
`value` = end-to-end solves/s over all ranks with everything device resident; the assembly throughput (Gentries/s) and the
per-kernel roofline numbers are reported beside it in the same JSON line.  Frequencies are sharded round-robin over the
ranks (weak scaling: every rank assembles and solves its own frequencies; no collective on the data path, one gather of
each solution to the writer rank in the e2e arm).

Workload: synthetic S-cube: unit cube, 6 faces x 40 x 40 cells x 2 tri3 = 19200 elements, 10086 nodes,
30258 DOF, BCs of the reference's ME-TH-EL-001 tutorial, rho = mu = 1, nu = 0.25, xi = 0.03, 64 frequencies.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# The CPU legs (--impl reference, and the cpu_baseline leg of a single-rank run) use every host core: torch.distributed.run exports
# OMP_NUM_THREADS=1 to its workers, which OpenBLAS reads when it is loaded, i.e. at `import numpy` -- so the count is pinned here, before
# that import (round-1 bug: the zgetrf sample of the reference arm ran single-threaded at N > 1).
_REFERENCE_ARM = any(a == "reference" or a.endswith("=reference") for a in sys.argv[1:])
if _REFERENCE_ARM or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(_host_cores())

import numpy as np


def blas_threads():
    """Effective thread count of the BLAS behind scipy.linalg.lapack (threadpoolctl), pinned to the host cores if it is lower."""
    try:
        import scipy.linalg  # noqa: F401  (loads the BLAS the LU leg uses)
        import threadpoolctl
        want = _host_cores()
        info = threadpoolctl.threadpool_info()
        if any(i.get("user_api") == "blas" and i.get("num_threads", want) < want for i in info):
            threadpoolctl.threadpool_limits(limits=want, user_api="blas")
            info = threadpoolctl.threadpool_info()
        return max([i.get("num_threads", 1) for i in info if i.get("user_api") == "blas"] or [1])
    except Exception:
        return None

# NCCL's own log lines (e.g. "NCCL version ..." under NCCL_DEBUG=VERSION) must not land on stdout, which carries the ONE JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # hardware work queues: several problems in flight per GPU (--workload c1), see multifebe_b200/capi.py

METRIC = "harmonic 3D BEM end-to-end solves/s (assemble + zgetrf + zgetrs per frequency) at N=30258 DOF; assembly Gentries/s beside it"
N_FREQ = 64


def workload(args):
    from multifebe_b200.host import Model, Material, cube_mesh, cube_bcs, shape
    et = {"tri3": shape.TRI3, "tri6": shape.TRI6, "quad4": shape.QUAD4, "quad8": shape.QUAD8, "quad9": shape.QUAD9}[args.etype]
    md = Model(cube_mesh(args.m, et), cube_bcs())
    mat = Material(1.0, 1.0, 0.25, 0.03)
    # omega_max such that |k2| * (element size) <= 1 (>= 6 elements per S wavelength): both branches of E_m(z) are exercised
    cl = (2.0 ** 0.5 if et in (shape.TRI3, shape.TRI6) else 1.0) / args.m
    om_max = 1.0 / cl
    from multifebe_b200.sweep import linear_frequencies
    freqs = linear_frequencies(0.05 * om_max, om_max, N_FREQ)
    name = "S-cube %s m=%d: %d elements, %d nodes, %d DOF, %d-frequency sweep" % (args.etype, args.m, md.n_elem, md.n_node, md.n_dof, N_FREQ)
    return md, mat, freqs, name


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed regions (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [v.strip() for v in line.split(",")]))

    def summary(self, windows):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons, pw = [], 0.0, set(), 0.0
        for t, r in self.rows:
            if len(r) < 7 or not any(a <= t <= b for a, b in windows):
                continue
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1])); pw = max(pw, float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "power_w_max": pw or None,
                "samples": len(sm)}


def cpu_arm(md, mat, omega, budget_points=6e7, lu_n=6144, repeat=1):
    """The reference algorithm on the host cores: the oracle (exact CPU restatement, OpenMP over integration elements +
    critical scatter like src/build_lse_mechanics_bem_harela.f90:231-237,1294-1319) for the assembly and OpenBLAS
    zgetrf/zgetrs (scipy-bundled) for the LU, on a BOUNDED SAMPLE of the step, scaled to one full step ("extrapolated": true;
    the scale factors are in "scale").  cpu_full_step below is the same thing measured in full."""
    from oracle import oracle as orc
    from scipy.linalg import lapack
    ncores = _host_cores()
    nblas = blas_threads()
    o = orc.Oracle(md)
    n = md.n_dof
    # sample: all elements x every `stride`-th collocation point (the sample costs about `budget_points` quadrature points)
    est_points = md.n_elem * md.n_colloc * 5.0
    stride = max(1, int(round(est_points / budget_points)))
    times = []
    for _ in range(repeat):
        t0 = time.time(); _, _, ns, pts = o.assemble_colloc_sample(omega, mat, stride // 2, stride, nthreads=ncores); t1 = time.time()
        times.append(t1 - t0)
    t_asm_sample = min(times)
    t_asm = t_asm_sample * md.n_colloc / ns
    lu_n = min(lu_n, n)
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.standard_normal((lu_n, lu_n)) + 1j * rng.standard_normal((lu_n, lu_n)))
    A[np.arange(lu_n), np.arange(lu_n)] += 0.5 * lu_n ** 0.5     # BEM-like: a strong diagonal (the free term), few row interchanges
    b = rng.standard_normal(lu_n) + 1j * rng.standard_normal(lu_n)
    t0 = time.time(); lu, piv, info = lapack.zgetrf(A, overwrite_a=True); x, info = lapack.zgetrs(lu, piv, b); t1 = time.time()
    t_lu = (t1 - t0) * (n / lu_n) ** 3
    sample = ("assembly: oracle on all %d elements x every %d-th collocation point (%d of %d points, %.1f s) scaled by %d/%d; "
              "LU: OpenBLAS zgetrf+zgetrs at n=%d (%.1f s, %s BLAS threads) scaled by (%d/%d)^3" % (
                  md.n_elem, stride, ns, md.n_colloc, t_asm_sample, md.n_colloc, ns, lu_n, t1 - t0, nblas, n, lu_n))
    return {"value": 1.0 / (t_asm + t_lu), "unit": "solves/s", "cores": ncores, "blas_threads": nblas, "kind": "port", "sample": sample, "extrapolated": True,
            "scale": {"assembly": md.n_colloc / ns, "lu": (n / lu_n) ** 3}, "sample_wall_s": t_asm_sample + (t1 - t0),
            "assembly_s_per_step": t_asm, "lu_s_per_step": t_lu, "assembly_gentries_per_s": n * n / t_asm / 1e9}


def cpu_full_step(md, mat, omega):
    """ONE FULL step of the reference algorithm on the host cores, nothing sampled or scaled: the oracle assembles the whole n_dof x n_dof matrix
    (all elements x all collocation points, OpenMP over integration elements), then OpenBLAS zgetrf + zgetrs factorise and solve THAT matrix.
    Needs 16 n^2 bytes of host memory (14.6 GB at 30258 DOF); returns None if the host cannot hold it."""
    from oracle import oracle as orc
    from scipy.linalg import lapack
    ncores = _host_cores()
    nblas = blas_threads()
    n = md.n_dof
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except Exception:
        avail = None
    if avail is not None and avail < 1.3 * 16.0 * n * n:
        return None
    o = orc.Oracle(md)
    try:
        t0 = time.time(); A, b, stats = o.assemble(omega, mat, nthreads=ncores); t1 = time.time()
        lu, piv, info = lapack.zgetrf(A, overwrite_a=True)
        x, info2 = lapack.zgetrs(lu, piv, b); t2 = time.time()
    except MemoryError:
        return None
    if info != 0 or info2 != 0 or not np.isfinite(x).all():
        raise RuntimeError("reference arm: zgetrf/zgetrs failed (info %d, %d)" % (info, info2))
    del A, lu
    return {"value": 1.0 / (t2 - t0), "unit": "solves/s", "cores": ncores, "blas_threads": nblas, "kind": "port", "extrapolated": False,
            "sample": "ONE FULL step measured, nothing scaled: oracle assembly of all %d elements x all %d collocation points (%.1f s, %d OpenMP threads) + OpenBLAS "
                      "zgetrf+zgetrs of that %d x %d matrix (%.1f s, %s BLAS threads)" % (md.n_elem, md.n_colloc, t1 - t0, ncores, n, n, t2 - t1, nblas),
            "assembly_s_per_step": t1 - t0, "lu_s_per_step": t2 - t1, "assembly_gentries_per_s": n * n / (t1 - t0) / 1e9, "x_checksum": float(np.abs(x).sum())}


# ---------------------------------------------------------------------------------------------------------------------
# --workload static: BASELINE config 2 (ME-ST-EL-002 refined: static elastic cube, quad9, ~10k DOF; staela assembly + real LU)
# ---------------------------------------------------------------------------------------------------------------------
STATIC_METRIC = "static 3D BEM end-to-end solves/s (assemble + dgetrf + dgetrs) on the ME-ST-EL-002 cube refined to quad9, 9522 DOF"


def static_workload(args):
    from multifebe_b200.host import Model, Material, cube_mesh, cube_bcs, shape
    et = {"tri3": shape.TRI3, "tri6": shape.TRI6, "quad4": shape.QUAD4, "quad8": shape.QUAD8, "quad9": shape.QUAD9}[args.static_etype]
    md = Model(cube_mesh(args.static_m, et), cube_bcs())
    mat = Material(1.0, 1.0, 0.25, 0.0)
    name = "static S-cube %s m=%d (ME-ST-EL-002 BCs): %d elements, %d nodes, %d DOF" % (args.static_etype, args.static_m, md.n_elem, md.n_node, md.n_dof)
    return md, mat, name


def cpu_arm_static(md, mat, budget_points=6e7, lu_n=6144):
    """Reference algorithm of the static path on the host cores: oracle (Kelvin kernels, OpenMP over integration elements +
    critical scatter) on a bounded sample of the collocation points + OpenBLAS dgetrf/dgetrs, scaled to one full solve."""
    from oracle import oracle as orc
    from scipy.linalg import lapack
    ncores = _host_cores(); blas_threads()
    o = orc.Oracle(md)
    n = md.n_dof
    stride = max(1, int(round(md.n_elem * md.n_colloc * 30.0 / budget_points)))
    t0 = time.time(); _, _, ns, pts = o.assemble_colloc_sample_static(mat, stride // 2, stride, nthreads=ncores); t1 = time.time()
    t_asm = (t1 - t0) * md.n_colloc / ns
    lu_n = min(lu_n, n)
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.standard_normal((lu_n, lu_n))); b = rng.standard_normal(lu_n)
    t0 = time.time(); lu, piv, info = lapack.dgetrf(A, overwrite_a=True); x, info = lapack.dgetrs(lu, piv, b); t2 = time.time()
    t_lu = (t2 - t0) * (n / lu_n) ** 3
    sample = ("assembly: static oracle on all %d elements x every %d-th collocation point (%d of %d points, %.1f s) scaled by %d/%d; "
              "LU: OpenBLAS dgetrf+dgetrs at n=%d (%.2f s) scaled by (%d/%d)^3" % (md.n_elem, stride, ns, md.n_colloc, t_asm * ns / md.n_colloc,
                                                                                   md.n_colloc, ns, lu_n, t2 - t0, n, lu_n))
    return {"value": 1.0 / (t_asm + t_lu), "unit": "solves/s", "cores": ncores, "kind": "port", "sample": sample,
            "assembly_s_per_step": t_asm, "lu_s_per_step": t_lu}


def run_static(args):
    """N ranks = N independent replicas of the same static solve (the static analysis has a single system: nothing to shard)."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    md, mat, name = static_workload(args)
    if args.impl == "reference":
        if rank != 0:
            return
        cpu_arm_static(md, mat, budget_points=2e6, lu_n=1024)
        t_ref0 = time.time(); res = [cpu_arm_static(md, mat) for _ in range(args.steps)]; wall_ref = time.time() - t_ref0
        v = float(np.mean([r["value"] for r in res])); cb = dict(res[-1]); cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": STATIC_METRIC, "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": wall_ref * 1e3 / max(args.steps, 1), "ms_per_full_step": 1e3 / v, "extrapolated": True, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": name, "note": "reference algorithm on host cores (oracle port; no Fortran compiler here), bounded sample scaled to a full solve"},
                          "cpu_baseline": cb, "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    from multifebe_b200 import capi
    ctx = capi.Context(local)
    t0 = time.time(); pr = capi.Problem(ctx, md); t_setup = time.time() - t0
    n = md.n_dof
    dev = torch.device("cuda", local)

    def barrier():
        ctx.mark(7); ctx.elapsed_ms(7, 7)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_max(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        x = pr.solve_static(mat)
    # device-resident arm: the library's own events around assemble + LU + solve (prescribed values resident, solution stays on the device
    # until the single download the call ends with; its 76 KB are inside the e2e arm below)
    acc = {}
    barrier(); w0 = time.time()
    for s in range(args.steps):
        x = pr.solve_static(mat)
        st = pr.stats()
        for k in ("MS_ASSEMBLE", "MS_REGULAR", "MS_ADAPTIVE", "MS_SINGULAR", "MS_ZERO", "MS_LU", "MS_SOLVE", "MS_GEMM", "MS_PANEL", "LAUNCHES", "LU_LAUNCHES", "GEMM_LAUNCHES", "GEMM_FLOPS"):
            acc[k] = acc.get(k, 0.0) + st[k]
    barrier(); windows = [(w0, time.time())]
    K = args.steps
    ms_dev = reduce_max((acc["MS_ASSEMBLE"] + acc["MS_LU"] + acc["MS_SOLVE"]) / K)
    # e2e arm: wall clock of the C-ABI call with host buffers (cvalue up, x down), max over ranks
    barrier(); t0 = time.time(); w0 = t0
    for s in range(args.steps):
        x = pr.solve_static(mat)
    barrier(); ms_e2e = reduce_max((time.time() - t0) * 1e3 / K); windows.append((w0, time.time()))
    peaks = ctx.measure_peaks() if rank == 0 else None
    if rank == 0:
        u, t = md.nodal_solution(x)
        lam2mu = 2.0 * mat.mu_r * mat.nu_r / (1.0 - 2.0 * mat.nu_r) + 2.0 * mat.mu_r
        err = float(np.abs(u[:, 0].real - md.node_x[:, 0] / lam2mu).max() * lam2mu)
        gemm_tf = acc["GEMM_FLOPS"] / max(acc["MS_GEMM"], 1e-9) / 1e9
        out = {"metric": STATIC_METRIC, "value": world * 1e3 / ms_dev, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": name, "sharding": "replicas only: the static analysis is one system, every rank solves the same problem",
                          "l2": "L2 flushed between steps by the assembly itself (it rewrites the %.2f GB matrix, larger than the 126 MB L2)" % (8.0 * n * n / 1e9),
                          "setup_s_once_per_mesh": t_setup},
               "clocks": clocks.summary(windows),
               "e2e": {"value": world * 1e3 / ms_e2e, "unit": "solves/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(md.cvalue.size * 8 + 1024), "d2h_bytes_per_step": int(8 * n + 4 * n),
                       "api": "mfb_staela3d_solve (host cvalue in, host x out)"},
               "gpu_launches": int(acc["LAUNCHES"] + acc["LU_LAUNCHES"]),
               "roofline": {"kernel": "k_dgemm_minus (real LU trailing update, DMMA.8x8x4)", "bound": "tensor", "achieved": gemm_tf, "peak": peaks["dmma_tflops"], "unit": "TFLOP/s",
                            "frac": gemm_tf / peaks["dmma_tflops"], "traffic": None, "avg_launch_ms": acc["MS_GEMM"] / max(acc["GEMM_LAUNCHES"], 1),
                            "launches_per_step": acc["GEMM_LAUNCHES"] / K, "share_of_step": (acc["MS_GEMM"] / K) / ms_dev,
                            "note": "2mnk flops per launch / CUDA-event time of the trailing updates; at this size the factorisation is bound by the panel "
                                    "(one grid barrier per column), not by the GEMM: see lu.ms_panel_on_lookahead_stream",
                            "peak_source": "FP64 tensor (DMMA) micro-benchmark measured live (mfb_measure_peaks)"},
               "assembly": {"ms": acc["MS_ASSEMBLE"] / K, "ms_regular": acc["MS_REGULAR"] / K, "ms_adaptive": acc["MS_ADAPTIVE"] / K, "ms_singular": acc["MS_SINGULAR"] / K,
                            "gentries_per_s": n * n / (acc["MS_ASSEMBLE"] / K * 1e-3) / 1e9, "matrix_write_gbs": 8.0 * n * n / (acc["MS_ASSEMBLE"] / K * 1e-3) / 1e9},
               "lu": {"ms": acc["MS_LU"] / K, "tflops": 2.0 / 3.0 * n ** 3 / (acc["MS_LU"] / K) / 1e9, "ms_gemm": acc["MS_GEMM"] / K, "ms_panel_on_lookahead_stream": acc["MS_PANEL"] / K,
                      "ms_dgetrs": acc["MS_SOLVE"] / K},
               "exact_solution_max_error": err, "peaks_measured_live": peaks}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_arm_static(md, mat)
        print(json.dumps(out), flush=True)
    pr.close(); ctx.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


ACOUSTIC_METRIC = "acoustic 3D BEM end-to-end solves/s (assemble + zgetrf + zgetrs of one frequency) on the ME-TH-AC-001 room refined to ~10k DOF"


def acoustic_workload(args):
    from multifebe_b200.host import FluidModel, Fluid, cube_mesh, room_bcs, shape
    et = {"tri3": shape.TRI3, "tri6": shape.TRI6, "quad4": shape.QUAD4, "quad8": shape.QUAD8, "quad9": shape.QUAD9}[args.acoustic_etype]
    md = FluidModel(cube_mesh(args.acoustic_m, et, L=3.0), room_bcs(1.0))
    fl = Fluid(rho=1.25, c=343.0)
    omega = 2.0 * np.pi * 40.0
    name = "acoustic room L=3 m (ME-TH-AC-001 BCs) %s m=%d at 40 Hz: %d elements, %d nodes, %d DOF" % (args.acoustic_etype, args.acoustic_m, md.n_elem, md.n_node, md.n_dof)
    return md, fl, name, omega


def cpu_arm_acoustic(args, md, fl, omega, lu_n=4096, budget_pairs=4e6):
    """Reference algorithm of the acoustic path on the host cores: the oracle (OpenMP over integration elements + critical scatter) on every
    stride-th collocation point of the same mesh (same mix of regular, quasi-singular and singular pairs), scaled to all of them, + OpenBLAS
    zgetrf/zgetrs scaled by n^3."""
    import copy
    from oracle import oracle as orc
    from scipy.linalg import lapack
    ncores = _host_cores(); blas_threads()
    stride = max(1, int(round(md.n_elem * md.n_colloc / budget_pairs)))
    sub = copy.copy(md)
    sel = np.arange(stride // 2, md.n_colloc, stride)
    for name in ("colloc_x", "colloc_node", "colloc_elem", "colloc_kn", "colloc_xi"):
        setattr(sub, name, np.ascontiguousarray(getattr(md, name)[sel]))
    sub.n_colloc = len(sel)
    o = orc.PotOracle(sub)
    t0 = time.time(); o.assemble(omega, fl, nthreads=ncores); t_sample = time.time() - t0
    t_asm = t_sample * md.n_colloc / len(sel)
    n = md.n_dof
    lu_n = min(lu_n, n)
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.standard_normal((lu_n, lu_n)) + 1j * rng.standard_normal((lu_n, lu_n))); b = rng.standard_normal(lu_n) + 0j
    t0 = time.time(); lu, piv, info = lapack.zgetrf(A, overwrite_a=True); x, info = lapack.zgetrs(lu, piv, b); t2 = time.time()
    t_lu = (t2 - t0) * (n / lu_n) ** 3
    sample = ("assembly: acoustic oracle on all %d elements x every %d-th collocation point (%d of %d points, %.1f s) scaled by %d/%d; LU: OpenBLAS "
              "zgetrf+zgetrs at n=%d (%.2f s) scaled by (%d/%d)^3" % (md.n_elem, stride, len(sel), md.n_colloc, t_sample, md.n_colloc, len(sel), lu_n, t2 - t0, n, lu_n))
    return {"value": 1.0 / (t_asm + t_lu), "unit": "solves/s", "cores": ncores, "kind": "port", "sample": sample, "assembly_s_per_step": t_asm, "lu_s_per_step": t_lu}


def run_acoustic(args):
    """One frequency of the acoustic room (ME-TH-AC-001 refined) per step; N ranks = N replicas solving the same frequency (a sweep would shard its
    frequencies exactly as the headline workload does; this line measures the per-frequency cost of the scalar path)."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    md, mat, name, omega = acoustic_workload(args)
    if args.impl == "reference":
        if rank != 0:
            return
        t_ref0 = time.time(); res = [cpu_arm_acoustic(args, md, mat, omega) for _ in range(args.steps)]; wall_ref = time.time() - t_ref0
        v = float(np.mean([r["value"] for r in res])); cb = dict(res[-1]); cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": ACOUSTIC_METRIC, "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": wall_ref * 1e3 / max(args.steps, 1), "ms_per_full_step": 1e3 / v, "extrapolated": True, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
                          "config": {"workload": name, "note": "reference algorithm on host cores (oracle port; no Fortran compiler here), bounded sample scaled to a full solve"},
                          "cpu_baseline": cb, "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    from multifebe_b200 import capi
    ctx = capi.Context(local)
    t0 = time.time(); pr = capi.Problem(ctx, md); t_setup = time.time() - t0
    n = md.n_dof
    dev = torch.device("cuda", local)

    def barrier():
        ctx.mark(7); ctx.elapsed_ms(7, 7)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_max(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        x = pr.solve_frequency_fluid(omega, mat)
    # device-resident arm: the library's own events around assemble + LU + solve (prescribed values resident, solution stays on the device
    # until the single download the call ends with; its 76 KB are inside the e2e arm below)
    acc = {}
    barrier(); w0 = time.time()
    for s in range(args.steps):
        x = pr.solve_frequency_fluid(omega, mat)
        st = pr.stats()
        for k in ("MS_ASSEMBLE", "MS_REGULAR", "MS_ADAPTIVE", "MS_SINGULAR", "MS_ZERO", "MS_LU", "MS_SOLVE", "MS_GEMM", "MS_PANEL", "LAUNCHES", "LU_LAUNCHES", "GEMM_LAUNCHES", "GEMM_FLOPS"):
            acc[k] = acc.get(k, 0.0) + st[k]
    barrier(); windows = [(w0, time.time())]
    K = args.steps
    ms_dev = reduce_max((acc["MS_ASSEMBLE"] + acc["MS_LU"] + acc["MS_SOLVE"]) / K)
    # e2e arm: wall clock of the C-ABI call with host buffers (cvalue up, x down), max over ranks
    barrier(); t0 = time.time(); w0 = t0
    for s in range(args.steps):
        x = pr.solve_frequency_fluid(omega, mat)
    barrier(); ms_e2e = reduce_max((time.time() - t0) * 1e3 / K); windows.append((w0, time.time()))
    peaks = ctx.measure_peaks() if rank == 0 else None
    if rank == 0:
        from multifebe_b200.host import room_analytic
        pn, un = md.nodal_solution(x)
        p_ex, _ = room_analytic(md.node_x[:, 0], omega, mat, L=3.0, P=1.0)
        err = float(np.abs(pn - p_ex).max() / np.abs(p_ex).max())
        gemm_tf = acc["GEMM_FLOPS"] / max(acc["MS_GEMM"], 1e-9) / 1e9
        out = {"metric": ACOUSTIC_METRIC, "value": world * 1e3 / ms_dev, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
               "config": {"workload": name, "sharding": "replicas: every rank solves the same frequency (a sweep shards its frequencies like the headline workload)",
                          "l2": "L2 flushed between steps by the assembly itself (it rewrites the %.2f GB matrix, larger than the 126 MB L2)" % (16.0 * n * n / 1e9),
                          "setup_s_once_per_mesh": t_setup},
               "clocks": clocks.summary(windows),
               "e2e": {"value": world * 1e3 / ms_e2e, "unit": "solves/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(md.cvalue.size * 16 + 1024), "d2h_bytes_per_step": int(16 * n + 4 * n),
                       "api": "mfb_harpot3d_solve_frequency (host cvalue in, host x out)"},
               "gpu_launches": int(acc["LAUNCHES"] + acc["LU_LAUNCHES"]),
               "roofline": {"kernel": "k_zgemm3m_tma (LU trailing update, 3M complex product on DMMA.8x8x4)", "bound": "tensor", "achieved": gemm_tf, "peak": peaks["dmma_tflops"], "unit": "TFLOP/s",
                            "frac": gemm_tf / peaks["dmma_tflops"], "traffic": None, "avg_launch_ms": acc["MS_GEMM"] / max(acc["GEMM_LAUNCHES"], 1),
                            "launches_per_step": acc["GEMM_LAUNCHES"] / K, "share_of_step": (acc["MS_GEMM"] / K) / ms_dev,
                            "note": "executed tensor-pipe flops (6mnk per complex product) / CUDA-event time of the trailing updates; at this size the factorisation is bound "
                                    "by the panel, not by the GEMM: see lu.ms_panel_on_lookahead_stream",
                            "peak_source": "FP64 tensor (DMMA) micro-benchmark measured live (mfb_measure_peaks)"},
               "assembly": {"ms": acc["MS_ASSEMBLE"] / K, "ms_regular": acc["MS_REGULAR"] / K, "ms_adaptive": acc["MS_ADAPTIVE"] / K, "ms_singular": acc["MS_SINGULAR"] / K,
                            "gentries_per_s": n * n / (acc["MS_ASSEMBLE"] / K * 1e-3) / 1e9, "matrix_write_gbs": 16.0 * n * n / (acc["MS_ASSEMBLE"] / K * 1e-3) / 1e9},
               "lu": {"ms": acc["MS_LU"] / K, "tflops": 8.0 / 3.0 * n ** 3 / (acc["MS_LU"] / K) / 1e9, "ms_gemm": acc["MS_GEMM"] / K, "ms_panel_on_lookahead_stream": acc["MS_PANEL"] / K,
                      "ms_zgetrs": acc["MS_SOLVE"] / K},
               "analytic_solution_rel_error": err, "peaks_measured_live": peaks}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_arm_acoustic(args, md, mat, omega)
        print(json.dumps(out), flush=True)
    pr.close(); ctx.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# --workload c1: BASELINE config 1 = the reference's own tutorial input docs/examples/ME-TH-EL-001 (t3.dat + t3.msh: 462 nodes, 744 tri3,
# 1386 DOF, 300 frequencies lin in [0.01, 15] rad/s), committed as a test vector under tests/golden/ME-TH-EL-001/.  One step = the WHOLE sweep.
# ---------------------------------------------------------------------------------------------------------------------
C1_METRIC = "ME-TH-EL-001 (reference input t3.dat: 1386 DOF, 300-frequency sweep) end-to-end solves/s (assemble + zgetrf + zgetrs per frequency)"
C1_CASE = os.path.join(ROOT, "tests", "golden", "ME-TH-EL-001", "t3.dat")


def cpu_arm_c1(case, md, every=10):
    """The reference algorithm on the host cores for every `every`-th frequency of the sweep, in full (oracle assembly of the whole 1386 x 1386 system +
    OpenBLAS zgetrf/zgetrs of it), scaled to the 300 frequencies by the count only."""
    from oracle import oracle as orc
    ncores = _host_cores(); nblas = blas_threads()
    o = orc.Oracle(md)
    ks = list(range(every // 2, len(case.omega), every))
    t0 = time.time()
    for kf in ks:
        A, b, _ = o.assemble(float(case.omega[kf]), case.material, nthreads=ncores)
        x, _, _ = orc.lu_solve(A, b)
    t = time.time() - t0
    return {"value": len(ks) / t, "unit": "solves/s", "cores": ncores, "blas_threads": nblas, "kind": "port", "extrapolated": False,
            "sample": "every %d-th frequency of the sweep (%d of %d), each one assembled and solved in full by the oracle + OpenBLAS zgetrf/zgetrs: %.1f s" % (every, len(ks), len(case.omega), t),
            "sample_wall_s": t}


def run_c1(args):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    from multifebe_b200.host.casefile import CaseFile
    from multifebe_b200.host import column_analytic_u
    case = CaseFile(C1_CASE); md = case.build_model(); mat = case.material
    nf = len(case.omega); n = md.n_dof
    name = "ME-TH-EL-001 t3.dat: %d elements, %d nodes, %d DOF, %d-frequency sweep (one step = the whole sweep)" % (md.n_elem, md.n_node, n, nf)
    config = {"workload": name, "sharding": "frequencies round-robin over ranks, mesh+plan replicated, no data-path collective",
              "l2": "every frequency rewrites the 30.7 MB system matrix; 300 different frequencies per step, nothing is reused between them"}
    if args.impl == "reference":
        if rank != 0:
            return
        t0 = time.time(); res = [cpu_arm_c1(case, md) for _ in range(args.steps)]; wall = time.time() - t0
        v = float(np.mean([r["value"] for r in res])); cb = dict(res[-1]); cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": C1_METRIC, "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": wall * 1e3 / max(args.steps, 1), "ms_per_full_step": nf * 1e3 / v, "extrapolated": True,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "reference input (tests/golden/ME-TH-EL-001)",
                          "config": config, "cpu_baseline": cb, "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    from multifebe_b200 import capi
    from multifebe_b200.sweep import owned_frequencies
    ctx = capi.Context(local); dev = torch.device("cuda", local)
    t0 = time.time(); pr = capi.Problem(ctx, md); t_setup = time.time() - t0
    mine = owned_frequencies(nf, rank, world)

    def barrier():
        ctx.mark(7); ctx.elapsed_ms(7, 7)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_max(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    clocks = ClockSampler(local) if rank == 0 else None
    for w in range(max(args.warmup, 3)):
        pr.solve_frequency(float(case.omega[mine[w % len(mine)]]), mat, host=(w == 0))
    keys = ("MS_ASSEMBLE", "MS_REGULAR", "MS_ADAPTIVE", "MS_SINGULAR", "MS_LU", "MS_SOLVE", "MS_GEMM", "MS_PANEL", "LAUNCHES", "LU_LAUNCHES", "GEMM_LAUNCHES", "GEMM_FLOPS", "GEMM_EXEC_FLOPS")
    acc = {k: 0.0 for k in keys}
    # arm 1, device resident: prescribed values already on the device, solutions stay there
    barrier(); w0 = time.time(); ctx.mark(0)
    for s in range(args.steps):
        for kf in mine:
            pr.solve_frequency(float(case.omega[kf]), mat, host=False)
            st = pr.stats()
            for k in keys:
                acc[k] += st[k]
    ctx.mark(1); ms_dev = ctx.elapsed_ms(0, 1)
    barrier(); windows = [(w0, time.time())]
    ms_dev = reduce_max(ms_dev)
    # arm 2, end to end through the C ABI with host buffers: cvalue up and x down for every frequency
    X = np.zeros((nf, n), dtype=np.complex128)
    barrier(); w0 = time.time()
    for s in range(args.steps):
        for kf in mine:
            X[kf] = pr.solve_frequency(float(case.omega[kf]), mat, host=True)
    barrier(); ms_e2e = reduce_max((time.time() - w0) * 1e3); windows.append((w0, time.time()))
    # arm 3, lanes: args.lanes (context, problem) pairs on this GPU, one host thread each, the rank's frequencies dealt round-robin to the lanes; host buffers
    # in and out for every frequency (this is the end-to-end number of the workload: the sweep as a user of the C ABI would run it on one GPU)
    lanes_out = None
    if args.lanes > 1:
        pl = capi.ProblemLanes(md, local, n_lanes=args.lanes)
        om_mine = [float(case.omega[kf]) for kf in mine]
        pl.run(om_mine[:2 * args.lanes], mat)                       # warm-up of every lane
        barrier(); w0 = time.time()
        for s in range(args.steps):
            XL = pl.run(om_mine, mat)
        barrier(); ms_lanes = reduce_max((time.time() - w0) * 1e3); windows.append((w0, time.time()))
        lanes_out = {"lanes": args.lanes, "ms_per_step": ms_lanes / args.steps, "solves_per_s": world * len(mine) * args.steps / (ms_lanes * 1e-3),
                     "max_rel_diff_vs_one_at_a_time": float(np.abs(XL - X[mine]).max() / np.abs(X[mine]).max())}
        pl.close()
    peaks = ctx.measure_peaks() if rank == 0 else None
    if rank == 0:
        K = args.steps
        # the analytic column of doc_src/ME-TH-EL-001.tex:32-56 at the lowest frequencies of this rank (mesh error ~1e-3)
        errs = []
        for kf in mine[:3]:
            u, t = md.nodal_solution(X[kf]); ua = column_analytic_u(md.node_x[:, 0], float(case.omega[kf]), mat)
            errs.append(float(np.abs(u[:, 0] - ua).max() / np.abs(ua).max()))
        nfr = len(mine) * K
        lu_tf = 8.0 / 3.0 * n ** 3 / (acc["MS_LU"] / nfr) / 1e9
        # the factorisation of a system this small runs as one CUDA graph: its trailing updates are not timed one by one (MS_GEMM = 0); the LU as a whole is
        gemm_timed = acc["MS_GEMM"] > 1e-3
        gemm_tf = acc["GEMM_EXEC_FLOPS"] / acc["MS_GEMM"] / 1e9 if gemm_timed else lu_tf
        out = {"metric": C1_METRIC, "value": world * nfr / (ms_dev * 1e-3), "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / K,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "reference input (tests/golden/ME-TH-EL-001)",
               "config": config, "setup_s_once_per_mesh": t_setup, "clocks": clocks.summary(windows),
               "e2e": {"value": world * nfr / (ms_e2e * 1e-3), "unit": "solves/s", "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": int(len(mine) * md.cvalue.size * 16), "d2h_bytes_per_step": int(len(mine) * (16 * n + 4 * n)),
                       "api": "mfb_harela3d_solve_frequency (host cvalue in, host x out), once per frequency"},
               "gpu_launches": int(acc["LAUNCHES"] + acc["LU_LAUNCHES"]),
               "per_frequency_ms": {k[3:].lower(): acc[k] / nfr for k in keys if k.startswith("MS_")},
               "roofline": {"kernel": "k_zgemm3m_tma (LU trailing update)" if gemm_timed else "zgetrf + zgetrs of one frequency as ONE CUDA graph (panel columns, interchanges, triangular solves, trailing updates)",
                            "bound": "tensor", "achieved": gemm_tf, "peak": peaks["dmma_tflops"], "unit": "TFLOP/s", "frac": gemm_tf / peaks["dmma_tflops"],
                            "traffic": None, "share_of_step": (acc["MS_GEMM"] if gemm_timed else acc["MS_LU"]) / ms_dev,
                            "note": "at 1386 DOF the step is latency-bound (panel column chain, small grids), not pipe-bound: LU %.2f TFLOP/s of 8/3 n^3" % lu_tf,
                            "peak_source": "FP64 tensor (DMMA) micro-benchmark measured live (mfb_measure_peaks)"},
               "analytic_column_rel_error_first_frequencies": errs, "peaks_measured_live": peaks}
        if lanes_out is not None:      # the sweep with several frequencies in flight IS the end-to-end number; the one-at-a-time figures stay beside it
            out["one_at_a_time"] = {"device_resident_solves_per_s": out["value"], "e2e_solves_per_s": out["e2e"]["value"], "ms_per_step_e2e": out["e2e"]["ms_per_step"]}
            out["e2e"].update({"value": lanes_out["solves_per_s"], "ms_per_step": lanes_out["ms_per_step"],
                               "api": "mfb_harela3d_solve_frequency (host cvalue in, host x out) on %d problems of the same mesh in flight (capi.ProblemLanes)" % args.lanes})
            out["lanes"] = lanes_out
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_arm_c1(case, md)
        print(json.dumps(out), flush=True)
    pr.close(); ctx.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()



COUPLED_METRIC = ("poroelastic (harpor) + acoustic (harpot) coupled 3D BEM end-to-end solves/s (both regions assembled, be-be interface combination, zgetrf + zgetrs "
                  "of one frequency) at ~20k DOF")


def coupled_workload(args):
    """BASELINE config 4: a water column (inviscid fluid) over a water-saturated poroelastic layer, two boxes of a 10 m cube that share one
    perfectly bonded, permeable be-be interface; pressure prescribed on the far fluid face, sliding impermeable lateral walls, rigid
    impermeable bottom."""
    from multifebe_b200.host import MultiRegionModel, Region, FLUID, two_box_mesh, shape, Fluid, Poro
    from multifebe_b200.host.multiregion import PORO
    et = {"tri3": shape.TRI3, "tri6": shape.TRI6, "quad4": shape.QUAD4, "quad8": shape.QUAD8, "quad9": shape.QUAD9}[args.coupled_etype]
    fl = Fluid(1000.0, 1500.0, 0.0)
    po = Poro(rhof=1000.0, rhos=2650.0, lam=1.0e8, mu=1.0e8, xi=0.03, phi=0.35, rhoa=150.0, R=4.0e8, Q=7.0e8, b=1.0e6)
    bcs = {1: (0, 1.0), 2: ([1, 0, 0, 0], [0, 0, 0, 0])}
    for q in (3, 4, 5, 6):
        bcs[q] = (1, 0.0)
    for q in (13, 14, 15, 16):
        ct = [1, 1, 1, 1]; ct[2 if q in (13, 14) else 3] = 0
        bcs[q] = (ct, [0, 0, 0, 0])
    parts = (1, 2, 3, 4, 5, 6, 7, 13, 14, 15, 16)
    mrm = MultiRegionModel(two_box_mesh(args.coupled_m, et, L=10.0), [Region(FLUID, fl, [1, 3, 4, 5, 6, 7]), Region(PORO, po, [-7, 2, 13, 14, 15, 16])],
                           {b: b for b in parts}, bcs, interface_ctype={7: 0})
    omega = 2.0 * np.pi * 50.0
    name = "water box | saturated poroelastic box, 10 m cube, %s m=%d per face at 50 Hz: %d DOF" % (args.coupled_etype, args.coupled_m, mrm.n_dof)
    return mrm, name, omega


def cpu_arm_coupled(args, mrm, omega, lu_n=4096, budget_pairs=4e6):
    """Reference algorithm on the host cores: each region's oracle (one integration pass yields both H and G, as in the reference) on every
    stride-th collocation point of the region's mesh, scaled to all of them, + OpenBLAS zgetrf/zgetrs scaled by n^3."""
    import copy
    from oracle import oracle as orc
    from scipy.linalg import lapack
    from multifebe_b200.host.coupled import local_models
    ncores = _host_cores(); blas_threads()
    t_asm = 0.0; notes = []
    for kr, region in enumerate(mrm.regions):
        mH, _, _ = local_models(mrm, kr)
        stride = max(1, int(round(mH.n_elem * mH.n_colloc / budget_pairs)))
        sub = copy.copy(mH)
        sel = np.arange(stride // 2, mH.n_colloc, stride)
        for name in ("colloc_x", "colloc_node", "colloc_elem", "colloc_kn", "colloc_xi"):
            setattr(sub, name, np.ascontiguousarray(getattr(mH, name)[sel]))
        sub.n_colloc = len(sel)
        o = (orc.PotOracle if region.kind == "fluid" else orc.PorOracle if region.kind == "poro" else orc.Oracle)(sub)
        t0 = time.time(); o.assemble(omega, region.material, nthreads=ncores); t_s = time.time() - t0
        t_asm += t_s * mH.n_colloc / len(sel)
        notes.append("%s region: %d elements x %d of %d collocation points in %.1f s" % (region.kind, mH.n_elem, len(sel), mH.n_colloc, t_s))
    n = mrm.n_dof
    lu_n = min(lu_n, n)
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.standard_normal((lu_n, lu_n)) + 1j * rng.standard_normal((lu_n, lu_n))); b = rng.standard_normal(lu_n) + 0j
    t0 = time.time(); lu, piv, info = lapack.zgetrf(A, overwrite_a=True); x, info = lapack.zgetrs(lu, piv, b); t2 = time.time()
    t_lu = (t2 - t0) * (n / lu_n) ** 3
    sample = "assembly: " + "; ".join(notes) + ", each scaled to all its points; LU: OpenBLAS zgetrf+zgetrs at n=%d (%.2f s) scaled by (%d/%d)^3" % (lu_n, t2 - t0, n, lu_n)
    return {"value": 1.0 / (t_asm + t_lu), "unit": "solves/s", "cores": ncores, "kind": "port", "sample": sample, "assembly_s_per_step": t_asm, "lu_s_per_step": t_lu}


def run_coupled(args):
    """One frequency of the fluid | poroelastic model per step (capi.CoupledProblem.solve_frequency_resident: single-region assemblies, device
    combination, LU); N ranks = N replicas (first hardware run: profiles/r02_first_contact.log)."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    mrm, name, omega = coupled_workload(args)
    if args.impl == "reference":
        if rank != 0:
            return
        t_ref0 = time.time(); res = [cpu_arm_coupled(args, mrm, omega) for _ in range(args.steps)]; wall_ref = time.time() - t_ref0
        v = float(np.mean([r["value"] for r in res])); cb = dict(res[-1]); cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": COUPLED_METRIC, "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": wall_ref * 1e3 / max(args.steps, 1), "ms_per_full_step": 1e3 / v, "extrapolated": True, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
                          "config": {"workload": name, "note": "reference algorithm on host cores (oracle port; no Fortran compiler here), bounded sample scaled to a full solve"},
                          "cpu_baseline": cb, "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    from multifebe_b200 import capi
    ctx = capi.Context(local)
    dev = torch.device("cuda", local)
    t0 = time.time(); cp = capi.CoupledProblem(ctx, mrm); t_setup = time.time() - t0

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_max(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        x = cp.solve_frequency_resident(omega)
    acc = {}; by_problem = {}
    barrier(); t0 = time.time(); w0 = t0
    for s in range(args.steps):
        x = cp.solve_frequency_resident(omega)
        for key, pr in list(cp.problems.items()):
            st = pr.stats()
            for k in ("MS_ASSEMBLE", "MS_REGULAR", "MS_ADAPTIVE", "MS_SINGULAR", "LAUNCHES"):
                acc[k] = acc.get(k, 0.0) + st[k]
            by_problem["%s ndof=%d" % (str(key), pr.ndof)] = by_problem.get("%s ndof=%d" % (str(key), pr.ndof), 0.0) + st["MS_REGULAR"]
        st = cp._solver.stats()
        for k in ("MS_LU", "MS_SOLVE", "MS_GEMM", "MS_PANEL", "LU_LAUNCHES", "GEMM_LAUNCHES", "GEMM_FLOPS"):
            acc[k] = acc.get(k, 0.0) + st[k]
    barrier(); ms_e2e = reduce_max((time.time() - t0) * 1e3 / args.steps); windows = [(w0, time.time())]
    K = args.steps
    ms_dev = reduce_max((acc["MS_ASSEMBLE"] + acc["MS_LU"] + acc["MS_SOLVE"]) / K)
    peaks = ctx.measure_peaks() if rank == 0 else None
    if rank == 0:
        n = mrm.n_dof
        gemm_tf = acc["GEMM_FLOPS"] / max(acc["MS_GEMM"], 1e-9) / 1e9
        out = {"metric": COUPLED_METRIC, "value": world * 1e3 / ms_dev, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
               "config": {"workload": name, "sharding": "replicas: every rank solves the same frequency",
                          "l2": "every step rewrites the local and the coupled matrices (%.2f GB for the coupled one alone, larger than the 126 MB L2)" % (16.0 * n * n / 1e9),
                          "setup_s_once_per_mesh": t_setup,
                          "note": "value = library events of the four local assemblies + LU + solve (the combination kernels and the host-side term lists are "
                                  "outside those events and inside e2e)"},
               "clocks": clocks.summary(windows),
               "e2e": {"value": world * 1e3 / ms_e2e, "unit": "solves/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(sum(m.cvalue.size for m3 in cp.locals for m in m3[:2]) * 16),
                       "d2h_bytes_per_step": int(16 * n + 4 * n), "api": "mfb_harpot3d_assemble / mfb_harpor3d_assemble x2, mfb_combine_columns, mfb_add_entries, mfb_zsolve, mfb_get_solution"},
               "gpu_launches": int(acc["LAUNCHES"] + acc["LU_LAUNCHES"]),
               "roofline": {"kernel": "k_zgemm3m_tma (LU trailing update)", "bound": "tensor", "achieved": gemm_tf, "peak": peaks["dmma_tflops"], "unit": "TFLOP/s",
                            "frac": gemm_tf / peaks["dmma_tflops"], "traffic": None, "share_of_step": (acc["MS_GEMM"] / K) / ms_dev,
                            "peak_source": "FP64 tensor (DMMA) micro-benchmark measured live (mfb_measure_peaks)"},
               "assembly": {"ms": acc["MS_ASSEMBLE"] / K, "ms_regular": acc["MS_REGULAR"] / K, "ms_adaptive": acc["MS_ADAPTIVE"] / K, "ms_singular": acc["MS_SINGULAR"] / K,
                            "ms_regular_by_problem": {k: v / K for k, v in by_problem.items()}},
               "lu": {"ms": acc["MS_LU"] / K, "tflops": 8.0 / 3.0 * n ** 3 / (acc["MS_LU"] / K) / 1e9, "ms_gemm": acc["MS_GEMM"] / K, "ms_zgetrs": acc["MS_SOLVE"] / K},
               "peaks_measured_live": peaks}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_arm_coupled(args, mrm, omega)
        print(json.dumps(out), flush=True)
    cp.close(); ctx.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


def headline_config(name):
    """The `config` object of the headline workload: the SAME dict in both arms (the driver compares them)."""
    return {"workload": name, "sharding": "frequencies round-robin over ranks, mesh+plan replicated, no data-path collective",
            "frequencies_timed": "step s on rank q solves frequency ((s*N+q)*21) mod 64 of the sweep (spread over the whole band)",
            "l2": "inputs larger than L2 (system matrix 14.6 GB >> 126 MB L2), no explicit flush"}


def run_reference(args):
    """The reference algorithm on the host cores (oracle port: no Fortran compiler in the image).  Timed step 0 is ONE FULL
    step, measured, nothing scaled (cpu_full_step); `value` is that measurement.  The other steps are bounded samples scaled to a full step
    (cpu_arm), reported beside it as a cross-check of the scaling used by the `cpu_baseline` leg of the GPU arm.  `ms_per_step` is the real
    wall time of this run per step, so steps x ms_per_step is what the driver's clock saw."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    md, mat, freqs, name = workload(args)
    for w in range(min(args.warmup, 1)):
        cpu_arm(md, mat, freqs[0], budget_points=2e6, lu_n=1024)
    t0 = time.time()
    full = None
    if not args.reference_sampled_only:
        full = cpu_full_step(md, mat, freqs[0])
    res = [cpu_arm(md, mat, freqs[(s * args.gpus * 21) % N_FREQ]) for s in range(1 if full else 0, args.steps)]
    wall = time.time() - t0
    v_samples = float(np.mean([r["value"] for r in res])) if res else None
    cb = dict(full if full else res[-1])
    v = cb["value"] if full else v_samples
    cb["value"] = v
    if full and res:
        cb["sampled_steps"] = {"n": len(res), "extrapolated_solves_per_s_mean": v_samples, "ratio_to_full_measurement": v_samples / v, "last": res[-1]}
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": wall * 1e3 / max(args.steps, 1), "ms_per_full_step": 1e3 / v, "extrapolated": not bool(full),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
           "config": headline_config(name),
           "note": "reference algorithm on host cores (oracle port: the Fortran reference cannot be compiled here -- no Fortran compiler). value = 1 / (wall time of "
                   "ONE FULL step: whole-matrix assembly + zgetrf + zgetrs at full size), measured in timed step 0; the remaining steps are bounded samples scaled to a "
                   "full step (cpu_baseline.sampled_steps); ms_per_step = wall time of this run / steps" if full else
                   "reference algorithm on host cores (oracle port); the host could not hold the full matrix: every step is a bounded sample SCALED to a full step",
           "cpu_baseline": cb, "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def single_frequency_arm(pr, ctx, capi, dist, torch, dev, rank, world, omega, mat, steps, barrier, reduce_max):
    """One frequency assembled and solved by all ranks together (mfb_dist_solve_frequency); max-over-ranks wall time per call
    (the call synchronises its stream before returning), checked against the single-GPU solution of the same frequency."""
    try:
        x1 = pr.solve_frequency(omega, mat, host=True)
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(capi.dist_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(uid, 0)
        pr.dist_init(rank, world, uid.cpu().numpy().tobytes(), 0)
        x2 = pr.dist_solve_frequency(omega, mat)     # warm-up
        calls = []
        barrier(); t0 = time.time()
        for _ in range(steps):
            t1 = time.time(); x2 = pr.dist_solve_frequency(omega, mat); calls.append((time.time() - t1) * 1e3)
        barrier(); ms = reduce_max((time.time() - t0) * 1e3) / steps
        st = pr.stats()
        print("rank %d single-frequency calls (ms): %s; device total of the last %.1f" % (rank, ", ".join("%.1f" % c for c in calls), st["MS_DIST_TOTAL"]), file=sys.stderr, flush=True)
        err = float(np.abs(x2 - x1).max() / np.abs(x1).max())
        return {"ms_per_frequency": ms, "solves_per_s": 1e3 / ms, "relerr_vs_one_gpu_solution": err,
                "rank0_device_ms_total": st["MS_DIST_TOTAL"],
                "rank0_ms": {"assemble_own_row_blocks": st["MS_ASSEMBLE"], "redistribute_nccl": st["MS_REDIST"], "lu_distributed": st["MS_DIST_LU"],
                             "back_substitution": st["MS_DIST_SOLVE"]},
                "lu_tflops_all_ranks": 8.0 / 3.0 * pr.m.n_dof ** 3 / (st["MS_DIST_LU"] * 1e-3) / 1e12,
                "layout": "assembly by collocation-row blocks, LU on block-cyclic columns (nb=256), NCCL send/recv for the slabs, "
                          "NCCL broadcast per panel with one panel of look-ahead",
                "api": "mfb_dist_solve_frequency (host cvalue in, host x out on every rank)"}
    except Exception as e:   # reported, never fatal for the headline numbers
        return {"error": "%s: %s" % (type(e).__name__, e)}


def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    from multifebe_b200 import capi
    from multifebe_b200.sweep import FrequencySweep
    md, mat, freqs, name = workload(args)
    ctx = capi.Context(local)
    t0 = time.time(); pr = capi.Problem(ctx, md); t_setup = time.time() - t0
    n = md.n_dof
    dev = torch.device("cuda", local)

    def barrier():
        ctx_sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def ctx_sync():
        ctx.mark(7); ctx.elapsed_ms(7, 7)

    def reduce_max(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def kf_of(step):
        # steps stride through the 64-frequency sweep (21 is coprime with 64) so that a short run samples low, middle and high
        # frequencies: the cost of a Gauss point depends on |k r| (series or direct branch of E_m, lib/fbem/src/numerical.f90:1258-1330)
        return ((step * world + rank) * 21) % N_FREQ

    clocks = ClockSampler(local) if rank == 0 else None
    windows = []
    # ---- warm-up (also uploads the prescribed values once for the device-resident arm) ----
    pr.solve_frequency(freqs[kf_of(0)], mat, host=True)
    for s in range(max(args.warmup, 3) - 1):
        pr.solve_frequency(freqs[kf_of(s + 1)], mat, host=False)
    # ---- arm 1: device resident (value) ----
    acc = {}
    barrier(); w0 = time.time()
    ctx.mark(0)
    for s in range(args.steps):
        pr.solve_frequency(freqs[kf_of(s)], mat, host=False)
        st = pr.stats()
        for k in ("MS_ASSEMBLE", "MS_REGULAR", "MS_ADAPTIVE", "MS_SINGULAR", "MS_ZERO", "MS_LU", "MS_SOLVE", "MS_GEMM", "MS_PANEL", "MS_TRSM", "MS_SWAP",
                  "LAUNCHES", "LU_LAUNCHES", "GEMM_LAUNCHES", "GEMM_FLOPS", "GEMM_EXEC_FLOPS"):
            acc[k] = acc.get(k, 0.0) + st[k]
    ctx.mark(1)
    ms_dev = ctx.elapsed_ms(0, 1)
    barrier(); windows.append((w0, time.time()))
    ms_dev = reduce_max(ms_dev)
    # ---- arm 2: end to end through the C ABI with host buffers (pinned), solution gathered to the writer rank ----
    cv_pinned = torch.empty(md.cvalue.size * 2, dtype=torch.float64, pin_memory=True)
    cv_pinned.numpy()[:] = np.ascontiguousarray(md.cvalue).view(np.float64).ravel()
    pr._cv = cv_pinned.numpy().view(np.complex128)
    sweep = FrequencySweep(freqs, n, lambda kf, om: pr.solve_frequency(om, mat, host=True), rank=rank, world=world, dist=dist, device=dev)
    if dist is not None:   # untimed: the first gather sets up the NCCL channels of this group
        t = torch.zeros(2 * n, dtype=torch.float64, device=dev)
        dist.gather(t, [torch.empty_like(t) for _ in range(world)] if rank == 0 else None, dst=0)
    barrier(); w0 = time.time(); t0 = time.time()
    ctx.mark(2)
    for s in range(args.steps):
        sweep.round(s % sweep.n_rounds())
    ctx.mark(3)
    ms_e2e = ctx.elapsed_ms(2, 3)
    barrier(); wall_e2e = time.time() - t0; windows.append((w0, time.time()))
    ms_e2e = reduce_max(max(ms_e2e, wall_e2e * 1e3))
    # ---- arm 3 (N > 1): ONE frequency over all N GPUs (row-block assembly, NCCL redistribution, distributed LU); reported
    # ---- beside the sharded-sweep throughput as the latency of a single frequency
    single = None
    if world > 1 and os.environ.get("MFB_BENCH_SINGLE_FREQ", "1") != "0":
        single = single_frequency_arm(pr, ctx, capi, dist, torch, dev, rank, world, freqs[0], mat, args.steps, barrier, reduce_max)
    # ---- optional arm: the TWO-SEAM path exactly as the unmodified Fortran loop body would drive it: seam 1 fills the host's
    # ---- A_c, b_c (14.6 GB device -> host), seam 2 takes the host's A_c back (host -> device), factorises, returns L\U into A_c and x into b_c
    two_seam = None
    if args.two_seam and rank == 0:
        A_c = np.zeros((n, n), dtype=np.complex128, order="F"); b_c = np.zeros(n, dtype=np.complex128)
        pr.build_lse_mechanics_bem_harela(freqs[0], mat, out=(A_c, b_c))            # first touch of the host pages is not timed
        t0 = time.time(); pr.build_lse_mechanics_bem_harela(freqs[kf_of(1)], mat, out=(A_c, b_c)); t1 = time.time()
        x2 = pr.solve_lse_c(A_c, b_c); t2 = time.time()
        xr = pr.solve_frequency(freqs[kf_of(1)], mat, host=True)
        two_seam = {"ms_seam1_assemble_to_host": (t1 - t0) * 1e3, "ms_seam2_solve_lse_c_host_matrix": (t2 - t1) * 1e3, "solves_per_s": 1.0 / (t2 - t0),
                    "host_matrix_bytes_each_way": int(16 * n * n), "relerr_vs_fused_call": float(np.abs(x2 - xr).max() / np.abs(xr).max()),
                    "api": "mfb_harela3d_assemble (host A_c, b_c out) + mfb_zsolve (host A_c in, L\\U and x out): pageable host arrays, as a Fortran host holds them"}
        del A_c
    peaks = ctx.measure_peaks() if rank == 0 else None

    if rank == 0:
        K = args.steps
        value = world * K / (ms_dev * 1e-3)
        e2e = world * K / (ms_e2e * 1e-3)
        st = pr.stats()
        mp = {}
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        gemm_ms = acc["MS_GEMM"] / max(acc["GEMM_LAUNCHES"], 1)
        gemm_alg_tf = acc["GEMM_FLOPS"] / max(acc["MS_GEMM"], 1e-9) / 1e9      # 8mnk per complex product (what zgemm is credited with)
        gemm_tf = acc["GEMM_EXEC_FLOPS"] / max(acc["MS_GEMM"], 1e-9) / 1e9     # flops the tensor pipe executes (3M form: 6mnk)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_zgemm3m_tma")
        except Exception:
            pass
        roof = {"kernel": "k_zgemm3m_tma (LU trailing update: 3M complex product on DMMA.8x8x4, operands staged by TMA)", "bound": "tensor", "achieved": gemm_tf, "peak": peaks["dmma_tflops"],
                "unit": "TFLOP/s", "frac": gemm_tf / peaks["dmma_tflops"], "traffic": traffic,
                "note": "achieved = EXECUTED tensor-pipe flops (6mnk: three real products per complex product) / CUDA-event time of the trailing updates "
                        "(includes waiting for the look-ahead panel); algorithmic_tflops credits the standard 8mnk; traffic = DRAM bytes of one ncu capture at "
                        "M = N = 20480, K = 256 (an average trailing update), 1.115 x its algorithmic bytes (profiles/traffic.json)",
                "algorithmic_tflops": gemm_alg_tf, "algorithmic_frac": gemm_alg_tf / peaks["dmma_tflops"],
                "peak_source": "FP64 tensor (DMMA) micro-benchmark measured live on this GPU (mfb_measure_peaks); MEASURED_PEAKS.json carries no FP64 figure",
                "avg_launch_ms": gemm_ms, "launches_per_step": acc["GEMM_LAUNCHES"] / K, "share_of_step": acc["MS_GEMM"] / (ms_dev if world == 1 else acc["MS_ASSEMBLE"] + acc["MS_LU"] + acc["MS_SOLVE"]),
                "algorithmic_flops_per_step": acc["GEMM_FLOPS"] / K, "executed_flops_per_step": acc["GEMM_EXEC_FLOPS"] / K}
        asm_ms = acc["MS_ASSEMBLE"] / K
        reg_tf = st["FLOPS_REGULAR"] / (acc["MS_REGULAR"] / K) / 1e9
        hbm = mp.get("hbm_gbs")
        out = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / K,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
               "config": headline_config(name), "setup_s_once_per_mesh": t_setup,
               "clocks": clocks.summary(windows),
               "e2e": {"value": e2e, "unit": "solves/s", "h2d_bytes_per_step": int(md.cvalue.size * 16 + 4 * n + 1024), "d2h_bytes_per_step": int(16 * n + 4 * n + 4),
                       "ms_per_step": ms_e2e / K, "api": "mfb_harela3d_solve_frequency (host cvalue in, host x out)" + (" + NCCL gather of x to the writer rank" if world > 1 else "")},
               "gpu_launches": int(acc["LAUNCHES"] + acc["LU_LAUNCHES"]),
               "roofline": roof,
               "assembly": {"gentries_per_s": world * n * n / (asm_ms * 1e-3) / 1e9, "ms_per_frequency": asm_ms, "k_regular_tflops": reg_tf,
                            "k_regular_frac_fp64_fma_peak": reg_tf / peaks["dfma_tflops"], "algorithmic_flops_regular": st["FLOPS_REGULAR"],
                            "k_regular_note": "algorithmic flops = the reference's count per Gauss point (585 + 72 n, SURVEY.md 8d) summed over the plan; the kernel forms "
                                              "only the kernel combinations the element's boundary conditions use, so executed FP64 work is lower (ncu: profiles/)",
                            "matrix_write_gbs": 16.0 * n * n / (asm_ms * 1e-3) / 1e9, "hbm_frac": (16.0 * n * n / (asm_ms * 1e-3) / 1e9) / hbm if hbm else None,
                            "pairs_regular": st["PAIRS_REGULAR"], "pairs_adaptive": st["PAIRS_ADAPTIVE"], "pairs_singular": st["PAIRS_SINGULAR"],
                            "ms_regular": acc["MS_REGULAR"] / K, "ms_adaptive": acc["MS_ADAPTIVE"] / K, "ms_singular": acc["MS_SINGULAR"] / K, "ms_zero": acc["MS_ZERO"] / K},
               "lu": {"ms_per_frequency": acc["MS_LU"] / K, "tflops": 8.0 / 3.0 * n ** 3 / (acc["MS_LU"] / K) / 1e9,
                      "frac_fp64_tensor_peak": 8.0 / 3.0 * n ** 3 / (acc["MS_LU"] / K) / 1e9 / peaks["dmma_tflops"], "lu_only_solves_per_s": world * 1e3 / ((acc["MS_LU"] + acc["MS_SOLVE"]) / K),
                      "ms_panel_on_lookahead_stream": acc["MS_PANEL"] / K, "ms_trsm": acc["MS_TRSM"] / K, "ms_swap": acc["MS_SWAP"] / K, "ms_gemm": acc["MS_GEMM"] / K, "ms_zgetrs": acc["MS_SOLVE"] / K},
               "peaks_measured_live": peaks}
        if two_seam is not None:
            out["two_seam"] = two_seam
        if single is not None:
            single["speedup_vs_one_gpu"] = (ms_dev / K) / single["ms_per_frequency"] if "ms_per_frequency" in single else None
            out["single_frequency"] = single
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_arm(md, mat, freqs[0])
        print(json.dumps(out), flush=True)
    pr.close(); ctx.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--etype", default="tri3")
    ap.add_argument("--m", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--two-seam", action="store_true", help="also time the two-seam path with host matrices (14.6 GB each way per frequency)")
    ap.add_argument("--reference-sampled-only", action="store_true", help="--impl reference: skip the full-size measured step (every step a scaled sample)")
    ap.add_argument("--workload", default="harmonic", choices=["harmonic", "static", "acoustic", "coupled", "c1"],
                    help="harmonic: the headline 30k-DOF sweep (default); static: BASELINE config 2; acoustic: one frequency of the ME-TH-AC-001 room at ~10k DOF; coupled: BASELINE config 4 (fluid | poroelastic, ~20k DOF; device path opt-in)")
    ap.add_argument("--lanes", type=int, default=16, help="--workload c1: problems of the same mesh in flight per GPU (1 = one frequency at a time only)")
    ap.add_argument("--coupled-etype", default="quad9")
    ap.add_argument("--coupled-m", type=int, default=13, help="cells per face side of the two-box model (13 -> 21870 DOF)")
    ap.add_argument("--acoustic-etype", default="quad9")
    ap.add_argument("--acoustic-m", type=int, default=20)
    ap.add_argument("--static-etype", default="quad9")
    ap.add_argument("--static-m", type=int, default=11)
    args = ap.parse_args()
    if args.workload == "c1":
        run_c1(args)
    elif args.workload == "acoustic":
        run_acoustic(args)
    elif args.workload == "coupled":
        run_coupled(args)
    elif args.workload == "static":
        run_static(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
