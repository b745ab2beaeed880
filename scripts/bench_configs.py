#!/usr/bin/env python
"""BASELINE configs 4 and 5 on one B200, one JSON object per run (bench.py starts one process per
config and embeds the objects as `extra.config4` / `extra.config5`).

  config 4: 3-D 7-point upwind convection-diffusion n^3 (default 256^3, nonsymmetric), AMG-preconditioned
            GMRES(30) (+ VGMRES(30)), classical RS hierarchy from FASP's host setup, polynomial smoother of
            degree 3, tol 1e-8                                                      (CSR path)
  config 5: 3x3-block 7-point "black-oil shaped" system on n^3 block rows (default 272^3 = 20.1 M block
            rows, 140 M blocks, 10.7 GB), BSR SpMV GB/s + UA-AMG (VMB) / block-Jacobi / VGMRES(30), tol 1e-8
                                                                                    (BSR path)
Each object carries: device ms of the solve (CUDA events, mean after warm-up), iterations, true residual of
the returned solution, e2e ms through the host-pointer call, a per-matrix kernel table + roofline of the
dominant kernel from one profiled solve (CUDA events around every launch, graphs off), and a CPU baseline:
the reference's own cycle (sequential libfasp, same hierarchy) timed ONCE and scaled by the number of
preconditioner applications of the solve (a bounded sample: one CPU cycle at these sizes is 10-25 s).

    python scripts/bench_configs.py --config 4 [--n 256]
    python scripts/bench_configs.py --config 5 [--n 272]
"""
import argparse
import contextlib
import ctypes as C
import fcntl
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

import bench as B  # noqa: E402
from faspsolver_b200 import api, fasp_types as T, problems as PB  # noqa: E402


def log(*a):
    print("[config]", *a, file=sys.stderr, flush=True)


@contextlib.contextmanager
def gpu_phase(lock_path):
    """Serialise the GPU timing phases of concurrently running config processes (their host setups overlap)."""
    if not lock_path:
        yield
        return
    with open(lock_path, "w") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)


def timed_solves(solver, b, it, reps=3, warm=2):
    """Device ms (CUDA events around the Krylov loop) and e2e ms (host-pointer call incl. H2D b, x0 and D2H x)."""
    zero = np.zeros_like(b)
    out = np.zeros_like(b)
    api.pin_host(b)
    api.pin_host(out)
    try:
        for _ in range(warm):
            st, x = solver.solve(b, zero, it, out=out)
            if st < 0:
                raise RuntimeError("solve failed: %d %s" % (st, api.last_error()))
        ms, e2e = [], []
        for _ in range(reps):
            st, x = solver.solve(b, zero, it, out=out)
            ms.append(solver.stat(2))
            e2e.append(solver.stat(4))
    finally:
        api.unpin_host(b)
        api.unpin_host(out)
    return st, x.copy(), float(np.mean(ms)), float(np.mean(e2e)), int(solver.stat(3))


def profiled_solve(L, solver, b, it):
    """One more solve with CUDA events around every matrix kernel (graphs off): per-matrix table."""
    L.fasp_cuda_set_option(b"profile", 1.0)
    L.fasp_cuda_profile_dump(None, 0)
    solver.solve(b, np.zeros_like(b), it)
    buf = C.create_string_buffer(64 << 20)
    L.fasp_cuda_profile_dump(buf, len(buf))
    L.fasp_cuda_set_option(b"profile", 0.0)
    recs = [ln.split() for ln in buf.value.decode().splitlines()]
    recs = [(int(k), int(r), int(z), float(ms), float(by)) for k, r, z, ms, by in recs]
    groups = {}
    for rec in recs:
        groups.setdefault(rec[:3], []).append(rec[3])
    med = {k: float(np.percentile(v, 90)) for k, v in groups.items()}   # gated launches can be the majority
    # drop gated launches that returned at once (conditional tag / after convergence)
    recs = [rec for rec in recs if rec[0] % 100 < 50 and rec[3] >= 0.25 * med[rec[:3]] and rec[4] > 0]
    levels = {}
    for k, r_, z, ms, by in recs:
        e = levels.setdefault((r_, z), [0, 0.0, 0.0])
        e[0] += 1
        e[1] += ms
        e[2] += by
    table = [{"rows": r_, "nnz": z, "launches": c, "ms": round(ms, 4), "GBps": round(by / ms * 1e-6, 1) if ms else 0}
             for (r_, z), (c, ms, by) in sorted(levels.items(), key=lambda kv: -kv[1][1])]
    return recs, table


def roofline_of(recs, rows, nnz, what):
    peak, src = B.peaks()
    top = [r for r in recs if r[1] == rows and r[2] == nnz]
    tot = sum(r[3] for r in recs)
    if not top:
        return None
    t = sum(r[3] for r in top)
    ach = sum(r[4] for r in top) / t * 1e-6
    return {"bound": "hbm", "kernel": what, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "frac_of_nominal_8000": ach / 8000.0, "traffic": None, "peak_source": src,
            "algorithmic_bytes_per_launch": float(np.mean([r[4] for r in top])), "launch_ms": t / len(top),
            "launches": len(top), "share_of_matrix_kernel_time": t / tot if tot else None,
            "all_matrix_kernels_GBps": sum(r[4] for r in recs) / tot * 1e-6 if tot else None}


def config4(a, L, hf):
    from oracle.ref import RefFasp
    n = a.n or 256
    t = time.time()
    A = PB.convdiff7(n)
    b = np.ones(A.shape[0])
    log("conv-diff %d^3 generated in %.1fs" % (n, time.time() - t))
    amg = hf.amg_param(print_level=0, smoother=T.SMOOTHER_POLY, polynomial_degree=3)
    it = hf.its_param(itsolver_type=T.SOLVER_GMRES, restart=30, tol=1e-8, maxit=500, print_level=0)
    itv = hf.its_param(itsolver_type=T.SOLVER_VGMRES, restart=30, tol=1e-8, maxit=500, print_level=0)
    t = time.time()
    mgl = hf.amg_setup(A, amg)
    ts = time.time() - t
    info = api.hierarchy_info(mgl)
    log("host RS setup %.1fs, %d levels" % (ts, len(info)))
    with gpu_phase(a.lock):
        t = time.time()
        s = api.KrylovAmgSolver(mgl, amg)
        tu = time.time() - t
        st, x, ms, e2e, launches = timed_solves(s, b, it)
        st2, x2, ms2, e2e2, _ = timed_solves(s, b, itv, reps=2, warm=2)
        recs, table = profiled_solve(L, s, b, it)
        s.close()
    rel = float(np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b))
    rel2 = float(np.linalg.norm(b - A.to_scipy() @ x2) / np.linalg.norm(b))
    roof = roofline_of(recs, A.shape[0], A.nnz, "csr_pipe_kernel on the level-0 matrix (SpMV, residual and the "
                       "polynomial-smoother steps), CUDA events around every launch of one solve")
    # CPU baseline: ONE application of the reference's own V-cycle preconditioner on the same hierarchy
    ref = RefFasp()
    pcdata = T.precond_data()
    ref.L.fasp_param_amg_to_prec(C.byref(pcdata), C.byref(amg))
    pcdata.max_levels = mgl[0].num_levels
    pcdata.mgl_data = mgl
    r, z = np.ones(A.shape[0]), np.zeros(A.shape[0])
    t = time.perf_counter()
    ref.L.fasp_precond_amg(T.as_preal(r), T.as_preal(z), C.cast(C.byref(pcdata), C.c_void_p))
    dt = time.perf_counter() - t
    napply = st + 1 + (st - 1) // 30   # one per iteration + one per restart cycle (KryPgmres.c)
    cpu = {"value": dt * napply * 1e3, "unit": "ms", "cores": 1, "kind": "reference",
           "sample": "ONE fasp_precond_amg application (V(1,1), polynomial smoother degree 3; sequential libfasp, same "
                     "hierarchy) timed = %.2f s, times the %d preconditioner applications of the %d-iteration GMRES(30) "
                     "solve; the Krylov part (SpMV + Gram-Schmidt) is not included" % (dt, napply, st)}
    hf.amg_free(mgl, amg)
    return {"workload": "configs[3]: 3D 7-point upwind convection-diffusion %d^3 (%d rows, %d nnz, cell Peclet "
                        "0.5/0.25/0.125), rhs=1, AMG-GMRES(30) tol 1e-8, classical RS (FASP host setup), V(1,1) polynomial "
                        "smoother degree 3" % (n, A.shape[0], A.nnz),
            "metric": "amg_gmres30_solve_time_convdiff3d_7pt", "unit": "ms", "value": ms, "dtype": "f64",
            "iterations": int(st), "true_relres": rel, "levels": len(info), "host_setup_s": round(ts, 1),
            "operator_complexity": round(sum(z for _, z, _ in info) / float(A.nnz), 2),
            "grid_complexity": round(sum(r for r, _, _ in info) / float(A.shape[0]), 2),
            "upload_s": round(tu, 2), "gpu_launches_per_solve": launches,
            "e2e": {"value": e2e, "unit": "ms", "h2d_bytes_per_step": 16 * A.shape[0], "d2h_bytes_per_step": 8 * A.shape[0]},
            "vgmres30": {"value": ms2, "iterations": int(st2), "true_relres": rel2, "e2e": e2e2},
            "roofline": roof, "cpu_baseline": cpu, "levels_table": table[:24]}


def blockoil7_lean(n):
    """PB.blockoil7 without the 3 x NNZ x 9 temporaries (config 5 at 272^3 is 10 GB of blocks)."""
    scal = PB.poisson7(n, scaled=False)
    ia, ja = scal.ia, scal.ja
    NNZ = int(ia[-1])
    val = np.empty((NNZ, 3, 3))
    N = n ** 3
    step = 1 << 20
    for r0 in range(0, N, step):
        r1 = min(N, r0 + step)
        k0, k1 = int(ia[r0]), int(ia[r1])
        rows = np.repeat(np.arange(r0, r1, dtype=np.int64), np.diff(ia[r0:r1 + 1]))
        d = ja[k0:k1].astype(np.int64) - rows
        s = np.full(k1 - k0, -1.0)
        s[d == -1] = -1.2
        s[d == 1] = -0.8
        blk = val[k0:k1]
        np.multiply(s[:, None, None], PB._B[None, :, :], out=blk)
        blk[d == 0] = 6.0 * PB._B + PB._D
    A = T.BSR(N, N, 3, ia, ja, val)
    rhs = 1.0 + 0.01 * (np.arange(3 * N) % 7)
    return A, rhs


def config5(a, L, hf):
    from oracle.ref import RefFasp
    n = a.n or 272
    try:
        import psutil
        gb = psutil.virtual_memory().available / 2 ** 30
    except Exception:
        gb = 1e9
    need = 5.5 * 76 * 7 * n ** 3 / 2 ** 30   # matrix + FASP's copy + hierarchy + staging
    note = None
    if gb < need:
        n2 = int((gb / need) ** (1 / 3) * n) // 16 * 16
        note = "host has %.0f GB available, %.0f GB needed for %d^3: reduced to %d^3" % (gb, need, n, n2)
        log(note)
        n = max(n2, 32)
    t = time.time()
    A, b = blockoil7_lean(n)
    log("block system %d^3 generated in %.1fs" % (n, time.time() - t))
    peak, _ = B.peaks()
    amg = hf.amg_param(print_level=0, AMG_type=T.UA_AMG, aggregation_type=T.VMB, smoother=T.SMOOTHER_JACOBI,
                       coarse_dof=100)
    it = hf.its_param(itsolver_type=T.SOLVER_VGMRES, restart=30, tol=1e-8, maxit=500, print_level=0)
    t = time.time()
    mgl = hf.bamg_setup(A, amg)
    ts = time.time() - t
    nl = int(mgl[0].num_levels)
    log("host UA setup %.1fs, %d levels" % (ts, nl))
    out = {}
    by = (8.0 * 9 + 4) * A.NNZ + 4.0 * (A.ROW + 1) + 8.0 * 3 * (A.COL + A.ROW)
    with gpu_phase(a.lock):
        h = L.fasp_cuda_dbsr_upload(A.ptr())
        if not h:
            raise RuntimeError(api.last_error())
        spmv = {}
        for what, nm, extra in ((0, "mxv", 0.0), (2, "resid", 24.0 * A.ROW), (10, "jacobi", 24.0 * A.ROW + 72.0 * A.ROW)):
            ms = L.fasp_cuda_dbsr_time_kernel(h, what, 5, 30, 0)
            spmv[nm] = {"ms": round(ms, 4), "GBps": round((by + extra) / ms * 1e-6, 1),
                        "frac_of_measured_peak": round((by + extra) / ms * 1e-6 / peak, 3),
                        "frac_of_nominal_8000": round((by + extra) / ms * 1e-6 / 8000.0, 3)}
        L.fasp_cuda_dbsr_free(h)
        t = time.time()
        s = api.KrylovAmgSolver(mgl, amg, bsr=True)
        tu = time.time() - t
        st, x, ms, e2e, launches = timed_solves(s, b, it)
        recs, table = profiled_solve(L, s, b, it)
        s.close()
    import scipy.sparse as sp
    S = sp.bsr_matrix((A.val.reshape(-1, 3, 3), A.ja, A.ia), shape=(3 * A.ROW, 3 * A.COL))
    rel = float(np.linalg.norm(b - S @ x) / np.linalg.norm(b))
    roof = roofline_of(recs, A.ROW, A.NNZ, "bsr_pipe_kernel on the level-0 block matrix (SpMV, residual, block-Jacobi "
                       "sweep), CUDA events around every launch of one solve")
    # CPU baseline: ONE application of the reference's own BSR cycle on the same hierarchy
    ref = RefFasp()
    nn = 3 * A.ROW
    bv = np.ctypeslib.as_array(mgl[0].b.val, shape=(nn,))
    xv = np.ctypeslib.as_array(mgl[0].x.val, shape=(nn,))
    bv[:] = b
    xv[:] = 0.0
    amg_c = hf.amg_param(print_level=0, AMG_type=T.UA_AMG, aggregation_type=T.VMB, smoother=T.SMOOTHER_JACOBI,
                         coarse_dof=100, tol=1e-6)
    t = time.perf_counter()
    ref.L.fasp_solver_mgcycle_bsr(mgl, C.byref(amg_c))
    dt = time.perf_counter() - t
    napply = st + 1 + (st - 1) // 30
    cpu = {"value": dt * napply * 1e3, "unit": "ms", "cores": 1, "kind": "reference",
           "sample": "ONE fasp_solver_mgcycle_bsr (V-cycle, block Jacobi; sequential libfasp, same hierarchy) timed = %.2f s, "
                     "times the %d preconditioner applications of the %d-iteration VGMRES(30) solve; Krylov part not "
                     "included" % (dt, napply, st)}
    hf.bamg_free(mgl, amg)
    out.update({"workload": "configs[4]: 3x3-block 7-point black-oil-shaped system on %d^3 block rows (%d block rows, %d "
                            "blocks, %.2f GB), BSR SpMV + UA-AMG (VMB aggregation, FASP host setup) / block Jacobi / "
                            "VGMRES(30) tol 1e-8" % (n, A.ROW, A.NNZ, by / 1e9),
                "metric": "amg_vgmres30_solve_time_blockoil_bsr3", "unit": "ms", "value": ms, "dtype": "f64",
                "iterations": int(st), "true_relres": rel, "levels": nl, "host_setup_s": round(ts, 1),
                "upload_s": round(tu, 2), "gpu_launches_per_solve": launches, "size_note": note,
                "e2e": {"value": e2e, "unit": "ms", "h2d_bytes_per_step": 16 * nn, "d2h_bytes_per_step": 8 * nn},
                "bsr_spmv": spmv, "roofline": roof, "cpu_baseline": cpu, "levels_table": table[:16]})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[4, 5])
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--lock", default="")
    ap.add_argument("--opt", action="append", default=[])
    a = ap.parse_args()
    L = api.lib()
    api.check(L.fasp_cuda_init(int(os.environ.get("LOCAL_RANK", "0"))))
    for kv in a.opt:
        k, v = kv.split("=")
        api.check(L.fasp_cuda_set_option(k.encode(), float(v)))
    hf = B.host_fasp()
    out = config4(a, L, hf) if a.config == 4 else config5(a, L, hf)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
