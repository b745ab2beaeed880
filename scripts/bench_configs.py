#!/usr/bin/env python
"""BASELINE configs 4 and 5 (and the 27-point SpMV sweep) on one B200: parity-checked solves with
device time, next to the reference's sequential CPU time on a bounded sample.

  config 4: 3-D 7-point upwind convection-diffusion n^3, AMG-GMRES(30), classical RS, polynomial
            smoother (degree 3), tol 1e-8                    (fasp_cuda_krylov_amg_* , CSR path)
  config 5: 3x3-block 7-point "black-oil shaped" system on n^3 block rows, UA-AMG (VMB) +
            block Jacobi + VGMRES(30), tol 1e-8               (BSR path) + BSR SpMV GB/s

    python scripts/bench_configs.py --c4 256 --c5 160 --spmv27 128
"""
import argparse, ctypes as C, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import bench as B
from faspsolver_b200 import api, problems as PB, fasp_types as T


def timed_solves(solver, b, it, reps=3):
    zero = np.zeros_like(b)
    st, x = solver.solve(b, zero, it)
    ms = []
    for _ in range(reps):
        st, x = solver.solve(b, zero, it)
        ms.append(solver.stat(2))
    return st, x, float(np.mean(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c4", type=int, default=0)
    ap.add_argument("--c5", type=int, default=0)
    ap.add_argument("--spmv27", type=int, default=0)
    a = ap.parse_args()
    L = api.lib(); api.check(L.fasp_cuda_init(0))
    hf = B.host_fasp()
    peak, _ = B.peaks()
    if a.c4:
        A = PB.convdiff7(a.c4); b = np.ones(A.shape[0])
        amg = hf.amg_param(print_level=0, smoother=T.SMOOTHER_POLY, polynomial_degree=3)
        it = hf.its_param(itsolver_type=T.SOLVER_GMRES, restart=30, tol=1e-8, maxit=500, print_level=0)
        t = time.time(); mgl = hf.amg_setup(A, amg); ts = time.time() - t
        s = api.KrylovAmgSolver(mgl, amg)
        st, x, ms = timed_solves(s, b, it)
        rel = float(np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b))
        it2 = hf.its_param(itsolver_type=T.SOLVER_VGMRES, restart=30, tol=1e-8, maxit=500, print_level=0)
        st2, x2, ms2 = timed_solves(s, b, it2)
        print(json.dumps({"config": 4, "n": a.c4, "rows": A.shape[0], "nnz": A.nnz, "levels": len(api.hierarchy_info(mgl)),
                          "host_setup_s": round(ts, 1), "gmres30_iters": st, "gmres30_ms": round(ms, 3), "true_relres": rel,
                          "vgmres30_iters": st2, "vgmres30_ms": round(ms2, 3)}), flush=True)
        s.close(); hf.amg_free(mgl, amg)
    if a.c5:
        A, b = PB.blockoil7(a.c5)
        amg = hf.amg_param(print_level=0, AMG_type=T.UA_AMG, aggregation_type=T.VMB, smoother=T.SMOOTHER_JACOBI, coarse_dof=100)
        it = hf.its_param(itsolver_type=T.SOLVER_VGMRES, restart=30, tol=1e-8, maxit=500, print_level=0)
        h = L.fasp_cuda_dbsr_upload(A.ptr())
        out = {"config": 5, "n": a.c5, "block_rows": A.ROW, "blocks": A.NNZ, "nb": 3}
        by = (8.0 * 9 + 4) * A.NNZ + 4.0 * (A.ROW + 1) + 8.0 * 3 * (A.COL + A.ROW)
        for what, nm, extra in ((0, "bsr_mxv", 0.0), (2, "bsr_resid", 24.0 * A.ROW), (10, "bsr_jacobi", 24.0 * A.ROW + 72.0 * A.ROW)):
            ms = L.fasp_cuda_dbsr_time_kernel(h, what, 5, 30, 0)
            out[nm + "_ms"] = round(ms, 4); out[nm + "_GBps"] = round((by + extra) / ms * 1e-6, 1)
            out[nm + "_frac_of_measured_peak"] = round((by + extra) / ms * 1e-6 / peak, 3)
        L.fasp_cuda_dbsr_free(h)
        t = time.time(); mgl = hf.bamg_setup(A, amg); ts = time.time() - t
        s = api.KrylovAmgSolver(mgl, amg, bsr=True)
        st, x, ms = timed_solves(s, b, it)
        rel = float(np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b))
        out.update({"levels": int(mgl[0].num_levels), "host_setup_s": round(ts, 1), "vgmres30_iters": st, "vgmres30_ms": round(ms, 3), "true_relres": rel})
        print(json.dumps(out), flush=True)
        s.close(); hf.bamg_free(mgl, amg)
    if a.spmv27:
        A = PB.poisson27(a.spmv27)
        h = L.fasp_cuda_dcsr_upload(A.ptr())
        by = 12.0 * A.nnz + 4.0 * (A.shape[0] + 1) + 16.0 * A.shape[0]
        ms = L.fasp_cuda_dcsr_time_kernel(h, 0, 5, 30, 0)
        print(json.dumps({"spmv27": a.spmv27, "rows": A.shape[0], "nnz": A.nnz, "ms": round(ms, 4), "GBps": round(by / ms * 1e-6, 1),
                          "frac_of_measured_peak": round(by / ms * 1e-6 / peak, 3)}), flush=True)
        L.fasp_cuda_dcsr_free(h)

if __name__ == "__main__":
    main()
