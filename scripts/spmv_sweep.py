#!/usr/bin/env python
"""Standalone CSR SpMV / smoother-sweep GB/s sweep (BASELINE configs[1] 'standalone CSR SpMV
GB/s sweep'): resident matrix, CUDA events around `reps` launches, algorithmic bytes model
12 nnz + 4 (rows+1) + 8 cols + 8 rows (+8 rows when y is read, +16 rows for smoother sweeps).

    python scripts/spmv_sweep.py [--sizes 64,128,256] [--stencil 7|27] [--reps 50] [--kernels 0,2,11]
"""
import argparse, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
from faspsolver_b200 import api, problems as PB

def bytes_model(A, what):
    rows, cols = A.shape
    b = 12.0 * A.nnz + 4.0 * (rows + 1) + 8.0 * cols + 8.0 * rows
    if what in (1, 2): b += 8.0 * rows
    if what in (10, 11): b += 24.0 * rows
    return b

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="64,96,128,192,256")
    ap.add_argument("--stencil", type=int, default=7)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--warm", type=int, default=10)
    ap.add_argument("--kernels", default="0,2,11")
    ap.add_argument("--flush", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], help="library option key=value")
    a = ap.parse_args()
    L = api.lib(); api.check(L.fasp_cuda_init(0))
    for kv in a.opt:
        k, v = kv.split("=")
        api.check(L.fasp_cuda_set_option(k.encode(), float(v)))
    peak = 6538.3
    try: peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
    except Exception: pass
    names = {0: "y=Ax", 1: "y-=Ax", 2: "r=b-Ax", 10: "jacobi", 11: "l1"}
    for n in [int(s) for s in a.sizes.split(",")]:
        A = PB.poisson7(n) if a.stencil == 7 else PB.poisson27(n)
        h = L.fasp_cuda_dcsr_upload(A.ptr())
        if not h: raise RuntimeError(api.last_error())
        for what in [int(k) for k in a.kernels.split(",")]:
            ms = L.fasp_cuda_dcsr_time_kernel(h, what, a.warm, a.reps, a.flush)
            gbs = bytes_model(A, what) / ms * 1e-6
            print(json.dumps({"stencil": a.stencil, "n": n, "rows": A.shape[0], "nnz": A.nnz, "kernel": names[what],
                              "ms": round(ms, 5), "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 4)}), flush=True)
        L.fasp_cuda_dcsr_free(h)

if __name__ == "__main__":
    main()
