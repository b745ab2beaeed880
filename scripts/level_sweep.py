#!/usr/bin/env python
"""Per-level kernel GB/s on a real classical-AMG hierarchy (FASP host setup of the n^3 7-point
Poisson problem): uploads every A_l / P_l / R_l separately and times the resident kernels.

    python scripts/level_sweep.py [--n 128] [--reps 30] [--levels 0,1,2] [--kernels 0,11]
"""
import argparse, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import ctypes as C
import numpy as np
from faspsolver_b200 import api, problems as PB, fasp_types as T

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--warm", type=int, default=5)
    ap.add_argument("--levels", default="")
    ap.add_argument("--kernels", default="0,11")
    ap.add_argument("--ops", default="A,P,R")
    ap.add_argument("--vec-min-avg", type=int, default=-1)
    ap.add_argument("--opt", action="append", default=[], help="library option key=value")
    ap.add_argument("--lprs", default="", help="comma list of vec_lpr overrides to compare per level (0 = default choice)")
    ap.add_argument("--optsets", default="", help="option sets to compare per matrix: 'k=v,k=v;k=v' (applied before the upload)")
    ap.add_argument("--reset", default="", help="options restored after every set: 'k=v,k=v'")
    a = ap.parse_args()
    L = api.lib(); api.check(L.fasp_cuda_init(0))
    for kv in a.opt:
        k, v = kv.split("=")
        api.check(L.fasp_cuda_set_option(k.encode(), float(v)))
    if a.vec_min_avg >= 0: L.fasp_cuda_set_option(b"vec_min_avg", float(a.vec_min_avg))
    hf = api.HostFasp(str(ROOT / "oracle" / "_ref" / "libfasp_seq.so"))
    A = PB.poisson7(a.n)
    amg = hf.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
    mgl = hf.amg_setup(A, amg)
    nl = mgl[0].num_levels
    levels = [int(s) for s in a.levels.split(",")] if a.levels else list(range(nl))
    names = {0: "y=Ax", 1: "y-=Ax", 2: "r=b-Ax", 10: "jacobi", 11: "l1"}
    for l in levels:
        for op in a.ops.split(","):
            if op != "A" and l >= nl - 1: continue
            m = getattr(mgl[l], op)
            sets = [c for c in a.optsets.split(";") if c] or [""]
            for oset in sets:
              for kv in [x for x in a.reset.split(",") if x] + [x for x in oset.split(",") if x]:
                  k, v = kv.split("=")
                  api.check(L.fasp_cuda_set_option(k.encode(), float(v)))
              for lpr in ([int(v) for v in a.lprs.split(",")] if a.lprs else [None]):
                if lpr is not None: L.fasp_cuda_set_option(b"vec_lpr", float(lpr))
                h = L.fasp_cuda_dcsr_upload(C.byref(m))
                if not h: raise RuntimeError(api.last_error())
                kernels = [int(k) for k in a.kernels.split(",")] if op == "A" else ([1] if op == "P" else [0])
                for what in kernels:
                    ms = L.fasp_cuda_dcsr_time_kernel(h, what, a.warm, a.reps, 0)
                    by = 12.0 * m.nnz + 4.0 * (m.row + 1) + 8.0 * m.col + 8.0 * m.row
                    if what in (1, 2): by += 8.0 * m.row
                    if what in (10, 11): by += 24.0 * m.row
                    print(json.dumps({"level": l, "op": op, "rows": m.row, "cols": m.col, "nnz": m.nnz,
                                      "nnz_per_row": round(m.nnz / max(1, m.row), 1), "kernel": names[what], "lpr": lpr,
                                      "opts": oset, "us": round(ms * 1e3, 2), "GBps": round(by / ms * 1e-6, 1)}), flush=True)
                L.fasp_cuda_dcsr_free(h)
    hf.amg_free(mgl, amg)

if __name__ == "__main__":
    main()
