#!/usr/bin/env python
"""One hierarchy, several library option sets: per-level kernel table of the AMG-PCG solve.
    python scripts/solve_sweep.py --n 256 --configs "vec_min_avg=24;vec_min_avg=24,rowwise_max=64"
"""
import argparse, ctypes as C, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import bench as B
from faspsolver_b200 import api, fasp_types as T

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--configs", default="")
    ap.add_argument("--top", type=int, default=14)
    a = ap.parse_args()
    L = api.lib(); api.check(L.fasp_cuda_init(0))
    hf = B.host_fasp()
    A, b = B.build_problem(a.n)
    n = A.shape[0]
    amg, it = B.amg_recipe(hf)
    mgl = hf.amg_setup(A, amg)
    d_b = L.fasp_cuda_dvec_alloc(n); d_x = L.fasp_cuda_dvec_alloc(n)
    api.check(L.fasp_cuda_dvec_h2d(d_b, T.as_preal(b), n))
    zero = np.zeros(n)
    for cfg in a.configs.split(";"):
        for kv in [c for c in cfg.split(",") if c]:
            k, v = kv.split("=")
            api.check(L.fasp_cuda_set_option(k.encode(), float(v)))
        solver = api.KrylovAmgSolver(mgl, amg)
        def solve():
            api.check(L.fasp_cuda_dvec_h2d(d_x, T.as_preal(zero), n))
            st = solver.solve_dev(d_b, d_x, it)
            return st, solver.stat(2)
        solve(); times = [solve()[1] for _ in range(3)]
        L.fasp_cuda_set_option(b"profile", 1.0); L.fasp_cuda_profile_dump(None, 0)
        st, _ = solve()
        buf = C.create_string_buffer(64 << 20); L.fasp_cuda_profile_dump(buf, len(buf)); L.fasp_cuda_set_option(b"profile", 0.0)
        recs = [ln.split() for ln in buf.value.decode().splitlines()]
        recs = [(int(k), int(r), int(z), float(ms), float(by)) for k, r, z, ms, by in recs]
        groups = {}
        for rec in recs: groups.setdefault(rec[:3], []).append(rec[3])
        med = {k: float(np.median(v)) for k, v in groups.items()}
        recs = [r for r in recs if r[0] < 50 and r[3] >= 0.25 * med[r[:3]]]
        lv = {}
        for k, r_, z, ms, by in recs:
            e = lv.setdefault((r_, z), [0, 0.0, 0.0]); e[0] += 1; e[1] += ms; e[2] += by
        print("== config [%s]: iters %d, solve %.3f ms, matrix-kernel sum %.3f ms" % (cfg, st, float(np.mean(times)), sum(e[1] for e in lv.values())))
        for (r_, z), (c, ms, by) in sorted(lv.items(), key=lambda kv: -kv[1][1])[:a.top]:
            print("   %9d %10d %3d %8.3f ms %7.0f GB/s  nnz/row %.1f" % (r_, z, c, ms, by / ms * 1e-6, z / r_))
        solver.close()
    hf.amg_free(mgl, amg)

if __name__ == "__main__":
    main()
