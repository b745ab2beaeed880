#!/usr/bin/env python
"""Does the speed of the level-0 kernels depend on WHERE the arrays of the hierarchy land in device memory?
One host hierarchy (7-pt 256^3), then for a list of dummy allocations made before the upload: upload, warm up,
one profiled solve, per-mode mean times of the level-0 kernel. (25d178e made the L1 sweep 10 % slower without
touching its SASS; the only change before the upload were two small allocations in ensure_init.)"""
import ctypes as C, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import bench as B
from faspsolver_b200 import api, fasp_types as T

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pads = [int(float(x) * (1 << 17)) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,1,2,3,5,8,13,64,127,1000".split(","))]  # MiB -> doubles
L = api.lib()
api.check(L.fasp_cuda_init(0))
hf = B.host_fasp()
A, b = B.build_problem(n)
N = A.shape[0]
amg, it = B.amg_recipe(hf)
mgl = hf.amg_setup(A, amg)
kind = {0: "mxv", 2: "resid", 4: "l1"}
keep = []
for pad in pads:
    d = L.fasp_cuda_dvec_alloc(max(pad, 1))
    keep.append(d)          # never freed: every later allocation really moves
    s = api.KrylovAmgSolver(mgl, amg)
    d_b, d_x = L.fasp_cuda_dvec_alloc(N), L.fasp_cuda_dvec_alloc(N)
    api.check(L.fasp_cuda_dvec_h2d(d_b, T.as_preal(b), N))
    zero = np.zeros(N)
    ms = []
    for _ in range(3):
        api.check(L.fasp_cuda_dvec_h2d(d_x, T.as_preal(zero), N))
        st = s.solve_dev(d_b, d_x, it)
        ms.append(s.stat(2))
    L.fasp_cuda_set_option(b"profile", 1.0); L.fasp_cuda_profile_dump(None, 0)
    api.check(L.fasp_cuda_dvec_h2d(d_x, T.as_preal(zero), N))
    s.solve_dev(d_b, d_x, it)
    buf = C.create_string_buffer(64 << 20); L.fasp_cuda_profile_dump(buf, len(buf)); L.fasp_cuda_set_option(b"profile", 0.0)
    recs = [ln.split() for ln in buf.value.decode().splitlines()]
    recs = [(int(k), int(r), int(z), float(m)) for k, r, z, m, by in recs]
    out = {"pad_MiB": pad / (1 << 17), "solve_ms": round(min(ms), 3), "iters": st}
    for (rows, nnz, tag) in ((N, A.nnz, "L0"), (mgl[1].A.row, mgl[1].A.nnz, "L1"), (mgl[2].A.row, mgl[2].A.nnz, "L2")):
        for k, nm in kind.items():
            v = [m for kk, r, z, m in recs if kk == k and r == rows and z == nnz]
            if v:
                v = [x for x in v if x >= 0.25 * np.percentile(v, 90)]
                out["%s_%s_us" % (tag, nm)] = round(float(np.mean(v)) * 1e3, 1)
    print(json.dumps(out), flush=True)
    L.fasp_cuda_dvec_free(d_b); L.fasp_cuda_dvec_free(d_x)
    s.close()
