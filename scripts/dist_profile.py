#!/usr/bin/env python
"""Per-kind time table of one distributed solve (run under torchrun): matrix kernels per level,
halo exchanges (400), all-reduces (401), all-gathers (402). Graphs are off in profile mode."""
import argparse, ctypes as C, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import bench as B
from faspsolver_b200 import api, multigpu as MG, fasp_types as T

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--agg-rows", type=int, default=200000)
a = ap.parse_args()
rank, world, local = MG.init_comm()
L = api.lib()
hf = B.host_fasp()
A, b = MG._quiet(B.build_problem, a.size)
amg, it = B.amg_recipe(hf)
mgl = hf.amg_setup(A, amg)
s = MG.DistSolver(mgl, amg, agg_rows=a.agg_rows)
hf.amg_free(mgl, amg)
bl = np.ascontiguousarray(b[s.row0:s.row1]); z = np.zeros(s.row1 - s.row0)
for _ in range(3): s.solve(bl, z, it)
MG.barrier()
t = [s.solve(bl, z, it) and s.stat(2) for _ in range(3)]
L.fasp_cuda_set_option(b"profile", 1.0); L.fasp_cuda_profile_dump(None, 0)
MG.barrier()
st, _ = s.solve(bl, z, it)
buf = C.create_string_buffer(64 << 20); L.fasp_cuda_profile_dump(buf, len(buf)); L.fasp_cuda_set_option(b"profile", 0.0)
recs = [ln.split() for ln in buf.value.decode().splitlines()]
recs = [(int(k), int(r), int(z_), float(ms)) for k, r, z_, ms, by in recs]
if rank == 0:
    print("world %d agg_rows %d: graph solve %.3f ms, iterations %d, profiled (no graph) solve %.3f ms" % (world, a.agg_rows, float(np.mean(t)), st, s.stat(2)))
    grp = {}
    for k, r, z_, ms in recs:
        key = ("halo" if k == 400 else "allreduce" if k == 401 else "allgather" if k == 402 else "dense" if k == 100 else "matrix", r if k < 400 else 0, z_ if k < 400 else 0)
        e = grp.setdefault(key, [0, 0.0]); e[0] += 1; e[1] += ms
    tot = sum(v[1] for v in grp.values())
    for key, (c, ms) in sorted(grp.items(), key=lambda kv: -kv[1][1])[:18]:
        print("  %-10s rows %9d nnz %10d  launches %4d  %8.3f ms  avg %7.1f us" % (key[0], key[1], key[2], c, ms, ms / c * 1e3))
    print("  total profiled kernel+comm time %.3f ms" % tot)
s.close()
L.fasp_cuda_comm_finalize()
