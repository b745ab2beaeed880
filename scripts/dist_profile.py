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
ap.add_argument("--opts", default="", help="option sets to time before the profile, e.g. 'p2p_fused=1;p2p_fused=0'")
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
for cfg in [c for c in a.opts.split(";") if c]:
    for kv in cfg.split(","):
        k, v = kv.split("=")
        api.check(L.fasp_cuda_set_option(k.encode(), float(v)))
    wl = []
    for _ in range(3):
        st_, x_ = s.solve(bl, z.copy(), it); wl.append((st_, round(s.stat(2), 2)))
    MG.barrier()
    tt = []
    for _ in range(5):
        st_, x_ = s.solve(bl, z.copy(), it)
        tt.append(s.stat(2)); wl.append((st_, round(s.stat(2), 2)))
    print("[%s] rank %d solves (status, ms): %s" % (cfg, rank, wl), flush=True)
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, (s.row0, np.array(x_, copy=True)))
    if rank == 0:
        full = np.empty(A.shape[0])
        for p0, xp in parts: full[p0:p0 + xp.size] = xp
        rel = float(np.linalg.norm(b - A.to_scipy() @ full) / np.linalg.norm(b))
        print("[%s] world %d agg_rows %d: solve %.3f ms (min %.3f), iterations %d, true relres %.3e" % (cfg, world, a.agg_rows, float(np.mean(tt)), min(tt), st_, rel), flush=True)
for _ in range(3): s.solve(bl, z.copy(), it)
MG.barrier()
t = [s.solve(bl, z.copy(), it) and s.stat(2) for _ in range(3)]
L.fasp_cuda_set_option(b"profile", 1.0); L.fasp_cuda_profile_dump(None, 0)
MG.barrier()
st, _ = s.solve(bl, z.copy(), it)
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
