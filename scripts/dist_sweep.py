#!/usr/bin/env python
"""Multi-GPU solve-time sweep (run under torchrun): one host hierarchy (built by rank 0, mapped by the others),
then for every agglomeration threshold a partitioned solver and, on it, every option set:

    torchrun --nproc-per-node N scripts/dist_sweep.py --size 256 --agg-list 8000,50000,300000 \
        --opts "overlap=1;overlap=0"

Prints device ms of the solve (max over ranks, mean of 5 after 3 warm-ups), iterations and the true residual of the
assembled solution; `--profile` adds the per-kind time table of one profiled solve of the first configuration."""
import argparse, ctypes as C, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import bench as B
from faspsolver_b200 import api, multigpu as MG, fasp_types as T

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--agg-list", default="8000")
ap.add_argument("--opts", default="overlap=1", help="option sets separated by ';', options by ','")
ap.add_argument("--profile", action="store_true")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
rank, world, local = MG.init_comm()
import torch.distributed as dist
L = api.lib()
hf = B.host_fasp()
if rank == 0:
    A, b = MG._quiet(B.build_problem, a.size)
    n = A.shape[0]
else:
    A, n = None, a.size ** 3
    b = np.ones(n)
amg, it = B.amg_recipe(hf)
sh = MG.SharedHierarchy(hf, A, amg, rank, world)
if rank == 0:
    print("# %d GPUs, 7-pt %d^3, hierarchy: %s" % (world, a.size, sh.how), flush=True)
first = True
for agg in [int(v) for v in a.agg_list.split(",")]:
    for cfg in [c for c in a.opts.split(";") if c]:
        for kv in cfg.split(","):
            k, v = kv.split("=")
            api.check(L.fasp_cuda_set_option(k.encode(), float(v)))
        s = MG.DistSolver(sh.mgl, amg, agg_rows=agg)     # after the options: upload-time choices included
        bl = np.ascontiguousarray(b[s.row0:s.row1]); z = np.zeros(s.row1 - s.row0)
        for _ in range(3):
            st_, x_ = s.solve(bl, z.copy(), it)
        tt = []
        for _ in range(a.reps):
            MG.barrier()
            st_, x_ = s.solve(bl, z.copy(), it)
            tt.append(MG.allreduce_max(s.stat(2)))
        parts = [None] * world
        dist.all_gather_object(parts, (s.row0, np.array(x_, copy=True)))
        if rank == 0:
            full = np.empty(n)
            for p0, xp in parts: full[p0:p0 + xp.size] = xp
            rel = float(np.linalg.norm(b - A.to_scipy() @ full) / np.linalg.norm(b))
            print("world %d agg_rows %7d [%s]: solve %.3f ms (min %.3f max %.3f), iterations %d, true relres %.3e, launches/solve %d"
                  % (world, agg, cfg, float(np.mean(tt)), min(tt), max(tt), st_, rel, int(s.stat(3))), flush=True)
        if a.profile and first:
            first = False
            L.fasp_cuda_set_option(b"profile", 1.0); L.fasp_cuda_profile_dump(None, 0)
            MG.barrier()
            st, _ = s.solve(bl, z.copy(), it)
            buf = C.create_string_buffer(64 << 20); L.fasp_cuda_profile_dump(buf, len(buf)); L.fasp_cuda_set_option(b"profile", 0.0)
            recs = [ln.split() for ln in buf.value.decode().splitlines()]
            recs = [(int(k), int(r), int(z_), float(ms)) for k, r, z_, ms, by in recs]
            if rank == 0:
                print("  profiled (no graph, serial) solve %.3f ms" % s.stat(2))
                grp = {}
                for k, r, z_, ms in recs:
                    key = ("halo" if k == 400 else "allreduce" if k == 401 else "allgather" if k == 402 else "dense" if k == 100 else "matrix", r if k < 400 else 0, z_ if k < 400 else 0)
                    e = grp.setdefault(key, [0, 0.0]); e[0] += 1; e[1] += ms
                for key, (c, ms) in sorted(grp.items(), key=lambda kv: -kv[1][1])[:16]:
                    print("    %-10s rows %9d nnz %10d  launches %4d  %8.3f ms  avg %7.1f us" % (key[0], key[1], key[2], c, ms, ms / c * 1e3))
        s.close()
        MG.barrier()
sh.close()
L.fasp_cuda_comm_finalize()
