#!/usr/bin/env python
"""Repeat-solve reproducibility of the multi-GPU solver (torchrun, 2+ ranks): the same solve four times on one solver
object, for every recipe and option set; prints iteration counts and max |x_k - x_0| / max |x_0|."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import bench as B
from faspsolver_b200 import api, problems as PB, multigpu as MG, fasp_types as T
rank, world, local = MG.init_comm()
L = api.lib()
hf = B.host_fasp()
A = PB.poisson7(40); b = np.ones(A.shape[0])
recipes = (("L1+CG", T.SMOOTHER_L1DIAG, T.SOLVER_CG), ("Jacobi+VGMRES", T.SMOOTHER_JACOBI, T.SOLVER_VGMRES),
           ("poly+CG", T.SMOOTHER_POLY, T.SOLVER_CG), ("poly+VGMRES", T.SMOOTHER_POLY, T.SOLVER_VGMRES))
optsets = ({}, {"ghost_redundant": 2.0}, {"overlap": 0.0}, {"ghost_redundant": 0.0})
for name, smoother, solver_type in recipes:
    amg = hf.amg_param(print_level=0, smoother=smoother, relaxation=0.67 if smoother == T.SMOOTHER_JACOBI else 1.0)
    it = hf.its_param(itsolver_type=solver_type, tol=1e-8, maxit=200, print_level=0, restart=30)
    sh = MG.SharedHierarchy(hf, A if rank == 0 else None, amg, rank, world)
    for opts in optsets:
        for k, v in opts.items():
            api.check(L.fasp_cuda_set_option(k.encode(), v))
        s = MG.DistSolver(sh.mgl, amg, agg_rows=2000)
        nloc = s.row1 - s.row0
        b_loc = np.ascontiguousarray(b[s.row0:s.row1])
        xs, sts = [], []
        for k in range(4):
            st, x = s.solve(b_loc, np.zeros(nloc), it)
            xs.append(x.copy()); sts.append(st)
        d = [float(np.abs(x - xs[0]).max() / np.abs(xs[0]).max()) for x in xs[1:]]
        dm = MG.allreduce_max(max(d))
        if rank == 0:
            print("%-14s %-42s iters %s  max rel diff of repeats %.3e" % (name, opts, sts, dm), flush=True)
        s.close()
        for k in opts:
            api.check(L.fasp_cuda_set_option(k.encode(), 1.0 if k != "lookahead" else 2.0))
    MG.barrier()
    sh.close()
L.fasp_cuda_comm_finalize()
