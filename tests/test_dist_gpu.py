"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the row-partitioned solve on 2 ranks gives
the same iteration count and solution as the single-GPU solve and the sequential reference."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch.distributed as dist
from faspsolver_b200 import api, problems as PB, multigpu as MG, fasp_types as T
import bench as B
rank, world, local = MG.init_comm()
L = api.lib()
hf = B.host_fasp()
A = PB.poisson7(40); b = np.ones(A.shape[0])
for smoother, solver_type in ((T.SMOOTHER_L1DIAG, T.SOLVER_CG), (T.SMOOTHER_JACOBI, T.SOLVER_VGMRES), (T.SMOOTHER_POLY, T.SOLVER_CG)):
    amg = hf.amg_param(print_level=0, smoother=smoother, relaxation=0.67 if smoother == T.SMOOTHER_JACOBI else 1.0)
    it = hf.its_param(itsolver_type=solver_type, tol=1e-8, maxit=200, print_level=0, restart=30)
    sh = MG.SharedHierarchy(hf, A if rank == 0 else None, amg, rank, world)   # setup on rank 0 only
    mgl = sh.mgl
    s = MG.DistSolver(mgl, amg, agg_rows=2000)
    nloc = s.row1 - s.row0
    b_loc = np.ascontiguousarray(b[s.row0:s.row1])
    st, x_loc = s.solve(b_loc, np.zeros(nloc), it)
    st_again, x_again = s.solve(b_loc, np.zeros(nloc), it)        # cached workspace / graphs
    assert st_again == st and np.array_equal(x_again, x_loc), (rank, st, st_again)
    # device-pointer form with caller vectors of EXACTLY n_local entries (no ghost room, not peer-mapped):
    # the solver must stage x itself instead of exchanging ghosts into the caller's allocation
    d_b, d_x = L.fasp_cuda_dvec_alloc(nloc), L.fasp_cuda_dvec_alloc(nloc)
    api.check(L.fasp_cuda_dvec_h2d(d_b, T.as_preal(b_loc), nloc))
    api.check(L.fasp_cuda_dvec_h2d(d_x, T.as_preal(np.zeros(nloc)), nloc))
    st_dev = s.solve_dev(d_b, d_x, it)
    x_dev = np.empty(nloc)
    api.check(L.fasp_cuda_dvec_d2h(T.as_preal(x_dev), d_x, nloc))
    L.fasp_cuda_dvec_free(d_b); L.fasp_cuda_dvec_free(d_x)
    assert st_dev == st and np.array_equal(x_dev, x_loc), (rank, st, st_dev, np.abs(x_dev - x_loc).max())
    # the exchange-saving variants: redundant ghost rows off (bit-identical), overlap off (same to rounding)
    for opts in ((("ghost_redundant", 0.0),), (("overlap", 0.0),), (("ghost_redundant", 0.0), ("overlap", 0.0))):
        for k, v in opts:
            api.check(L.fasp_cuda_set_option(k.encode(), v))
        s2 = MG.DistSolver(mgl, amg, agg_rows=2000)
        st2, x2 = s2.solve(b_loc, np.zeros(nloc), it)
        s2.close()
        for k, v in opts:
            api.check(L.fasp_cuda_set_option(k.encode(), 1.0))
        # Same iterates up to rounding: the appended ghost rows change the row blocks of P / R (a row near a block
        # boundary may be summed by a lane group instead of one thread), and without the interior / boundary split
        # the fused dot products are summed in one piece instead of two
        assert st2 == st and np.abs(x2 - x_loc).max() <= 1e-10 * np.abs(x_loc).max(), (rank, opts, st, st2)
    parts = [None] * world
    dist.all_gather_object(parts, (s.row0, x_loc))
    s.close()
    if rank == 0:
        x = np.concatenate([p[1] for p in sorted(parts, key=lambda t: t[0])])
        from oracle.ref import RefFasp
        ref = RefFasp()
        amg_r = ref.amg_param(print_level=0, smoother=smoother, relaxation=0.67 if smoother == T.SMOOTHER_JACOBI else 1.0)
        st_ref, x_ref = ref.krylov_amg(A, b, np.zeros_like(b), it, amg_r)
        # the one-GPU solve on the same hierarchy, while the communicator is active
        s1 = api.KrylovAmgSolver(mgl, amg)
        st1, x1 = s1.solve(b, np.zeros_like(b), it)
        s1.close()
        rel = float(np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b))
        dx = float(np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref))
        dx1 = float(np.linalg.norm(x - x1) / np.linalg.norm(x1))
        print("RESULT", json.dumps({"smoother": smoother, "st": st, "st_ref": st_ref, "st1": st1, "rel": rel, "dx": dx, "dx1": dx1}))
    MG.barrier()
    sh.close()
L.fasp_cuda_comm_finalize()
'''


SLAB_WORKER = r'''
import os, sys, json, ctypes as C
sys.path.insert(0, %(root)r)
import numpy as np, torch.distributed as dist
from faspsolver_b200 import api, problems as PB, multigpu as MG, fasp_types as T, slabsetup as SS
from oracle.ref import RefFasp
rank, world, local = MG.init_comm()
L = api.lib()
ref = RefFasp()
comm = SS.HostComm(rank, world)
for name, gen, n, smoother in (("p27", PB.poisson27, 24, T.SMOOTHER_L1DIAG), ("p7", PB.poisson7, 40, T.SMOOTHER_JACOBI)):
    zoff = MG.plane_partition(n, world)
    off = [z * n * n for z in zoff]
    As = gen(n, zrange=(zoff[rank], zoff[rank + 1]))          # this rank's slab only
    relax = 0.67 if smoother == T.SMOOTHER_JACOBI else 1.0
    amg = ref.amg_param(print_level=0, smoother=smoother, relaxation=relax)
    it = ref.its_param(itsolver_type=T.SOLVER_CG, tol=1e-8, maxit=200, print_level=0)
    # p27: slab by slab on every level (seam rows re-interpolated, P and R with foreign columns); p7: the default,
    # levels that fit one FASP call are split as one piece
    sh = SS.SlabHierarchy(ref, As, off, amg, comm, agg_rows=1500, **({"max_piece_nnz": 0} if name == "p27" else {}))
    assert len(sh.levels) >= 2
    nloc = As.shape[0]
    b_loc = np.ones(nloc)
    res = {}
    for redundant in (1.0, 0.0):
        api.check(L.fasp_cuda_set_option(b"ghost_redundant", redundant))
        s = MG.SlabSolver(sh)
        assert (s.row0, s.row1) == (off[rank], off[rank + 1])
        st, x_loc = s.solve(b_loc, np.zeros(nloc), it)
        st2, x2 = s.solve(b_loc, np.zeros(nloc), it)
        assert st2 == st and np.array_equal(x2, x_loc)
        s.close()
        res[redundant] = (st, x_loc)
    api.check(L.fasp_cuda_set_option(b"ghost_redundant", 1.0))
    # redundant ghost rows: the same iterates up to rounding (the row blocks of P / R differ)
    assert res[1.0][0] == res[0.0][0] and np.abs(res[1.0][1] - res[0.0][1]).max() <= 1e-10 * np.abs(res[0.0][1]).max()
    st, x_loc = res[1.0]
    parts = [None] * world
    dist.all_gather_object(parts, (off[rank], x_loc))
    # oracle: the reference's PCG + fasp_precond_amg on the SAME hierarchy, assembled into global matrices;
    # and the one-GPU device solve on it
    mgl, amg_g = SS.assemble_mgl(ref, sh)
    if rank == 0:
        A = gen(n)
        N = A.shape[0]
        b = np.ones(N)
        x = np.concatenate([p[1] for p in sorted(parts, key=lambda t: t[0])])
        pcdata = T.precond_data()
        ref.L.fasp_param_amg_to_prec(C.byref(pcdata), C.byref(amg_g))
        pcdata.max_levels = mgl[0].num_levels
        pcdata.mgl_data = mgl
        pc = T.precond(C.cast(C.byref(pcdata), C.c_void_p), C.cast(ref.L.fasp_precond_amg, T.PRECOND_FCT))
        vb, vx = T.Vec(b), T.Vec(np.zeros(N))
        st_ref = ref.L.fasp_solver_dcsr_pcg(A.ptr(), vb.ptr(), vx.ptr(), C.byref(pc), it.tol, it.abstol, it.maxit, it.stop_type, 0)
        s1 = api.KrylovAmgSolver(mgl, amg_g)
        st1, x1 = s1.solve(b, np.zeros(N), it)
        s1.close()
        rel = float(np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b))
        print("RESULT", json.dumps({"name": name, "st": st, "st_ref": st_ref, "st1": st1, "rel": rel,
                                    "dx": float(np.linalg.norm(x - vx.a) / np.linalg.norm(vx.a)),
                                    "dx1": float(np.linalg.norm(x - x1) / np.linalg.norm(x1))}))
    MG.barrier()
    sh.close()
L.fasp_cuda_comm_finalize()
'''


def _ngpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return len([l for l in out.splitlines() if l.startswith("GPU ")])
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
def test_two_rank_solve_matches_reference(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": str(ROOT)})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = [json.loads(l.split("RESULT", 1)[1]) for l in r.stdout.splitlines() if "RESULT" in l]
    assert len(res) == 3
    for d in res:
        assert d["st"] > 0 and abs(d["st"] - d["st_ref"]) <= 1, d
        assert d["rel"] <= 1e-8 * 1.001 and d["dx"] <= 1e-8, d
        assert abs(d["st"] - d["st1"]) <= 1 and d["dx1"] <= 1e-8, d


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
def test_two_rank_slab_hierarchy_matches_reference_on_the_same_hierarchy(tmp_path):
    """No rank holds the global matrix: slabs are generated per rank, the hierarchy is built slab by slab
    (slabsetup.py), uploaded through fasp_cuda_dist_krylov_amg_create_slabs. Oracle: the reference's CPU PCG +
    fasp_precond_amg on the same hierarchy assembled into global matrices; and the one-GPU device solve."""
    script = tmp_path / "worker.py"
    script.write_text(SLAB_WORKER % {"root": str(ROOT)})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29534", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = [json.loads(l.split("RESULT", 1)[1]) for l in r.stdout.splitlines() if "RESULT" in l]
    assert len(res) == 2
    for d in res:
        assert d["st"] > 0 and abs(d["st"] - d["st_ref"]) <= 1 and abs(d["st"] - d["st1"]) <= 1, d
        assert d["rel"] <= 1e-8 * 1.001 and d["dx"] <= 1e-8 and d["dx1"] <= 1e-8, d
