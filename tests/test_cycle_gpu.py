"""One multigrid cycle and full AMG-Krylov solves vs the sequential reference on the same
hierarchy (built by the reference's own host setup)."""
import ctypes as C

import numpy as np
import pytest

from faspsolver_b200 import api
from faspsolver_b200 import fasp_types as T
from faspsolver_b200 import problems as PB

pytestmark = pytest.mark.gpu


def _cycle_both(gpu, ref, A, b, amg_kw):
    """Run fasp_solver_mgcycle (reference) and fasp_cuda_solver_mgcycle on the same mgl."""
    amg = ref.amg_param(print_level=0, **amg_kw)
    mgl = ref.amg_setup(A, amg)
    try:
        n = A.shape[0]
        nl = mgl[0].num_levels
        bv = np.ctypeslib.as_array(mgl[0].b.val, shape=(n,))
        xv = np.ctypeslib.as_array(mgl[0].x.val, shape=(n,))
        bv[:] = b
        xv[:] = 0.0
        ref.L.fasp_solver_mgcycle(mgl, C.byref(amg))
        x_ref = xv.copy()
        bv[:] = b
        xv[:] = 0.0
        st = gpu.fasp_cuda_solver_mgcycle(mgl, C.byref(amg))
        assert st == 0, gpu.fasp_cuda_last_error()
        x_gpu = xv.copy()
        # second cycle from a non-zero guess
        ref.L.fasp_solver_mgcycle(mgl, C.byref(amg))
        x_ref2 = xv.copy()
        xv[:] = x_gpu
        assert gpu.fasp_cuda_solver_mgcycle(mgl, C.byref(amg)) == 0
        x_gpu2 = xv.copy()
    finally:
        ref.amg_free(mgl, amg)
    return nl, x_ref, x_gpu, x_ref2, x_gpu2


CYCLE_CASES = [
    ("FE", dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67)),
    ("FE", dict(smoother=T.SMOOTHER_L1DIAG)),
    ("FE", dict(smoother=T.SMOOTHER_POLY, polynomial_degree=3)),
    ("FE", dict(smoother=T.SMOOTHER_L1DIAG, cycle_type=T.W_CYCLE)),
    ("FE", dict(smoother=T.SMOOTHER_L1DIAG, cycle_type=T.VW_CYCLE)),
    ("FE", dict(smoother=T.SMOOTHER_L1DIAG, cycle_type=T.WV_CYCLE)),
    ("FE", dict(smoother=T.SMOOTHER_L1DIAG, presmooth_iter=2, postsmooth_iter=3)),
    ("FE", dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67, coarse_scaling=T.ON)),
    ("FE", dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67, AMG_type=T.SA_AMG)),
    ("FE", dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67, AMG_type=T.UA_AMG)),
    ("FD", dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67, coarse_dof=20)),
    ("FD", dict(smoother=T.SMOOTHER_L1DIAG)),   # single level: the cycle is the coarse solve
    ("p7", dict(smoother=T.SMOOTHER_L1DIAG)),
    ("cd7", dict(smoother=T.SMOOTHER_POLY, polynomial_degree=3)),
]


@pytest.mark.parametrize("prob,amg_kw", CYCLE_CASES)
def test_mgcycle_matches_reference(gpu, ref, data, prob, amg_kw):
    if prob in ("FD", "FE"):
        A, b = data[prob], data[prob + "_b"]
    elif prob == "p7":
        A = PB.poisson7(20)
        b = np.ones(A.shape[0])
    else:
        A = PB.convdiff7(20)
        b = np.ones(A.shape[0])
    nl, x_ref, x_gpu, x_ref2, x_gpu2 = _cycle_both(gpu, ref, A, b, amg_kw)
    # the reference's coarsest solve is CG to 1e-10 (PreMGUtil.inl:43), ours is direct
    for xr, xg in ((x_ref, x_gpu), (x_ref2, x_gpu2)):
        rel = np.linalg.norm(xg - xr) / np.linalg.norm(xr)
        assert rel < 1e-8, (prob, amg_kw, nl, rel)


SOLVE_CASES = [
    ("FD", dict(itsolver_type=T.SOLVER_CG), dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67, coarse_dof=20)),
    ("FD", dict(itsolver_type=T.SOLVER_CG), dict(smoother=T.SMOOTHER_L1DIAG, coarse_dof=20)),
    ("FD", dict(itsolver_type=T.SOLVER_CG), dict(smoother=T.SMOOTHER_L1DIAG)),
    ("FE", dict(itsolver_type=T.SOLVER_CG), dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67)),
    ("FE", dict(itsolver_type=T.SOLVER_CG), dict(smoother=T.SMOOTHER_L1DIAG)),
    ("FE", dict(itsolver_type=T.SOLVER_CG), dict(smoother=T.SMOOTHER_POLY, polynomial_degree=3)),
    ("FE", dict(itsolver_type=T.SOLVER_CG), dict(smoother=T.SMOOTHER_L1DIAG, cycle_type=T.W_CYCLE)),
    ("FE", dict(itsolver_type=T.SOLVER_GMRES, restart=30), dict(smoother=T.SMOOTHER_L1DIAG)),
    ("FE", dict(itsolver_type=T.SOLVER_VGMRES, restart=30), dict(smoother=T.SMOOTHER_POLY, polynomial_degree=3)),
    ("FE", dict(itsolver_type=T.SOLVER_GMRES, restart=3), dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67)),
    ("FE", dict(itsolver_type=T.SOLVER_VGMRES, restart=4), dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67)),
    ("FE", dict(itsolver_type=T.SOLVER_VFGMRES, restart=30), dict(smoother=T.SMOOTHER_L1DIAG)),
    ("FE", dict(itsolver_type=T.SOLVER_VFGMRES, restart=4), dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67)),
    ("cd7", dict(itsolver_type=T.SOLVER_VFGMRES, restart=30), dict(smoother=T.SMOOTHER_POLY, polynomial_degree=3)),
    ("FE", dict(itsolver_type=T.SOLVER_CG), dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67, AMG_type=T.SA_AMG)),
    ("FE", dict(itsolver_type=T.SOLVER_CG), dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67, AMG_type=T.UA_AMG)),
    ("p7", dict(itsolver_type=T.SOLVER_CG), dict(smoother=T.SMOOTHER_L1DIAG)),
    ("cd7", dict(itsolver_type=T.SOLVER_GMRES, restart=30), dict(smoother=T.SMOOTHER_POLY, polynomial_degree=3)),
    ("cd7", dict(itsolver_type=T.SOLVER_VGMRES, restart=30), dict(smoother=T.SMOOTHER_POLY, polynomial_degree=3)),
]


@pytest.mark.parametrize("prob,it_kw,amg_kw", SOLVE_CASES)
def test_krylov_amg_matches_reference(gpu, ref, data, prob, it_kw, amg_kw):
    """Drop-in fasp_cuda_solver_dcsr_krylov_amg vs fasp_solver_dcsr_krylov_amg: iteration
    count +-1, true relative residual <= tol, solution within 1e-8 relative."""
    if prob in ("FD", "FE"):
        A, b = data[prob], data[prob + "_b"]
    elif prob == "p7":
        A = PB.poisson7(24)
        b = np.ones(A.shape[0])
    else:
        A = PB.convdiff7(24)
        b = np.ones(A.shape[0])
    tol = 1e-8
    it = ref.its_param(tol=tol, maxit=500, print_level=0, **it_kw)
    amg = ref.amg_param(print_level=0, **amg_kw)
    st_ref, x_ref = ref.krylov_amg(A, b, np.zeros_like(b), it, amg)
    amg2 = ref.amg_param(print_level=0, **amg_kw)
    st, x = api.fasp_cuda_solver_dcsr_krylov_amg(A, b, np.zeros_like(b), it, amg2)
    assert st >= 0, (st, api.last_error())
    assert abs(st - st_ref) <= 1, (st, st_ref)
    r = b - A.to_scipy() @ x
    assert np.linalg.norm(r) / np.linalg.norm(b) <= tol * 1.0000001
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) <= 1e-8, np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)


def test_golden_iteration_counts(gpu, ref, data, golden_answers):
    """The committed oracle answers (tests/golden/oracle_answers.json) for the recipes above."""
    by_name = {r["name"]: r for r in golden_answers["recipes"]}
    for name, prob in (("FE_pcg_jacobi067", "FE"), ("FE_pcg_l1", "FE"), ("FE_pcg_poly3", "FE"),
                       ("FE_gmres30_l1", "FE"), ("FE_vfgmres30_l1", "FE"), ("FD_pcg_l1_cdof20", "FD")):
        rec = by_name[name]
        A, b = data[prob], data[prob + "_b"]
        it = ref.its_param(tol=1e-8, maxit=500, print_level=0, **rec["it"])
        amg = ref.amg_param(print_level=0, **rec["amg"])
        st, x = api.fasp_cuda_solver_dcsr_krylov_amg(A, b, np.zeros_like(b), it, amg)
        assert abs(st - rec["status"]) <= 1, (name, st, rec["status"])


def test_reg_gcc_golden_l1diag_amg_solver(gpu, ref, data, golden_answers):
    """test/out/reg.gcc:412 — 'Classical AMG V-cycle with L1_DIAG smoother as iterative
    solver' on the FE problem: 19 iterations, relres 8.612004e-11 (regression.c:275-289)."""
    g = golden_answers["reg_gcc"]["FE_amg_solver_L1DIAG_tol1e-10"]
    A, b = data["FE"], data["FE_b"]
    amg = ref.amg_param(print_level=0, maxit=500, tol=1e-10, smoother=T.SMOOTHER_L1DIAG)
    mgl = ref.amg_setup(A, amg)
    try:
        n = A.shape[0]
        np.ctypeslib.as_array(mgl[0].b.val, shape=(n,))[:] = b
        np.ctypeslib.as_array(mgl[0].x.val, shape=(n,))[:] = 0.0
        st = gpu.fasp_cuda_amg_solve(mgl, C.byref(amg))
        x = np.ctypeslib.as_array(mgl[0].x.val, shape=(n,)).copy()
    finally:
        ref.amg_free(mgl, amg)
    assert st == g["iters"], st
    relres = np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b)
    assert abs(relres - g["relres"]) / g["relres"] < 1e-3, relres
    assert np.abs(x - data["FE_sol"]).max() < 1e-4   # regression.c:56 tolerance


def test_pcg_with_null_precond(gpu, ref, data):
    """fasp_cuda_solver_dcsr_pcg with pc == NULL (identity, KryPcg.c:128-131). The host-callback and
    device-callback forms of the plug-in contract are in tests/test_plugins_gpu.py."""
    A, b = data["FE"], data["FE_b"]
    n = A.shape[0]
    vb, vx, vxr = T.Vec(b), T.Vec(np.zeros(n)), T.Vec(np.zeros(n))
    st = gpu.fasp_cuda_solver_dcsr_pcg(A.ptr(), vb.ptr(), vx.ptr(), None, 1e-8, 1e-18, 1000, 1, 0)
    st_ref = ref.L.fasp_solver_dcsr_pcg(A.ptr(), vb.ptr(), vxr.ptr(), None, 1e-8, 1e-18, 1000, 1, 0)
    assert st > 0 and abs(st - st_ref) <= 1, (st, st_ref)
    assert np.linalg.norm(vx.a - vxr.a) / np.linalg.norm(vxr.a) < 1e-7
