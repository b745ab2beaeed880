"""The plug-in side of the boundary: the reference's OWN Krylov loops driving the device
through its two callback interfaces, and the device Krylov loops driven with host callbacks.

  plug-in #1  precond.fct      (fasp.h:1095-1103): fasp_solver_dcsr_pcg (KryPcg.c:96) with
              pc->fct = fasp_cuda_precond_amg from fasp_cuda_precond_setup / _from_mgl
  plug-in #2  mxv_matfree.fct  (fasp.h:1109-1117): fasp_solver_pcg / fasp_solver_pvgmres
              (KryPcg.c:1260, KryPvgmres.c:1468) with mf.fct = fasp_cuda_blas_mxv_csr / _bsr
  HostPrec    fasp_cuda_solver_dcsr_pcg / _dbsr_pcg / _dbsr_pgmres with a host pc->fct
              (the reference's fasp_precond_amg, fasp_precond_diag, fasp_precond_dbsr_diag)
"""
import ctypes as C

import numpy as np
import pytest

from faspsolver_b200 import api
from faspsolver_b200 import fasp_types as T
from faspsolver_b200 import problems as PB

pytestmark = pytest.mark.gpu
P = C.POINTER


def _bind_ref_extras(ref):
    L = ref.L
    L.fasp_solver_pcg.restype = T.INT
    L.fasp_solver_pcg.argtypes = [P(T.mxv_matfree), P(T.dvector), P(T.dvector), P(T.precond), T.REAL, T.REAL, T.INT,
                                  T.SHORT, T.SHORT]
    L.fasp_solver_pvgmres.restype = T.INT
    L.fasp_solver_pvgmres.argtypes = [P(T.mxv_matfree), P(T.dvector), P(T.dvector), P(T.precond), T.REAL, T.REAL,
                                      T.INT, T.SHORT, T.SHORT, T.SHORT]
    L.fasp_solver_matfree_init.restype = None
    L.fasp_solver_matfree_init.argtypes = [T.INT, P(T.mxv_matfree), C.c_void_p]
    L.fasp_precond_diag.restype = None
    L.fasp_precond_dbsr_diag.restype = None
    return L


def _ref_amg_precond(ref, mgl, amg):
    """precond {data = precond_data, fct = fasp_precond_amg} exactly as SolCSR.c:525-538 builds it."""
    pcdata = T.precond_data()
    ref.L.fasp_param_amg_to_prec(C.byref(pcdata), C.byref(amg))
    pcdata.max_levels = mgl[0].num_levels
    pcdata.mgl_data = mgl
    pc = T.precond(C.cast(C.byref(pcdata), C.c_void_p), C.cast(ref.L.fasp_precond_amg, T.PRECOND_FCT))
    pc._keep = pcdata
    return pc


@pytest.mark.parametrize("how", ["setup", "from_mgl"])
def test_reference_pcg_with_device_precond_callback(gpu, ref, data, how):
    """plug-in #1: the reference's fasp_solver_dcsr_pcg calls fasp_cuda_precond_amg with HOST r, z."""
    A, b = data["FE"], data["FE_b"]
    n = A.shape[0]
    amg = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
    mgl = None
    if how == "setup":
        pc = gpu.fasp_cuda_precond_setup(T.PREC_AMG, C.byref(amg), None, A.ptr())
    else:
        mgl = ref.amg_setup(A, amg)
        pc = gpu.fasp_cuda_precond_from_mgl(mgl, C.byref(amg))
    assert pc, gpu.fasp_cuda_last_error()
    try:
        vb, vx = T.Vec(b), T.Vec(np.zeros(n))
        st = ref.L.fasp_solver_dcsr_pcg(A.ptr(), vb.ptr(), vx.ptr(), pc, 1e-8, 1e-18, 500, 1, 0)
        # the callback on its own: z = B r equals the device cycle through the handle API
        r = np.random.default_rng(3).uniform(-1, 1, n)
        z = np.zeros(n)
        gpu.fasp_cuda_precond_amg(T.as_preal(r), T.as_preal(z), pc.contents.data)
        # the same precond object keeps the device Krylov loop resident (fast path)
        vx2 = T.Vec(np.zeros(n))
        st_dev = gpu.fasp_cuda_solver_dcsr_pcg(A.ptr(), vb.ptr(), vx2.ptr(), pc, 1e-8, 1e-18, 500, 1, 0)
    finally:
        gpu.fasp_cuda_precond_free(pc)
        if mgl is not None:
            ref.amg_free(mgl, amg)
    it = ref.its_param(itsolver_type=T.SOLVER_CG, tol=1e-8, maxit=500, print_level=0)
    amg_r = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
    st_ref, x_ref = ref.krylov_amg(A, b, np.zeros(n), it, amg_r)
    assert st > 0 and abs(st - st_ref) <= 1, (st, st_ref)
    assert st_dev > 0 and abs(st_dev - st_ref) <= 1, (st_dev, st_ref)
    assert np.linalg.norm(vx.a - x_ref) / np.linalg.norm(x_ref) <= 1e-8
    assert np.linalg.norm(vx2.a - x_ref) / np.linalg.norm(x_ref) <= 1e-8
    assert np.linalg.norm(b - A.to_scipy() @ vx.a) / np.linalg.norm(b) <= 1e-8 * 1.0000001
    # z = B r against the reference's own cycle on the reference's hierarchy
    amg_c = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
    mgl_c = ref.amg_setup(A, amg_c)
    try:
        pcr = _ref_amg_precond(ref, mgl_c, amg_c)
        z_ref = np.zeros(n)
        ref.L.fasp_precond_amg(T.as_preal(r.copy()), T.as_preal(z_ref), pcr.data)
        assert np.linalg.norm(z - z_ref) / np.linalg.norm(z_ref) < 1e-8
        # ... and through the handle API with host vectors
        h = gpu.fasp_cuda_amg_upload(mgl_c, C.byref(amg_c))
        assert h, gpu.fasp_cuda_last_error()
        z2 = np.zeros(n)
        assert gpu.fasp_cuda_amg_cycle_host(h, T.as_preal(r), T.as_preal(z2)) == 0, gpu.fasp_cuda_last_error()
        assert gpu.fasp_cuda_amg_levels(h) == mgl_c[0].num_levels and gpu.fasp_cuda_amg_bytes(h) > 0
        gpu.fasp_cuda_amg_free(h)
        assert np.linalg.norm(z2 - z_ref) / np.linalg.norm(z_ref) < 1e-8
    finally:
        ref.amg_free(mgl_c, amg_c)


def test_reference_matfree_krylov_with_device_spmv_csr(gpu, ref, data):
    """plug-in #2, CSR: fasp_solver_pcg (KryPcg.c:1260) with mf.fct = fasp_cuda_blas_mxv_csr. The FE rows are
    short, so the device SpMV is bit-identical and the whole CG run must be bit-identical too."""
    L = _bind_ref_extras(ref)
    A, b = data["FE"], data["FE_b"]
    n = A.shape[0]
    mf, mf_ref = T.mxv_matfree(), T.mxv_matfree()
    assert gpu.fasp_cuda_solver_matfree_init(T.MAT_CSR, C.byref(mf), C.cast(A.ptr(), C.c_void_p)) == 0
    L.fasp_solver_matfree_init(T.MAT_CSR, C.byref(mf_ref), C.cast(A.ptr(), C.c_void_p))
    vb, vx, vxr = T.Vec(b), T.Vec(np.zeros(n)), T.Vec(np.zeros(n))
    st = L.fasp_solver_pcg(C.byref(mf), vb.ptr(), vx.ptr(), None, 1e-8, 1e-18, 1000, 1, 0)
    st_ref = L.fasp_solver_pcg(C.byref(mf_ref), vb.ptr(), vxr.ptr(), None, 1e-8, 1e-18, 1000, 1, 0)
    assert st == st_ref and st > 0, (st, st_ref)
    assert np.array_equal(vx.a, vxr.a)
    assert gpu.fasp_cuda_solver_matfree_init(3, C.byref(mf), None) == T.ERROR_DATA_STRUCTURE   # MAT_STR: not ours


def test_reference_matfree_krylov_with_device_spmv_bsr(gpu, ref, data):
    """plug-in #2, BSR: fasp_solver_pvgmres (KryPvgmres.c:1468) with mf.fct = fasp_cuda_blas_mxv_bsr on SPE01."""
    L = _bind_ref_extras(ref)
    A, b = data["SPE"], data["SPE_b"]
    n = b.size
    mf, mf_ref = T.mxv_matfree(), T.mxv_matfree()
    assert gpu.fasp_cuda_solver_matfree_init(T.MAT_BSR, C.byref(mf), C.cast(A.ptr(), C.c_void_p)) == 0
    L.fasp_solver_matfree_init(T.MAT_BSR, C.byref(mf_ref), C.cast(A.ptr(), C.c_void_p))
    vb, vx, vxr = T.Vec(b), T.Vec(np.zeros(n)), T.Vec(np.zeros(n))
    st = L.fasp_solver_pvgmres(C.byref(mf), vb.ptr(), vx.ptr(), None, 1e-6, 1e-18, 200, 30, 1, 0)
    st_ref = L.fasp_solver_pvgmres(C.byref(mf_ref), vb.ptr(), vxr.ptr(), None, 1e-6, 1e-18, 200, 30, 1, 0)
    assert st == st_ref, (st, st_ref)
    assert np.array_equal(vx.a, vxr.a)   # the BSR kernel is bit-identical to the CPU block loops


def test_device_pcg_with_host_precond_callbacks(gpu, ref, data):
    """HostPrec: fasp_cuda_solver_dcsr_pcg with the reference's own host preconditioners as pc->fct —
    fasp_precond_amg on the reference's hierarchy (D2H r, CPU V-cycle, H2D z per iteration) and fasp_precond_diag."""
    _bind_ref_extras(ref)
    A, b = data["FE"], data["FE_b"]
    n = A.shape[0]
    amg = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
    mgl = ref.amg_setup(A, amg)
    try:
        pc = _ref_amg_precond(ref, mgl, amg)
        vb, vx, vxr = T.Vec(b), T.Vec(np.zeros(n)), T.Vec(np.zeros(n))
        st = gpu.fasp_cuda_solver_dcsr_pcg(A.ptr(), vb.ptr(), vx.ptr(), C.byref(pc), 1e-8, 1e-18, 500, 1, 0)
        st_ref = ref.L.fasp_solver_dcsr_pcg(A.ptr(), vb.ptr(), vxr.ptr(), C.byref(pc), 1e-8, 1e-18, 500, 1, 0)
    finally:
        ref.amg_free(mgl, amg)
    assert st > 0 and abs(st - st_ref) <= 1, (st, st_ref, gpu.fasp_cuda_last_error())
    assert np.linalg.norm(vx.a - vxr.a) / np.linalg.norm(vxr.a) <= 1e-8
    # diagonal preconditioner, data = dvector* of the diagonal (PreCSR.c:172)
    S = A.to_scipy()
    diag = T.Vec(S.diagonal().copy())
    pcd = T.precond(C.cast(diag.ptr(), C.c_void_p), C.cast(ref.L.fasp_precond_diag, T.PRECOND_FCT))
    vx, vxr = T.Vec(np.zeros(n)), T.Vec(np.zeros(n))
    st = gpu.fasp_cuda_solver_dcsr_pcg(A.ptr(), vb.ptr(), vx.ptr(), C.byref(pcd), 1e-8, 1e-18, 1000, 1, 0)
    st_ref = ref.L.fasp_solver_dcsr_pcg(A.ptr(), vb.ptr(), vxr.ptr(), C.byref(pcd), 1e-8, 1e-18, 1000, 1, 0)
    assert st > 0 and abs(st - st_ref) <= 1, (st, st_ref)
    assert np.linalg.norm(vx.a - vxr.a) / np.linalg.norm(vxr.a) <= 1e-7


class precond_diag_bsr(C.Structure):
    """fasp_block.h:255-263"""
    _fields_ = [("nb", T.INT), ("diag", T.dvector)]


def test_device_bsr_pcg_and_pgmres(gpu, ref, data):
    """fasp_cuda_solver_dbsr_pcg / _dbsr_pgmres (KryPcg.c:386, KryPgmres.c:376): pc == NULL and the reference's
    block-diagonal host preconditioner fasp_precond_dbsr_diag (SolBSR.c:215-262) as a callback."""
    _bind_ref_extras(ref)
    # SPD block system: kron(7-point Laplacian, SPD 3x3 block) so that CG applies
    s = PB.poisson7(8, scaled=False)
    Bk = np.array([[2.0, -0.5, 0.1], [-0.5, 2.0, -0.3], [0.1, -0.3, 1.5]])
    val = s.val[:, None, None] * Bk[None, :, :]
    A = T.BSR(s.shape[0], s.shape[1], 3, s.ia, s.ja, val)
    n = A.ROW * A.nb
    b = 1.0 + 0.01 * (np.arange(n) % 5)
    vb = T.Vec(b)
    dinv = ref.L.fasp_dbsr_getdiaginv(A.ptr())
    pdata = precond_diag_bsr(A.nb, dinv)
    pc = T.precond(C.cast(C.byref(pdata), C.c_void_p), C.cast(ref.L.fasp_precond_dbsr_diag, T.PRECOND_FCT))
    for pcp in (None, C.byref(pc)):
        vx, vxr = T.Vec(np.zeros(n)), T.Vec(np.zeros(n))
        st = gpu.fasp_cuda_solver_dbsr_pcg(A.ptr(), vb.ptr(), vx.ptr(), pcp, 1e-8, 1e-18, 500, 1, 0)
        st_ref = ref.L.fasp_solver_dbsr_pcg(A.ptr(), vb.ptr(), vxr.ptr(), pcp, 1e-8, 1e-18, 500, 1, 0)
        assert st > 0 and abs(st - st_ref) <= 1, (st, st_ref, gpu.fasp_cuda_last_error())
        assert np.linalg.norm(vx.a - vxr.a) / np.linalg.norm(vxr.a) <= 1e-8
        vx, vxr = T.Vec(np.zeros(n)), T.Vec(np.zeros(n))
        st = gpu.fasp_cuda_solver_dbsr_pgmres(A.ptr(), vb.ptr(), vx.ptr(), pcp, 1e-8, 1e-18, 500, 20, 1, 0)
        st_ref = ref.L.fasp_solver_dbsr_pgmres(A.ptr(), vb.ptr(), vxr.ptr(), pcp, 1e-8, 1e-18, 500, 20, 1, 0)
        assert st > 0 and abs(st - st_ref) <= 1, (st, st_ref, gpu.fasp_cuda_last_error())
        assert np.linalg.norm(vx.a - vxr.a) / np.linalg.norm(vxr.a) <= 1e-7
    # the BSR driver with SOLVER_CG (SolBSR.c:55) on the same system
    it = ref.its_param(itsolver_type=T.SOLVER_CG, tol=1e-8, maxit=500, print_level=0)
    vx, vxr = T.Vec(np.zeros(n)), T.Vec(np.zeros(n))
    st = gpu.fasp_cuda_solver_dbsr_itsolver(A.ptr(), vb.ptr(), vx.ptr(), None, C.byref(it))
    st_ref = ref.L.fasp_solver_dbsr_pcg(A.ptr(), vb.ptr(), vxr.ptr(), None, 1e-8, it.abstol, 500, 1, 0)
    assert st > 0 and abs(st - st_ref) <= 1


@pytest.mark.parametrize("solver", ["pcg", "pgmres", "pvgmres"])
def test_stop_rel_precres(gpu, ref, data, solver):
    """STOP_REL_PRECRES (KryPcg.c:141-150,195-203; KryPvgmres.c:160-167,312-319): ||r||_B = sqrt((B r, r)) with
    a device-resident AMG preconditioner, against the reference with its own fasp_precond_amg."""
    A, b = data["FE"], data["FE_b"]
    n = A.shape[0]
    amg = ref.amg_param(print_level=0, smoother=T.SMOOTHER_JACOBI, relaxation=0.67)
    mgl = ref.amg_setup(A, amg)
    pc = gpu.fasp_cuda_precond_from_mgl(mgl, C.byref(amg))
    assert pc, gpu.fasp_cuda_last_error()
    try:
        pcr = _ref_amg_precond(ref, mgl, amg)
        vb, vx, vxr = T.Vec(b), T.Vec(np.zeros(n)), T.Vec(np.zeros(n))
        if solver == "pcg":
            st = gpu.fasp_cuda_solver_dcsr_pcg(A.ptr(), vb.ptr(), vx.ptr(), pc, 1e-8, 1e-18, 500, T.STOP_REL_PRECRES, 0)
            st_ref = ref.L.fasp_solver_dcsr_pcg(A.ptr(), vb.ptr(), vxr.ptr(), C.byref(pcr), 1e-8, 1e-18, 500,
                                                T.STOP_REL_PRECRES, 0)
        else:
            f = getattr(gpu, "fasp_cuda_solver_dcsr_" + solver)
            fr = getattr(ref.L, "fasp_solver_dcsr_" + solver)
            st = f(A.ptr(), vb.ptr(), vx.ptr(), pc, 1e-8, 1e-18, 500, 30, T.STOP_REL_PRECRES, 0)
            st_ref = fr(A.ptr(), vb.ptr(), vxr.ptr(), C.byref(pcr), 1e-8, 1e-18, 500, 30, T.STOP_REL_PRECRES, 0)
    finally:
        gpu.fasp_cuda_precond_free(pc)
        ref.amg_free(mgl, amg)
    assert st > 0 and abs(st - st_ref) <= 1, (solver, st, st_ref, gpu.fasp_cuda_last_error())
    assert np.linalg.norm(vx.a - vxr.a) / np.linalg.norm(vxr.a) <= 1e-7


def test_blas1_drop_ins(gpu, ref):
    """fasp_cuda_blas_darray_* against fasp_blas_darray_* (BlaArray.c): element-wise results bit for bit
    (including the a == 1 / a == -1 shortcuts), reductions to 1e-14 relative."""
    L = ref.L
    for nm, args in (("ax", [T.INT, T.REAL, T.PREAL]), ("axpy", [T.INT, T.REAL, T.PREAL, T.PREAL]),
                     ("axpby", [T.INT, T.REAL, T.PREAL, T.REAL, T.PREAL])):
        f = getattr(L, "fasp_blas_darray_" + nm)
        f.restype, f.argtypes = None, args
    for nm, args in (("dotprod", [T.INT, T.PREAL, T.PREAL]), ("norm2", [T.INT, T.PREAL]), ("norm1", [T.INT, T.PREAL]),
                     ("norminf", [T.INT, T.PREAL])):
        f = getattr(L, "fasp_blas_darray_" + nm)
        f.restype, f.argtypes = T.REAL, args
    rng = np.random.default_rng(11)
    for n in (1, 7, 1000, 300001):
        x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        for a in (1.0, -1.0, 0.37, 0.0):
            yg, yr = y.copy(), y.copy()
            assert gpu.fasp_cuda_blas_darray_axpy(n, a, T.as_preal(x), T.as_preal(yg)) == 0
            L.fasp_blas_darray_axpy(n, a, T.as_preal(x), T.as_preal(yr))
            assert np.array_equal(yg, yr), ("axpy", n, a)
            yg, yr = y.copy(), y.copy()
            assert gpu.fasp_cuda_blas_darray_axpby(n, a, T.as_preal(x), -0.6, T.as_preal(yg)) == 0
            L.fasp_blas_darray_axpby(n, a, T.as_preal(x), -0.6, T.as_preal(yr))
            assert np.array_equal(yg, yr), ("axpby", n, a)
            xg, xr = x.copy(), x.copy()
            assert gpu.fasp_cuda_blas_darray_ax(n, a, T.as_preal(xg)) == 0
            L.fasp_blas_darray_ax(n, a, T.as_preal(xr))
            assert np.array_equal(xg, xr), ("ax", n, a)
        import math
        exact = {"dotprod": math.fsum(x * y), "norm2": math.sqrt(math.fsum(x * x)), "norm1": math.fsum(np.abs(x)),
                 "norminf": float(np.abs(x).max())}
        for nm, fa in (("dotprod", (T.as_preal(x), T.as_preal(y))), ("norm2", (T.as_preal(x),)),
                       ("norm1", (T.as_preal(x),)), ("norminf", (T.as_preal(x),))):
            vg = getattr(gpu, "fasp_cuda_blas_darray_" + nm)(n, *fa)
            vr = getattr(L, "fasp_blas_darray_" + nm)(n, *fa)
            scale = float(np.abs(x * y).sum()) if nm == "dotprod" else abs(vr)
            # the device tree sum against the exactly rounded value: 1e-14 relative
            assert abs(vg - exact[nm]) <= 1e-14 * scale, (nm, n, vg, exact[nm])
            # against the CPU's left-to-right sum, which itself carries O(sqrt(n) eps) of rounding error
            assert abs(vg - vr) <= max(1e-14, 8 * 2.3e-16 * math.sqrt(n)) * scale, (nm, n, vg, vr)
    assert gpu.fasp_cuda_blas_darray_norminf(0, None) == 0.0


def test_dense_inverse_blocked_gauss_jordan(gpu):
    """The coarsest-level factorisation (dense.cu): sizes around the panel width 32 and the 64-wide update tiles,
    a matrix that needs row exchanges in every panel, and the singular-matrix report."""
    rng = np.random.default_rng(21)
    for n in (1, 2, 31, 32, 33, 64, 65, 100, 257, 700):
        M = rng.uniform(-1, 1, (n, n))
        if n > 3:
            M[np.arange(n), np.arange(n)] = 0.0   # zero diagonal: pivoting is mandatory
        X = np.zeros((n, n))
        st = gpu.fasp_cuda_dense_inverse(n, T.as_preal(np.ascontiguousarray(M)), T.as_preal(X))
        assert st == 0, (n, gpu.fasp_cuda_last_error())
        Xr = np.linalg.inv(M)
        err = np.abs(X @ M - np.eye(n)).max()
        assert err <= 1e-9 * max(1.0, np.linalg.cond(M) * 1e-4), (n, err)
        assert np.abs(X - Xr).max() <= 1e-8 * np.abs(Xr).max() * max(1.0, np.linalg.cond(M) * 1e-6), n
    S = np.ones((40, 40))
    X = np.zeros((40, 40))
    assert gpu.fasp_cuda_dense_inverse(40, T.as_preal(S), T.as_preal(X)) == T.ERROR_AMG_SETUP
    assert "singular" in gpu.fasp_cuda_last_error().decode()


def test_coarse_level_solvers(gpu, ref, data):
    """Coarsest level: (i) a large one (two-level hierarchy, 864 rows) through the dense inverse, (ii) the same
    through the iterative fallback (coarse_dense_max below its size: CG in one cooperative kernel, replacing
    fasp_coarse_itsolver PreMGUtil.inl:37), (iii) a singular coarsest operator (pure Neumann): the factorisation
    reports the vanishing pivot and the solver falls back to CG instead of applying an inf/NaN inverse."""
    A = PB.poisson7(12)
    b = np.ones(A.shape[0])
    it = ref.its_param(itsolver_type=T.SOLVER_CG, tol=1e-8, maxit=200, print_level=0)
    kw = dict(print_level=0, smoother=T.SMOOTHER_L1DIAG, max_levels=2)
    st_ref, x_ref = ref.krylov_amg(A, b, np.zeros_like(b), it, ref.amg_param(**kw))
    for dense_max in (8192, 100):
        assert gpu.fasp_cuda_set_option(b"coarse_dense_max", float(dense_max)) == 0
        try:
            st, x = api.fasp_cuda_solver_dcsr_krylov_amg(A, b, np.zeros_like(b), it, ref.amg_param(**kw))
        finally:
            gpu.fasp_cuda_set_option(b"coarse_dense_max", 8192.0)
        assert st > 0 and abs(st - st_ref) <= 1, (dense_max, st, st_ref, api.last_error())
        assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) <= 1e-8, dense_max
    # one cycle through the iterative coarse solve against the reference's cycle (same tolerance 1e-10)
    amg = ref.amg_param(**kw)
    mgl = ref.amg_setup(A, amg)
    try:
        n = A.shape[0]
        bv = np.ctypeslib.as_array(mgl[0].b.val, shape=(n,))
        xv = np.ctypeslib.as_array(mgl[0].x.val, shape=(n,))
        bv[:] = b
        xv[:] = 0.0
        ref.L.fasp_solver_mgcycle(mgl, C.byref(amg))
        xr = xv.copy()
        xv[:] = 0.0
        gpu.fasp_cuda_set_option(b"coarse_dense", 0.0)
        try:
            assert gpu.fasp_cuda_solver_mgcycle(mgl, C.byref(amg)) == 0, gpu.fasp_cuda_last_error()
        finally:
            gpu.fasp_cuda_set_option(b"coarse_dense", 1.0)
        assert np.linalg.norm(xv - xr) / np.linalg.norm(xr) < 1e-8
    finally:
        ref.amg_free(mgl, amg)
    # singular (pure Neumann) operator with a consistent right-hand side
    import scipy.sparse as sp
    m = 6
    T1 = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(m, m)).tolil()
    T1[0, 0] = T1[m - 1, m - 1] = 1.0
    I = sp.identity(m)
    N = (sp.kron(sp.kron(T1, I), I) + sp.kron(sp.kron(I, T1), I) + sp.kron(sp.kron(I, I), T1)).tocsr()
    N.sort_indices()
    An = T.CSR.from_scipy(N)
    bn = np.random.default_rng(5).uniform(-1, 1, An.shape[0])
    bn -= bn.mean()
    amg1 = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG, max_levels=1)   # the cycle IS the coarse solve
    st, xn = api.fasp_cuda_solver_dcsr_krylov_amg(An, bn, np.zeros_like(bn), it, amg1)
    assert st >= 0, (st, api.last_error())
    assert np.all(np.isfinite(xn))
    assert np.linalg.norm(bn - N @ xn) / np.linalg.norm(bn) <= 1e-8 * 1.001


def test_remaining_entry_points(gpu, ref, data):
    """The entry points no other GPU test names: the device-vector cycle, the itsolver driver with the AMG
    preconditioner object, the relres history, the kernel timer, launch counter, profile dump and sync."""
    A, b = data["FE"], data["FE_b"]
    n = A.shape[0]
    amg = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
    mgl = ref.amg_setup(A, amg)
    try:
        # z = B r with device vectors equals the host-vector cycle on the same handle
        h = gpu.fasp_cuda_amg_upload(mgl, C.byref(amg))
        assert h, gpu.fasp_cuda_last_error()
        r = np.random.default_rng(8).uniform(-1, 1, n)
        z_host = np.zeros(n)
        assert gpu.fasp_cuda_amg_cycle_host(h, T.as_preal(r), T.as_preal(z_host)) == 0
        d_r, d_z = gpu.fasp_cuda_dvec_alloc(n), gpu.fasp_cuda_dvec_alloc(n)
        assert gpu.fasp_cuda_dvec_h2d(d_r, T.as_preal(r), n) == 0
        gpu.fasp_cuda_launch_count_reset()
        assert gpu.fasp_cuda_amg_cycle_dev(h, d_r, d_z) == 0, gpu.fasp_cuda_last_error()
        assert gpu.fasp_cuda_sync() == 0
        assert gpu.fasp_cuda_launch_count() > 0
        z_dev = np.empty(n)
        assert gpu.fasp_cuda_dvec_d2h(T.as_preal(z_dev), d_z, n) == 0
        assert np.array_equal(z_dev, z_host)
        gpu.fasp_cuda_dvec_free(d_r), gpu.fasp_cuda_dvec_free(d_z)
        gpu.fasp_cuda_amg_free(h)
        # fasp_solver_dcsr_itsolver (SolCSR.c:56) with the device preconditioner object, CG and VGMRES
        pc = gpu.fasp_cuda_precond_from_mgl(mgl, C.byref(amg))
        assert pc, gpu.fasp_cuda_last_error()
        for solver in (T.SOLVER_CG, T.SOLVER_VGMRES):
            it = ref.its_param(itsolver_type=solver, tol=1e-8, maxit=200, print_level=0, restart=30)
            vb, vx = T.Vec(b), T.Vec(np.zeros(n))
            st = gpu.fasp_cuda_solver_dcsr_itsolver(A.ptr(), vb.ptr(), vx.ptr(), pc, C.byref(it))
            amg_r = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
            st_ref, x_ref = ref.krylov_amg(A, b, np.zeros(n), it, amg_r)
            assert st > 0 and abs(st - st_ref) <= 1, (solver, st, st_ref, gpu.fasp_cuda_last_error())
            assert np.linalg.norm(vx.a - x_ref) / np.linalg.norm(x_ref) <= 1e-8
        gpu.fasp_cuda_precond_free(pc)
        # relres history of a resident solver: entry 0 is the initial residual (1 for x0 = 0), the last meets the tolerance
        s = api.KrylovAmgSolver(mgl, amg)
        it = ref.its_param(itsolver_type=T.SOLVER_CG, tol=1e-8, maxit=200, print_level=0)
        st, x = s.solve(b, np.zeros(n), it)
        hist = s.history()
        assert st > 0 and hist.size == st + 1 and abs(hist[0] - 1.0) < 1e-12 and hist[-1] <= 1e-8
        assert s.stat(0) == st and s.stat(3) > 0
        s.close()
    finally:
        ref.amg_free(mgl, amg)
    # kernel timer and profile records on a resident matrix
    P7 = PB.poisson7(32)
    hA = gpu.fasp_cuda_dcsr_upload(P7.ptr())
    assert hA, gpu.fasp_cuda_last_error()
    for what in (0, 1, 2, 10, 11):
        ms = gpu.fasp_cuda_dcsr_time_kernel(hA, what, 1, 3, 0)
        assert ms > 0, (what, gpu.fasp_cuda_last_error())
    gpu.fasp_cuda_set_option(b"profile", 1.0)
    try:
        gpu.fasp_cuda_profile_dump(None, 0)
        x = np.ones(P7.shape[1]); y = np.empty(P7.shape[0])
        assert gpu.fasp_cuda_blas_dcsr_mxv(P7.ptr(), T.as_preal(x), T.as_preal(y)) == 0
        buf = C.create_string_buffer(1 << 16)
        gpu.fasp_cuda_profile_dump(buf, len(buf))
        recs = [ln.split() for ln in buf.value.decode().splitlines()]
        assert any(int(k) == 0 and int(rows) == P7.shape[0] and int(nnz) == P7.nnz for k, rows, nnz, ms, by in recs)
    finally:
        gpu.fasp_cuda_set_option(b"profile", 0.0)
    gpu.fasp_cuda_dcsr_free(hA)
