"""CPU tests (no GPU): the oracle restatement (oracle/fasp_oracle.c) is pinned against
 (a) the committed golden vectors / answers produced by the unmodified reference, and
 (b) the reference library itself (oracle/_ref/libfasp_seq.so) when it has been built here,
including the reference's own golden log test/out/reg.gcc (via tests/golden/oracle_answers.json)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from faspsolver_b200 import fasp_types as T
from faspsolver_b200 import problems as PB


@pytest.fixture(scope="module")
def orc():
    from oracle.port import Oracle
    return Oracle()


def test_oracle_kernels_match_golden_vectors(orc, data, golden_vectors):
    A, g, b = data["FE"], golden_vectors, data["FE_b"]
    assert np.array_equal(orc.mxv(A, g["x"]), g["mxv"])
    assert np.array_equal(orc.aAxpy(-1.0, A, g["x"], b), g["aAxpy_m1"])
    assert np.array_equal(orc.aAxpy(0.3, A, g["x"], b), g["aAxpy_0p3"])
    assert np.array_equal(orc.jacobi(A, b, g["x"], 2, 0.67), g["smooth_jacobi067"])
    assert np.array_equal(orc.l1diag(A, b, g["x"], 2), g["smooth_l1diag"])
    assert np.array_equal(orc.poly(A, b, g["x"], 3, 2), g["smooth_poly3"])


def test_oracle_kernels_match_reference_live(orc, ref, data):
    rng = np.random.default_rng(41)
    for A in (data["FD"], data["FE"], PB.poisson7(12), PB.convdiff7(10), PB.poisson27(8)):
        n = A.shape[0]
        x, b = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        assert np.array_equal(orc.mxv(A, x), ref.mxv(A, x))
        assert np.array_equal(orc.aAxpy(0.7, A, x, b), ref.aAxpy(0.7, A, x, b))
        u = T.Vec(x.copy())
        ref.L.fasp_smoother_dcsr_L1diag(u.ptr(), 0, n - 1, 1, A.ptr(), T.Vec(b).ptr(), 2)
        assert np.array_equal(orc.l1diag(A, b, x, 2), u.a)
        u = T.Vec(x.copy())
        ref.L.fasp_smoother_dcsr_jacobi(u.ptr(), n - 1, 0, -1, A.ptr(), T.Vec(b).ptr(), 3, 0.8)
        assert np.array_equal(orc.jacobi(A, b, x, 3, 0.8), u.a)
        u = T.Vec(x.copy())
        ref.L.fasp_smoother_dcsr_poly(A.ptr(), T.Vec(b).ptr(), u.ptr(), n, 4, 1)
        assert np.array_equal(orc.poly(A, b, x, 4, 1), u.a)


def test_oracle_bsr_matches_reference(orc, ref, data):
    from oracle.port import _pd, _pi
    rng = np.random.default_rng(42)
    for A in (data["SPE"], PB.blockoil7(4)[0]):
        n = A.ROW * A.nb
        x, y0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        y, yr = np.empty(n), np.empty(n)
        orc.L.oracle_dbsr_mxv(A.ROW, A.nb, _pi(A.ia), _pi(A.ja), _pd(A.val), _pd(x), _pd(y))
        ref.L.fasp_blas_dbsr_mxv(A.ptr(), T.as_preal(x), T.as_preal(yr))
        assert np.array_equal(y, yr)
        for alpha in (1.0, -1.0, 0.4):
            y, yr = y0.copy(), y0.copy()
            orc.L.oracle_dbsr_aAxpy(alpha, A.ROW, A.nb, _pi(A.ia), _pi(A.ja), _pd(A.val), _pd(x), _pd(y))
            ref.L.fasp_blas_dbsr_aAxpy(alpha, A.ptr(), T.as_preal(x), T.as_preal(yr))
            assert np.array_equal(y, yr)
        dinv = ref.L.fasp_dbsr_getdiaginv(A.ptr())
        dnp = np.ctypeslib.as_array(dinv.val, shape=(A.ROW * A.nb * A.nb,)).copy()
        u, ur = x.copy(), T.Vec(x.copy())
        orc.L.oracle_dbsr_jacobi1(A.ROW, A.nb, _pi(A.ia), _pi(A.ja), _pd(A.val), _pd(y0), _pd(u), _pd(dnp))
        ref.L.fasp_smoother_dbsr_jacobi1(A.ptr(), T.Vec(y0).ptr(), ur.ptr(), dinv.val)
        assert np.array_equal(u, ur.a)


ORACLE_SOLVES = [
    ("FE_pcg_jacobi067", "pcg", {}), ("FE_pcg_l1", "pcg", {}), ("FE_pcg_poly3", "pcg", {}),
    ("FE_pcg_l1_W", "pcg", {}), ("FE_gmres30_l1", "gmres", dict(variable=False)),
    ("FE_vgmres30_poly3", "gmres", dict(variable=True)), ("FD_pcg_jacobi067_cdof20", "pcg", {}),
    ("FD_pcg_l1_cdof20", "pcg", {}), ("FE_vfgmres30_l1", "fgmres", {}),
]


@pytest.mark.parametrize("name,method,kw", ORACLE_SOLVES)
def test_oracle_solves_reproduce_reference_answers(orc, ref, data, golden_answers, name, method, kw):
    """Iteration counts of the restated PCG / GMRES + V-cycle equal the committed answers of the
    unmodified reference (hierarchy from the reference's own setup)."""
    from oracle.port import OracleMG, hierarchy_from_mgl
    rec = {r["name"]: r for r in golden_answers["recipes"]}[name]
    prob = name.split("_")[0]
    A, b = data[prob], data[prob + "_b"]
    amg = ref.amg_param(print_level=0, **rec["amg"])
    mgl = ref.amg_setup(A, amg)
    try:
        lv = hierarchy_from_mgl(mgl)
    finally:
        ref.amg_free(mgl, amg)
    mg = OracleMG(orc, lv, smoother=amg.smoother, cycle_type=amg.cycle_type, presmooth=amg.presmooth_iter,
                  postsmooth=amg.postsmooth_iter, ndeg=amg.polynomial_degree, relax=amg.relaxation, tol=1e-6)
    if method == "pcg":
        st, x, rel = mg.pcg(A, b, tol=1e-8)
    elif method == "fgmres":
        st, x, rel = mg.fgmres(A, b, tol=1e-8, restart=rec["it"]["restart"])
    else:
        st, x, rel = mg.gmres(A, b, tol=1e-8, restart=rec["it"]["restart"], **kw)
    mg.close()
    assert st == rec["status"], (name, st, rec["status"])
    assert np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b) <= 1e-8 * 1.001


def test_oracle_cycle_matches_reference_cycle(orc, ref, data):
    from oracle.port import OracleMG, hierarchy_from_mgl
    A, b = data["FE"], data["FE_b"]
    for kw in (dict(smoother=T.SMOOTHER_L1DIAG), dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67, cycle_type=T.W_CYCLE),
               dict(smoother=T.SMOOTHER_POLY, polynomial_degree=3, coarse_scaling=T.ON)):
        amg = ref.amg_param(print_level=0, **kw)
        mgl = ref.amg_setup(A, amg)
        try:
            n = A.shape[0]
            np.ctypeslib.as_array(mgl[0].b.val, shape=(n,))[:] = b
            np.ctypeslib.as_array(mgl[0].x.val, shape=(n,))[:] = 0.0
            ref.L.fasp_solver_mgcycle(mgl, C.byref(amg))
            x_ref = np.ctypeslib.as_array(mgl[0].x.val, shape=(n,)).copy()
            lv = hierarchy_from_mgl(mgl)
        finally:
            ref.amg_free(mgl, amg)
        mg = OracleMG(orc, lv, smoother=amg.smoother, cycle_type=amg.cycle_type, ndeg=amg.polynomial_degree,
                      relax=amg.relaxation, coarse_scaling=amg.coarse_scaling, tol=amg.tol)
        x = mg.cycle(b, np.zeros(n))
        mg.close()
        assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-9, kw


def test_reference_reproduces_reg_gcc_golden_lines(ref, data, golden_answers):
    """The compiled reference (our oracle's anchor) reproduces test/out/reg.gcc bit for bit on the
    lines that pin this path: FE AMG-PCG 6 it / 2.728796e-11, FD 1 it / 4.938174e-15, and the
    L1_DIAG AMG solver 19 it / 8.612004e-11."""
    g = golden_answers["reg_gcc"]
    for prob, key in (("FE", "FE_amg_pcg_default_tol1e-10"), ("FD", "FD_amg_pcg_default_tol1e-10")):
        A, b = data[prob], data[prob + "_b"]
        it = ref.its_param(maxit=500, tol=1e-10, print_level=0)
        amg = ref.amg_param(print_level=0)
        st, x = ref.krylov_amg(A, b, np.zeros_like(b), it, amg)
        assert st == g[key]["iters"]
        assert np.abs(x - data[prob + "_sol"]).max() < 1e-4
    A, b = data["FE"], data["FE_b"]
    amg = ref.amg_param(print_level=0, maxit=500, tol=1e-10, smoother=T.SMOOTHER_L1DIAG)
    vb, vx = T.Vec(b), T.Vec(np.zeros_like(b))
    st = ref.L.fasp_solver_amg(A.ptr(), vb.ptr(), vx.ptr(), C.byref(amg))
    assert st == g["FE_amg_solver_L1DIAG_tol1e-10"]["iters"]
    rel = np.linalg.norm(b - A.to_scipy() @ vx.a) / np.linalg.norm(b)
    assert abs(rel - g["FE_amg_solver_L1DIAG_tol1e-10"]["relres"]) / rel < 1e-5


def test_oracle_flexible_gmres_reproduces_reg_gcc_line(orc, ref, data, golden_answers):
    """Unpreconditioned VFGMRES on the FE problem (test/main/regression.c:492-505, tol 1e-12, default
    restart 25): reg.gcc pins 493 iterations / 7.667271e-13. The restatement and the compiled
    reference both reproduce it, bit-identical to each other."""
    from oracle.port import OracleMG
    g = golden_answers["reg_gcc"]["FE_vfgmres_unprec_tol1e-12"]
    A, b = data["FE"], data["FE_b"]
    n = A.shape[0]
    mg = OracleMG.__new__(OracleMG)   # no hierarchy needed for pc == NULL
    mg.orc, mg.h = orc, None
    st, x, rel = mg.fgmres(A, b, tol=1e-12, maxit=5000, restart=25, precond=False)
    vb, vx = T.Vec(b), T.Vec(np.zeros(n))
    st_ref = ref.L.fasp_solver_dcsr_pvfgmres(A.ptr(), vb.ptr(), vx.ptr(), None, 1e-12, 1e-20, 5000, 25, 1, 0)
    assert st == st_ref == g["iters"]
    assert float("%.6e" % rel) == g["relres"]
    assert np.array_equal(x, vx.a)
    assert np.abs(x - data["FE_sol"]).max() < 1e-4


def test_oracle_amg_solver_reproduces_reg_gcc_line(orc, ref, data, golden_answers):
    """AMG V-cycle with the L1_DIAG smoother as an iterative solver on the FE problem (regression.c:275-289):
    reg.gcc:412 pins 19 iterations / 8.612004e-11. The restated loop (oracle_amg_solve over the restated
    cycle; only the coarsest solve is a plain instead of a safeguarded CG) gives the same count and the same
    printed residual on the reference's own hierarchy."""
    from oracle.port import OracleMG, hierarchy_from_mgl
    g = golden_answers["reg_gcc"]["FE_amg_solver_L1DIAG_tol1e-10"]
    A, b = data["FE"], data["FE_b"]
    amg = ref.amg_param(print_level=0, maxit=500, tol=1e-10, smoother=T.SMOOTHER_L1DIAG)
    mgl = ref.amg_setup(A, amg)
    try:
        lv = hierarchy_from_mgl(mgl)
    finally:
        ref.amg_free(mgl, amg)
    mg = OracleMG(orc, lv, smoother=amg.smoother, cycle_type=amg.cycle_type, presmooth=amg.presmooth_iter,
                  postsmooth=amg.postsmooth_iter, ndeg=amg.polynomial_degree, relax=amg.relaxation, tol=amg.tol)
    st, x, rel = mg.amg_solve(b, tol=1e-10, maxit=500)
    mg.close()
    assert st == g["iters"], st
    assert abs(rel - g["relres"]) / g["relres"] < 1e-5, rel
    assert np.abs(x - data["FE_sol"]).max() < 1e-4


_OMP_WORKER = r'''
import ctypes as C, json, sys
import numpy as np
z = np.load(sys.argv[2])
ia, ja, val = (np.ascontiguousarray(z[k]) for k in ("ia", "ja", "val"))
b, u0 = np.ascontiguousarray(z["b"]), np.ascontiguousarray(z["u0"])
n = ia.size - 1
PI, PD = C.POINTER(C.c_int), C.POINTER(C.c_double)
class dCSRmat_omp(C.Structure):   # fasp.h:151-180 with MULTI_COLOR_ORDER (the OpenMP build's layout)
    _fields_ = [("row", C.c_int), ("col", C.c_int), ("nnz", C.c_int), ("IA", PI), ("JA", PI), ("val", PD),
                ("color", C.c_int), ("IC", PI), ("ICMAP", PI)]
class dvector(C.Structure):
    _fields_ = [("row", C.c_int), ("val", PD)]
L = C.CDLL(sys.argv[1])
A = dCSRmat_omp(n, n, int(ia[n]), ia.ctypes.data_as(PI), ja.ctypes.data_as(PI), val.ctypes.data_as(PD), 0, None, None)
rowmax, groups = C.c_int(0), C.c_int(0)
L.dCSRmat_Multicoloring(C.byref(A), C.byref(rowmax), C.byref(groups))
ic = [A.IC[k] for k in range(A.color + 1)]
icmap = [A.ICMAP[k] for k in range(n)]
out = {"color": A.color, "IC": ic, "ICMAP": icmap, "u": {}}
for order in (1, -1):
    u = u0.copy()
    vu, vb = dvector(n, u.ctypes.data_as(PD)), dvector(n, b.ctypes.data_as(PD))
    L.fasp_smoother_dcsr_gs_multicolor(C.byref(vu), C.byref(A), C.byref(vb), 2, order)
    out["u"][str(order)] = u.tolist()
print(json.dumps(out))
'''


@pytest.mark.parametrize("prob", ["FE", "p7", "p27"])
def test_multicolour_gs_restatement_equals_openmp_reference(orc, data, tmp_path, prob):
    """Colouring (dCSRmat_Multicoloring, BlaSparseCSR.c:1687) and two multicolour GS sweeps in both colour
    orders (fasp_smoother_dcsr_gs_multicolor, :2123) of the unmodified OpenMP reference build, run in a
    separate process (its dCSRmat layout differs from the sequential build's): the plain-C restatement
    gives the same colour classes and bit-identical iterates, and so does the host colouring the
    library ships for its device smoother (fasp_cuda_multicolor_host)."""
    import subprocess
    import sys
    from oracle import ref as R
    from oracle.port import _pd, _pi
    from faspsolver_b200 import api
    if not R.OMP.exists():
        pytest.skip("oracle/_ref/libfasp_omp.so not built")
    A = data["FE"] if prob == "FE" else (PB.poisson7(9) if prob == "p7" else PB.poisson27(7))
    n = A.shape[0]
    rng = np.random.default_rng(17)
    b, u0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    np.savez(tmp_path / "in.npz", ia=A.ia, ja=A.ja, val=A.val, b=b, u0=u0)
    script = tmp_path / "omp_worker.py"
    script.write_text(_OMP_WORKER)
    r = subprocess.run([sys.executable, str(script), str(R.OMP), str(tmp_path / "in.npz")], capture_output=True,
                       text=True, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0, r.stderr[-2000:]
    want = json.loads(r.stdout.strip().splitlines()[-1])
    ic, icmap = np.zeros(n + 2, dtype=np.int32), np.zeros(n, dtype=np.int32)
    ncol = orc.L.oracle_multicolor(n, _pi(A.ia), _pi(A.ja), _pi(ic), _pi(icmap))
    assert ncol == want["color"]
    assert ic[:ncol + 1].tolist() == want["IC"] and icmap.tolist() == want["ICMAP"]
    ic2, icmap2 = (C.c_int * (n + 2))(), (C.c_int * n)()
    assert api.lib().fasp_cuda_multicolor_host(n, _pi(A.ia), _pi(A.ja), ic2, icmap2) == ncol
    assert list(ic2[:ncol + 1]) == want["IC"] and list(icmap2) == want["ICMAP"]
    for order in (1, -1):
        u = u0.copy()
        orc.L.oracle_gs_multicolor(n, _pi(A.ia), _pi(A.ja), _pd(A.val), _pd(b), _pd(u), 2, order, ncol, _pi(ic), _pi(icmap))
        assert np.array_equal(u, np.array(want["u"][str(order)])), (prob, order)


def test_unpreconditioned_krylov_golden_lines(orc, ref, data, golden_answers):
    """test/out/reg.gcc, FE problem, methods on this path without a preconditioner (regression.c:300-640):
    the restated CG / GMRES(25) / vGMRES(25) reproduce the iteration counts and the printed residuals to all
    7 digits, and the compiled reference reproduces the "BSR format" lines (the same matrix as 1x1 blocks:
    CG 244, GMRES stops at MaxIt = 500, vGMRES / vFGMRES 339) including the logged max-norm error."""
    g = golden_answers["reg_gcc"]
    A, b, sol = data["FE"], data["FE_b"], data["FE_sol"]
    n = A.shape[0]
    from oracle.port import OracleMG
    mg = OracleMG.__new__(OracleMG)   # no hierarchy: pc == NULL
    mg.orc, mg.h = orc, None
    for key, run in (("FE_cg_unprec_tol1e-12", lambda: mg.pcg(A, b, tol=1e-12, maxit=5000, precond=False)),
                     ("FE_gmres_unprec_tol1e-12",
                      lambda: mg.gmres(A, b, tol=1e-12, maxit=5000, restart=25, variable=False, precond=False)),
                     ("FE_vgmres_unprec_tol1e-12",
                      lambda: mg.gmres(A, b, tol=1e-12, maxit=5000, restart=25, variable=True, precond=False))):
        st, x, rel = run()
        assert st == g[key]["iters"], (key, st)
        assert float("%.6e" % rel) == g[key]["relres"], (key, rel)
        assert np.abs(x - sol).max() < 1e-4
    Ab = T.BSR(n, n, 1, A.ia, A.ja, A.val)
    S = A.to_scipy()
    for key, fn, extra in (("FE_bsr_cg_unprec_tol1e-12", ref.L.fasp_solver_dbsr_pcg, (1e-12, 1e-20, 500, 1, 0)),
                           ("FE_bsr_gmres_unprec_tol1e-8", ref.L.fasp_solver_dbsr_pgmres, (1e-8, 1e-20, 500, 25, 1, 0)),
                           ("FE_bsr_vgmres_unprec_tol1e-8", ref.L.fasp_solver_dbsr_pvgmres, (1e-8, 1e-20, 500, 25, 1, 0)),
                           ("FE_bsr_vfgmres_unprec_tol1e-8", ref.L.fasp_solver_dbsr_pvfgmres, (1e-8, 1e-20, 500, 25, 1, 0))):
        vb, vx = T.Vec(b), T.Vec(np.zeros(n))
        st = fn(Ab.ptr(), vb.ptr(), vx.ptr(), None, *extra)
        want = g[key]["iters"]
        assert st == (want if want < 500 else T.ERROR_SOLVER_MAXIT), (key, st)
        rel = np.linalg.norm(b - S @ vx.a) / np.linalg.norm(b)
        assert abs(rel - g[key]["relres"]) / g[key]["relres"] < 1e-5, (key, rel)
        assert float("%.4e" % np.abs(vx.a - sol).max()) == g[key]["maxdiff"], key
