"""Error behaviour and printed output of the drop-in entry points (FASP's conventions)."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from faspsolver_b200 import api
from faspsolver_b200 import fasp_types as T
from faspsolver_b200 import problems as PB

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_unsupported_smoother_and_solver_codes(gpu, ref, data):
    A, b = data["FE"], data["FE_b"]
    it = ref.its_param(itsolver_type=T.SOLVER_CG, tol=1e-8, maxit=100, print_level=0)
    amg = ref.amg_param(print_level=0)          # default smoother = sequential GS (no data-parallel form)
    st, _ = api.fasp_cuda_solver_dcsr_krylov_amg(A, b, np.zeros_like(b), it, amg)
    assert st == T.ERROR_AMG_SMOOTH_TYPE and "smoother" in api.last_error()
    amg = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
    it2 = ref.its_param(itsolver_type=2, tol=1e-8, maxit=100, print_level=0)   # BiCGstab: not on the path
    st, _ = api.fasp_cuda_solver_dcsr_krylov_amg(A, b, np.zeros_like(b), it2, amg)
    assert st == T.ERROR_SOLVER_TYPE
    it3 = ref.its_param(itsolver_type=T.SOLVER_CG, tol=1e-8, maxit=100, print_level=0, stop_type=7)   # no such type
    st, _ = api.fasp_cuda_solver_dcsr_krylov_amg(A, b, np.zeros_like(b), it3, amg)
    assert st == T.ERROR_INPUT_PAR
    it4 = ref.its_param(itsolver_type=T.SOLVER_GMRES, tol=1e-8, maxit=100, print_level=0, stop_type=7)
    st, _ = api.fasp_cuda_solver_dcsr_krylov_amg(A, b, np.zeros_like(b), it4, amg)
    assert st == T.ERROR_INPUT_PAR
    # a BSR coarsest level beyond coarse_dense_max has no iterative fallback: loud error, not a wrong answer
    Ab, bb = PB.blockoil7(6)
    gpu.fasp_cuda_set_option(b"coarse_dense_max", 16.0)
    try:
        vb, vx = T.Vec(bb), T.Vec(np.zeros_like(bb))
        itb = ref.its_param(itsolver_type=T.SOLVER_VGMRES, tol=1e-8, maxit=100, print_level=0)
        amgb = ref.amg_param(print_level=0, AMG_type=T.UA_AMG, aggregation_type=T.VMB, smoother=T.SMOOTHER_JACOBI,
                             coarse_dof=100)
        st = gpu.fasp_cuda_solver_dbsr_krylov_amg(Ab.ptr(), vb.ptr(), vx.ptr(), C.byref(itb), C.byref(amgb))
    finally:
        gpu.fasp_cuda_set_option(b"coarse_dense_max", 8192.0)
    assert st == T.ERROR_AMG_SETUP, st


def test_gs_as_multicolor_option(gpu, ref, data):
    """SMOOTHER_GS is accepted as multicolour GS (what FASP's OpenMP build runs) only on request."""
    A, b = data["FE"], data["FE_b"]
    it = ref.its_param(itsolver_type=T.SOLVER_CG, tol=1e-8, maxit=100, print_level=0)
    gpu.fasp_cuda_set_option(b"gs_multicolor", 1.0)
    try:
        st, x = api.fasp_cuda_solver_dcsr_krylov_amg(A, b, np.zeros_like(b), it, ref.amg_param(print_level=0))
    finally:
        gpu.fasp_cuda_set_option(b"gs_multicolor", 0.0)
    assert 0 < st <= 12, (st, api.last_error())
    assert np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b) <= 1e-8


def test_maxit_and_zero_rhs(gpu, ref, data):
    A, b = data["FE"], data["FE_b"]
    n = A.shape[0]
    vb, vx = T.Vec(b), T.Vec(np.zeros(n))
    st = gpu.fasp_cuda_solver_dcsr_pcg(A.ptr(), vb.ptr(), vx.ptr(), None, 1e-12, 1e-30, 5, 1, 0)
    vxr = T.Vec(np.zeros(n))
    st_ref = ref.L.fasp_solver_dcsr_pcg(A.ptr(), vb.ptr(), vxr.ptr(), None, 1e-12, 1e-30, 5, 1, 0)
    assert st == st_ref == T.ERROR_SOLVER_MAXIT
    assert np.linalg.norm(vx.a - vxr.a) / np.linalg.norm(vxr.a) < 1e-10     # same 5 iterates
    z, vx0 = T.Vec(np.zeros(n)), T.Vec(np.zeros(n))                         # b = 0: zero iterations
    assert gpu.fasp_cuda_solver_dcsr_pcg(A.ptr(), z.ptr(), vx0.ptr(), None, 1e-8, 1e-18, 50, 1, 0) == 0
    assert not vx0.a.any()
    g1, g2 = T.Vec(np.zeros(n)), T.Vec(np.zeros(n))   # keep the arrays alive across the calls
    st = gpu.fasp_cuda_solver_dcsr_pvgmres(A.ptr(), vb.ptr(), g1.ptr(), None, 1e-12, 1e-30, 7, 5, 1, 0)
    st_ref = ref.L.fasp_solver_dcsr_pvgmres(A.ptr(), vb.ptr(), g2.ptr(), None, 1e-12, 1e-30, 7, 5, 1, 0)
    assert st == st_ref == T.ERROR_SOLVER_MAXIT
    # flexible GMRES: MaxIt, zero right-hand side (returns 0 without touching x), and an initial
    # guess that already solves the system (KryPvfgmres.c:161 "no need to iterate")
    f1, f2 = T.Vec(np.zeros(n)), T.Vec(np.zeros(n))
    st = gpu.fasp_cuda_solver_dcsr_pvfgmres(A.ptr(), vb.ptr(), f1.ptr(), None, 1e-12, 1e-30, 7, 5, 1, 0)
    st_ref = ref.L.fasp_solver_dcsr_pvfgmres(A.ptr(), vb.ptr(), f2.ptr(), None, 1e-12, 1e-30, 7, 5, 1, 0)
    assert st == st_ref == T.ERROR_SOLVER_MAXIT
    assert np.linalg.norm(f1.a - f2.a) / np.linalg.norm(f2.a) < 1e-10
    f0 = T.Vec(np.zeros(n))
    assert gpu.fasp_cuda_solver_dcsr_pvfgmres(A.ptr(), z.ptr(), f0.ptr(), None, 1e-8, 1e-18, 50, 5, 1, 0) == 0
    assert not f0.a.any()
    xs = T.Vec(data["FE_sol"].copy())
    bs = T.Vec(A.to_scipy() @ xs.a)
    assert gpu.fasp_cuda_solver_dcsr_pvfgmres(A.ptr(), bs.ptr(), xs.ptr(), None, 1e-8, 1e-18, 50, 5, 1, 0) == 0


def test_printed_iteration_table_matches_reference_format(gpu, ref, data, tmp_path):
    """print_level >= PRINT_SOME prints FASP's own table (fasp_itinfo, AuxMessage.c:41-76) and final
    line (ITS_FINAL, KryUtil.inl:93-103): compare with the reference's stdout line by line."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from faspsolver_b200 import api, fasp_types as T
from oracle.ref import RefFasp
z = np.load(%r)
A = T.CSR(z["FE_ia"].size - 1, z["FE_ia"].size - 1, z["FE_ia"], z["FE_ja"], z["FE_val"]); b = z["FE_b"]
ref = RefFasp()
which = sys.argv[1]
kind = int(sys.argv[2])
x0 = np.zeros_like(b)
if kind in (1, 3):     # zero right-hand side: FASP jumps to FINISHED before the iteration table (KryPcg.c:156)
    b = np.zeros_like(b)
it = ref.its_param(itsolver_type=(T.SOLVER_CG, T.SOLVER_CG, T.SOLVER_VGMRES, T.SOLVER_VGMRES)[kind], tol=1e-8,
                   maxit=100, print_level=2)
amg = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
if which == "gpu":
    L = api.lib(); L.fasp_cuda_init(0)
    st, x = api.fasp_cuda_solver_dcsr_krylov_amg(A, b, x0, it, amg)
else:
    st, x = ref.krylov_amg(A, b, x0, it, amg)
sys.stdout.flush()
''' % (str(ROOT), str(ROOT / "tests" / "golden" / "fasp_data.npz"))
    for kind in (0, 1, 2, 3):
        outs = {}
        for which in ("gpu", "ref"):
            r = subprocess.run([sys.executable, "-c", code, which, str(kind)], capture_output=True, text=True,
                               cwd=str(ROOT))
            assert r.returncode == 0, r.stderr[-2000:]
            outs[which] = [l for l in r.stdout.splitlines()
                           if "|" in l or l.startswith("Number of iterations") or l.startswith("---")]
        assert len(outs["gpu"]) == len(outs["ref"]), (kind, outs)
        assert len(outs["gpu"]) > (8 if kind in (0, 2) else 0), (kind, outs)
        for lg, lr in zip(outs["gpu"], outs["ref"]):
            if lg == lr:
                continue
            # same layout; numbers may differ in the last printed digit (reduction order)
            fg, fr = lg.replace("|", " ").split(), lr.replace("|", " ").split()
            assert len(fg) == len(fr) and len(lg) == len(lr), (lg, lr)
            for a, b_ in zip(fg, fr):
                a, b_ = a.rstrip("."), b_.rstrip(".")
                try:
                    assert abs(float(a) - float(b_)) <= 2e-6 * max(abs(float(b_)), 1e-300) + 1.01e-4, (lg, lr)
                except ValueError:
                    assert a == b_, (lg, lr)
