"""Setup-phase pieces on the device: the transpose R = P^T and the Galerkin product
A_c = R A P against the reference's fasp_dcsr_trans (BlaSparseCSR.c:952) / fasp_blas_dcsr_rap (BlaSpmvCSR.c:999),
entry for entry and bit for bit, and the UNMODIFIED fasp_amg_setup_rs driven through the interposition shim."""
import ctypes as C
import hashlib
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from faspsolver_b200 import fasp_types as T
from faspsolver_b200 import problems as PB

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
P = C.POINTER


def _bind(ref):
    L = ref.L
    L.fasp_dcsr_trans.restype = T.INT
    L.fasp_dcsr_trans.argtypes = [P(T.dCSRmat), P(T.dCSRmat)]
    L.fasp_blas_dcsr_rap.restype = None
    L.fasp_blas_dcsr_rap.argtypes = [P(T.dCSRmat)] * 4
    L.fasp_dcsr_free.restype = None
    L.fasp_dcsr_free.argtypes = [P(T.dCSRmat)]
    return L


def _arrays(m):
    ia = np.ctypeslib.as_array(m.IA, shape=(m.row + 1,)).copy()
    ja = np.ctypeslib.as_array(m.JA, shape=(m.nnz,)).copy() if m.nnz else np.zeros(0, np.int32)
    va = np.ctypeslib.as_array(m.val, shape=(m.nnz,)).copy() if (m.nnz and m.val) else None
    return ia, ja, va


def _same(a, b):
    assert (a.row, a.col, a.nnz) == (b.row, b.col, b.nnz), ((a.row, a.col, a.nnz), (b.row, b.col, b.nnz))
    ia, ja, va = _arrays(a)
    ib, jb, vb = _arrays(b)
    assert np.array_equal(ia, ib)
    assert np.array_equal(ja, jb)          # same entry ORDER inside every row
    assert (va is None) == (vb is None)
    if va is not None:
        assert np.array_equal(va, vb)      # same bits


@pytest.mark.parametrize("prob", ["FE", "p7", "cd7", "p27"])
def test_transpose_and_rap_match_reference_bit_for_bit(gpu, ref, data, prob):
    L = _bind(ref)
    A = {"FE": lambda: data["FE"], "p7": lambda: PB.poisson7(18), "cd7": lambda: PB.convdiff7(14),
         "p27": lambda: PB.poisson27(10)}[prob]()
    amg = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG, coarse_dof=50)
    mgl = ref.amg_setup(A, amg)
    try:
        nl = mgl[0].num_levels
        assert nl >= 3
        for l in range(nl - 1):
            Pm, Rm, Am = mgl[l].P, mgl[l].R, mgl[l].A
            for M in (Pm, Am):    # transposes: P^T (= R) and A^T, with and without values
                t_ref, t_gpu = T.dCSRmat(), T.dCSRmat()
                L.fasp_dcsr_trans(C.byref(M), C.byref(t_ref))
                assert gpu.fasp_cuda_dcsr_trans(C.byref(M), C.byref(t_gpu)) == 0, gpu.fasp_cuda_last_error()
                _same(t_gpu, t_ref)
                L.fasp_dcsr_free(C.byref(t_ref)), L.fasp_dcsr_free(C.byref(t_gpu))
            pat = T.dCSRmat(Pm.row, Pm.col, Pm.nnz, Pm.IA, Pm.JA, None)
            t_ref, t_gpu = T.dCSRmat(), T.dCSRmat()
            L.fasp_dcsr_trans(C.byref(pat), C.byref(t_ref))
            assert gpu.fasp_cuda_dcsr_trans(C.byref(pat), C.byref(t_gpu)) == 0
            _same(t_gpu, t_ref)
            L.fasp_dcsr_free(C.byref(t_ref)), L.fasp_dcsr_free(C.byref(t_gpu))
            # Galerkin product: must equal both a fresh reference call and the next level the setup stored
            c_ref, c_gpu = T.dCSRmat(), T.dCSRmat()
            L.fasp_blas_dcsr_rap(C.byref(Rm), C.byref(Am), C.byref(Pm), C.byref(c_ref))
            assert gpu.fasp_cuda_blas_dcsr_rap(C.byref(Rm), C.byref(Am), C.byref(Pm), C.byref(c_gpu)) == 0, \
                gpu.fasp_cuda_last_error()
            _same(c_gpu, c_ref)
            _same(c_gpu, mgl[l + 1].A)
            L.fasp_dcsr_free(C.byref(c_ref)), L.fasp_dcsr_free(C.byref(c_gpu))
    finally:
        ref.amg_free(mgl, amg)


def test_transpose_edge_cases(gpu, ref):
    L = _bind(ref)
    rng = np.random.default_rng(9)
    import scipy.sparse as sp
    for shape, dens in (((1, 1), 1.0), ((5, 9), 0.5), ((40, 3), 0.7), ((300, 200), 0.02), ((64, 64), 0.0)):
        M = sp.random(*shape, density=dens, random_state=rng, format="csr")
        A = T.CSR.from_scipy(M)
        t_ref, t_gpu = T.dCSRmat(), T.dCSRmat()
        L.fasp_dcsr_trans(A.ptr(), C.byref(t_ref))
        assert gpu.fasp_cuda_dcsr_trans(A.ptr(), C.byref(t_gpu)) == 0, gpu.fasp_cuda_last_error()
        _same(t_gpu, t_ref)


SETUP_WORKER = r'''
import sys, time, hashlib
sys.path.insert(0, %(root)r)
import numpy as np
from oracle.ref import RefFasp
from faspsolver_b200 import problems as PB, fasp_types as T
ref = RefFasp()
A = PB.poisson7(40)
amg = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG, AMG_type=int(sys.argv[1]))
t = time.time(); mgl = ref.amg_setup(A, amg); dt = time.time() - t
h = hashlib.sha256()
nl = mgl[0].num_levels
for l in range(nl):
    for nm in ("A", "P", "R"):
        if nm != "A" and l == nl - 1: continue
        m = getattr(mgl[l], nm)
        h.update(np.ctypeslib.as_array(m.IA, shape=(m.row + 1,)).tobytes())
        h.update(np.ctypeslib.as_array(m.JA, shape=(m.nnz,)).tobytes())
        h.update(np.ctypeslib.as_array(m.val, shape=(m.nnz,)).tobytes())
print("DIGEST", h.hexdigest(), nl, "%%.3f" %% dt)
'''


@pytest.mark.parametrize("amg_type", [T.CLASSIC_AMG, T.SA_AMG])
def test_unmodified_reference_setup_through_the_shim(gpu, tmp_path, amg_type):
    """LD_PRELOAD=libfasp_cuda_setup.so: FASP's own fasp_amg_setup_rs / _sa call fasp_dcsr_trans and
    fasp_blas_dcsr_rap, the shim forwards them to the device, and the hierarchy is IDENTICAL (sha256 of every
    array of every level) to the one the plain CPU setup builds."""
    if not (ROOT / "oracle" / "_ref" / "libfasp_seq.so").exists():
        pytest.skip("oracle/_ref/libfasp_seq.so not built")
    script = tmp_path / "w.py"
    script.write_text(SETUP_WORKER % {"root": str(ROOT)})
    out = {}
    for how in ("cpu", "shim"):
        env = dict(os.environ)
        if how == "shim":
            env["LD_PRELOAD"] = str(ROOT / "faspsolver_b200" / "lib" / "libfasp_cuda_setup.so")
        r = subprocess.run([sys.executable, str(script), str(amg_type)], env=env, capture_output=True, text=True,
                           cwd=str(ROOT))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out[how] = [l for l in r.stdout.splitlines() if l.startswith("DIGEST")][0].split()
    assert out["cpu"][1] == out["shim"][1] and out["cpu"][2] == out["shim"][2], out
