"""TEST INFRASTRUCTURE ONLY — ctypes access to the UNMODIFIED reference built by
oracle/build_ref.sh (oracle/_ref/libfasp_seq.so = sequential FASP 2.8.7, the parity oracle;
libfasp_omp.so = OpenMP build, CPU timing baseline).

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py
may import this module. The product (faspsolver_b200/lib/libfasp_cuda.so) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

_ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(_ROOT))
from faspsolver_b200 import fasp_types as T  # noqa: E402
from faspsolver_b200.api import HostFasp  # noqa: E402
from faspsolver_b200.fasp_types import (AMG_data, AMG_data_bsr, AMG_param, ITS_param, dBSRmat,  # noqa: E402
                                        dCSRmat, dvector, precond, INT, REAL, SHORT, PREAL)

REF_DIR = Path(__file__).resolve().parent / "_ref"
SEQ = REF_DIR / "libfasp_seq.so"
OMP = REF_DIR / "libfasp_omp.so"


def available() -> bool:
    return SEQ.exists()


class RefFasp(HostFasp):
    """The sequential reference library: both the 'host application's FASP' (setup phase)
    and the oracle for every solve-phase function."""

    def __init__(self, path=None):
        super().__init__(path or SEQ)
        L, P = self.L, C.POINTER
        sig = {
            "fasp_blas_dcsr_mxv": (None, [P(dCSRmat), PREAL, PREAL]),
            "fasp_blas_dcsr_aAxpy": (None, [REAL, P(dCSRmat), PREAL, PREAL]),
            "fasp_blas_dcsr_mxv_agg": (None, [P(dCSRmat), PREAL, PREAL]),
            "fasp_blas_dcsr_aAxpy_agg": (None, [REAL, P(dCSRmat), PREAL, PREAL]),
            "fasp_blas_dbsr_mxv": (None, [P(dBSRmat), PREAL, PREAL]),
            "fasp_blas_dbsr_aAxpy": (None, [REAL, P(dBSRmat), PREAL, PREAL]),
            "fasp_smoother_dcsr_jacobi": (None, [P(dvector), INT, INT, INT, P(dCSRmat), P(dvector), INT, REAL]),
            "fasp_smoother_dcsr_L1diag": (None, [P(dvector), INT, INT, INT, P(dCSRmat), P(dvector), INT]),
            "fasp_smoother_dcsr_poly": (None, [P(dCSRmat), P(dvector), P(dvector), INT, INT, INT]),
            "fasp_smoother_dbsr_jacobi1": (None, [P(dBSRmat), P(dvector), P(dvector), PREAL]),
            "fasp_solver_mgcycle": (None, [P(AMG_data), P(AMG_param)]),
            "fasp_solver_mgcycle_bsr": (None, [P(AMG_data_bsr), P(AMG_param)]),
            "fasp_amg_solve": (INT, [P(AMG_data), P(AMG_param)]),
            "fasp_solver_dcsr_pcg": (INT, [P(dCSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT]),
            "fasp_solver_dcsr_pvgmres": (INT, [P(dCSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT, SHORT]),
            "fasp_solver_dcsr_pgmres": (INT, [P(dCSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT, SHORT]),
            "fasp_solver_dbsr_pcg": (INT, [P(dBSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT]),
            "fasp_solver_dbsr_pvgmres": (INT, [P(dBSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT, SHORT]),
            "fasp_solver_dcsr_pvfgmres": (INT, [P(dCSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT, SHORT]),
            "fasp_solver_dbsr_pgmres": (INT, [P(dBSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT, SHORT]),
            "fasp_solver_dbsr_pvfgmres": (INT, [P(dBSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT, SHORT]),
            "fasp_solver_dcsr_krylov_amg": (INT, [P(dCSRmat), P(dvector), P(dvector), P(ITS_param), P(AMG_param)]),
            "fasp_solver_dbsr_krylov_amg": (INT, [P(dBSRmat), P(dvector), P(dvector), P(ITS_param), P(AMG_param)]),
            "fasp_solver_dcsr_itsolver": (INT, [P(dCSRmat), P(dvector), P(dvector), P(precond), P(ITS_param)]),
            "fasp_precond_amg": (None, [PREAL, PREAL, C.c_void_p]),
            "fasp_param_amg_to_prec": (None, [C.c_void_p, P(AMG_param)]),
            "fasp_dbsr_getdiaginv": (dvector, [P(dBSRmat)]),
            "fasp_solver_amg": (INT, [P(dCSRmat), P(dvector), P(dvector), P(AMG_param)]),
            "fasp_blas_dcsr_vmv": (REAL, [P(dCSRmat), PREAL, PREAL]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args

    # convenience wrappers returning numpy
    def mxv(self, A, x):
        y = np.empty(A.shape[0])
        self.L.fasp_blas_dcsr_mxv(A.ptr(), T.as_preal(np.ascontiguousarray(x, dtype=np.float64)), T.as_preal(y))
        return y

    def aAxpy(self, alpha, A, x, y):
        y = np.array(y, dtype=np.float64, copy=True)
        self.L.fasp_blas_dcsr_aAxpy(alpha, A.ptr(), T.as_preal(np.ascontiguousarray(x, dtype=np.float64)), T.as_preal(y))
        return y

    def krylov_amg(self, A, b, x0, itparam, amgparam):
        vb, vx = T.Vec(b), T.Vec(np.array(x0, dtype=np.float64, copy=True))
        st = self.L.fasp_solver_dcsr_krylov_amg(A.ptr(), vb.ptr(), vx.ptr(), C.byref(itparam), C.byref(amgparam))
        return st, vx.a
