"""TEST INFRASTRUCTURE ONLY — ctypes access to oracle/fasp_oracle.c (the plain-C restatement of
the reference's solve-phase algorithms), built by faspsolver_b200.build.build_oracle() into
oracle/_ref/libfasp_oracle.so. Same usage rules as oracle/ref.py."""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

_ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(_ROOT))
from faspsolver_b200 import build as _B  # noqa: E402
from faspsolver_b200.fasp_types import CSR  # noqa: E402

PI, PD, I, D, VP = C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int, C.c_double, C.c_void_p


def _pi(a):
    return a.ctypes.data_as(PI)


def _pd(a):
    return a.ctypes.data_as(PD) if a is not None else None


class Oracle:
    def __init__(self):
        path = _B.build_oracle()
        L = C.CDLL(str(path))
        L.oracle_dcsr_mxv.argtypes = [I, PI, PI, PD, PD, PD]
        L.oracle_dcsr_aAxpy.argtypes = [D, I, PI, PI, PD, PD, PD]
        L.oracle_smoother_jacobi.argtypes = [I, PI, PI, PD, PD, PD, I, D]
        L.oracle_smoother_l1diag.argtypes = [I, PI, PI, PD, PD, PD, I]
        L.oracle_smoother_poly.argtypes = [I, PI, PI, PD, PD, PD, I, I]
        L.oracle_dbsr_mxv.argtypes = [I, I, PI, PI, PD, PD, PD]
        L.oracle_dbsr_aAxpy.argtypes = [D, I, I, PI, PI, PD, PD, PD]
        L.oracle_dbsr_jacobi1.argtypes = [I, I, PI, PI, PD, PD, PD, PD]
        L.oracle_mg_new.argtypes = [I, I, I, I, I, I, I, D, D]
        L.oracle_mg_new.restype = VP
        L.oracle_mg_set_level.argtypes = [VP, I, I, PI, PI, PD, I, I, PI, PI, PD, I, I, PI, PI, PD]
        L.oracle_mg_free.argtypes = [VP]
        L.oracle_precond_amg.argtypes = [VP, PD, PD, I]
        L.oracle_mg_cycle_on.argtypes = [VP, PD, PD]
        L.oracle_pcg.argtypes = [I, PI, PI, PD, PD, PD, VP, D, D, I, PD]
        L.oracle_pcg.restype = I
        L.oracle_gmres.argtypes = [I, PI, PI, PD, PD, PD, VP, D, D, I, I, I, PD]
        L.oracle_gmres.restype = I
        L.oracle_multicolor.argtypes = [I, PI, PI, PI, PI]
        L.oracle_multicolor.restype = I
        L.oracle_gs_multicolor.argtypes = [I, PI, PI, PD, PD, PD, I, I, I, PI, PI]
        L.oracle_gs_multicolor.restype = None
        L.oracle_amg_solve.argtypes = [VP, PD, PD, D, I, PD]
        L.oracle_amg_solve.restype = I
        L.oracle_fgmres.argtypes = [I, PI, PI, PD, PD, PD, VP, D, D, I, I, PD]
        L.oracle_fgmres.restype = I
        for f in ("oracle_dcsr_mxv", "oracle_dcsr_aAxpy", "oracle_smoother_jacobi", "oracle_smoother_l1diag",
                  "oracle_smoother_poly", "oracle_dbsr_mxv", "oracle_dbsr_aAxpy", "oracle_dbsr_jacobi1",
                  "oracle_mg_set_level", "oracle_mg_free", "oracle_precond_amg", "oracle_mg_cycle_on"):
            getattr(L, f).restype = None
        self.L = L

    def mxv(self, A: CSR, x, pattern=False):
        y = np.empty(A.shape[0])
        self.L.oracle_dcsr_mxv(A.shape[0], _pi(A.ia), _pi(A.ja), None if pattern else _pd(A.val),
                               _pd(np.ascontiguousarray(x, dtype=np.float64)), _pd(y))
        return y

    def aAxpy(self, alpha, A: CSR, x, y):
        y = np.array(y, dtype=np.float64, copy=True)
        self.L.oracle_dcsr_aAxpy(alpha, A.shape[0], _pi(A.ia), _pi(A.ja), _pd(A.val),
                                 _pd(np.ascontiguousarray(x, dtype=np.float64)), _pd(y))
        return y

    def jacobi(self, A, b, u, L, w):
        u = np.array(u, dtype=np.float64, copy=True)
        self.L.oracle_smoother_jacobi(A.shape[0], _pi(A.ia), _pi(A.ja), _pd(A.val), _pd(np.ascontiguousarray(b)), _pd(u), L, w)
        return u

    def l1diag(self, A, b, u, L):
        u = np.array(u, dtype=np.float64, copy=True)
        self.L.oracle_smoother_l1diag(A.shape[0], _pi(A.ia), _pi(A.ja), _pd(A.val), _pd(np.ascontiguousarray(b)), _pd(u), L)
        return u

    def poly(self, A, b, u, ndeg, L):
        u = np.array(u, dtype=np.float64, copy=True)
        self.L.oracle_smoother_poly(A.shape[0], _pi(A.ia), _pi(A.ja), _pd(A.val), _pd(np.ascontiguousarray(b)), _pd(u), ndeg, L)
        return u


class OracleMG:
    """The restated multigrid cycle on a hierarchy given as lists of CSR (A_l, P_l, R_l)."""

    def __init__(self, orc: Oracle, levels, smoother, cycle_type=1, presmooth=1, postsmooth=1, ndeg=3,
                 coarse_scaling=0, relax=1.0, tol=1e-6, pattern_transfer=False):
        self.orc, self.levels = orc, levels   # keep arrays alive
        nl = len(levels)
        self.h = orc.L.oracle_mg_new(nl, smoother, cycle_type, presmooth, postsmooth, ndeg, coarse_scaling, relax, tol)
        z = np.zeros(1, dtype=np.int32)
        for l, (A, P, R) in enumerate(levels):
            def parts(M):
                if M is None:
                    return 0, 0, _pi(z), _pi(z), None
                return M.shape[0], M.shape[1], _pi(M.ia), _pi(M.ja), (None if pattern_transfer else _pd(M.val))
            orc.L.oracle_mg_set_level(self.h, l, A.shape[0], _pi(A.ia), _pi(A.ja), _pd(A.val), *parts(P), *parts(R))

    def cycle(self, b, x):
        x = np.array(x, dtype=np.float64, copy=True)
        self.orc.L.oracle_mg_cycle_on(self.h, _pd(np.ascontiguousarray(b)), _pd(x))
        return x

    def pcg(self, A, b, tol=1e-8, maxit=500, precond=True):
        u = np.zeros_like(b)
        rel = C.c_double(0)
        st = self.orc.L.oracle_pcg(A.shape[0], _pi(A.ia), _pi(A.ja), _pd(A.val), _pd(np.ascontiguousarray(b)), _pd(u),
                                   self.h if precond else None, tol, 1e-18, maxit, C.byref(rel))
        return st, u, rel.value

    def gmres(self, A, b, tol=1e-8, maxit=500, restart=30, variable=True, precond=True):
        x = np.zeros_like(b)
        rel = C.c_double(0)
        st = self.orc.L.oracle_gmres(A.shape[0], _pi(A.ia), _pi(A.ja), _pd(A.val), _pd(np.ascontiguousarray(b)), _pd(x),
                                     self.h if precond else None, tol, 1e-18, maxit, restart, int(variable), C.byref(rel))
        return st, x, rel.value

    def amg_solve(self, b, tol=1e-8, maxit=500):
        x = np.zeros_like(b)
        rel = C.c_double(0)
        st = self.orc.L.oracle_amg_solve(self.h, _pd(np.ascontiguousarray(b)), _pd(x), tol, maxit, C.byref(rel))
        return st, x, rel.value

    def fgmres(self, A, b, tol=1e-8, maxit=500, restart=30, precond=True):
        x = np.zeros_like(b)
        rel = C.c_double(0)
        st = self.orc.L.oracle_fgmres(A.shape[0], _pi(A.ia), _pi(A.ja), _pd(A.val), _pd(np.ascontiguousarray(b)), _pd(x),
                                      self.h if precond else None, tol, 1e-18, maxit, restart, C.byref(rel))
        return st, x, rel.value

    def close(self):
        if self.h:
            self.orc.L.oracle_mg_free(self.h)
            self.h = None


def hierarchy_from_mgl(mgl):
    """Copy (A_l, P_l, R_l) of a host FASP hierarchy into CSR objects."""
    nl = mgl[0].num_levels
    out = []
    for l in range(nl):
        A = CSR.from_struct(mgl[l].A)
        P = CSR.from_struct(mgl[l].P) if l < nl - 1 else None
        R = CSR.from_struct(mgl[l].R) if l < nl - 1 else None
        out.append((A, P, R))
    return out
