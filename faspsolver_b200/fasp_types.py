"""ctypes mirrors of the FASP structs that cross the libfasp_cuda boundary.

Same member names and order as include/fasp_cuda_types.h (which mirrors base/include/fasp.h
and fasp_block.h of FASP 2.8.7, sequential ABI). tests/test_boundary.py checks the sizes against
the C header and the reference's own headers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

SHORT, INT, REAL = C.c_short, C.c_int, C.c_double
PREAL = C.POINTER(REAL)
PINT = C.POINTER(INT)

# ---- constants (fasp_const.h) ---------------------------------------------------------
FASP_SUCCESS = 0
ERROR_INPUT_PAR, ERROR_ALLOC_MEM, ERROR_DATA_STRUCTURE = -13, -20, -21
ERROR_AMG_SMOOTH_TYPE, ERROR_AMG_SETUP = -31, -39
ERROR_SOLVER_TYPE, ERROR_SOLVER_PRECTYPE, ERROR_SOLVER_STAG = -40, -41, -42
ERROR_SOLVER_SOLSTAG, ERROR_SOLVER_TOLSMALL, ERROR_SOLVER_MISC, ERROR_SOLVER_MAXIT = -43, -44, -46, -48
PRINT_NONE, PRINT_MIN, PRINT_SOME, PRINT_MORE = 0, 1, 2, 4
SOLVER_DEFAULT, SOLVER_CG, SOLVER_GMRES, SOLVER_VGMRES, SOLVER_VFGMRES = 0, 1, 4, 5, 6
STOP_REL_RES, STOP_REL_PRECRES, STOP_MOD_REL_RES = 1, 2, 3
PREC_NULL, PREC_DIAG, PREC_AMG = 0, 1, 2
CLASSIC_AMG, SA_AMG, UA_AMG = 1, 2, 3
PAIRWISE, VMB = 1, 2
V_CYCLE, W_CYCLE, AMLI_CYCLE, NL_AMLI_CYCLE, VW_CYCLE, WV_CYCLE = 1, 2, 3, 4, 12, 21
SMOOTHER_JACOBI, SMOOTHER_GS, SMOOTHER_SGS, SMOOTHER_POLY, SMOOTHER_L1DIAG = 1, 2, 3, 9, 10
COARSE_RS = 1
MIN_CDOF = 20   # fasp_const.h:260
INTERP_DIR = 1
ON, OFF = 1, 0


class dCSRmat(C.Structure):
    _fields_ = [("row", INT), ("col", INT), ("nnz", INT), ("IA", PINT), ("JA", PINT), ("val", PREAL)]


class dvector(C.Structure):
    _fields_ = [("row", INT), ("val", PREAL)]


class ivector(C.Structure):
    _fields_ = [("row", INT), ("val", PINT)]


class dBSRmat(C.Structure):
    _fields_ = [("ROW", INT), ("COL", INT), ("NNZ", INT), ("nb", INT), ("storage_manner", INT),
                ("val", PREAL), ("IA", PINT), ("JA", PINT)]


class ITS_param(C.Structure):
    _fields_ = [("print_level", SHORT), ("itsolver_type", SHORT), ("decoup_type", SHORT),
                ("precond_type", SHORT), ("stop_type", SHORT), ("restart", INT), ("maxit", INT),
                ("tol", REAL), ("abstol", REAL)]


class ILU_param(C.Structure):
    _fields_ = [("print_level", SHORT), ("ILU_type", SHORT), ("ILU_lfil", INT),
                ("ILU_droptol", REAL), ("ILU_relax", REAL), ("ILU_permtol", REAL)]


class AMG_param(C.Structure):
    _fields_ = [
        ("AMG_type", SHORT), ("print_level", SHORT), ("maxit", INT), ("tol", REAL),
        ("max_levels", SHORT), ("coarse_dof", INT), ("cycle_type", SHORT),
        ("quality_bound", REAL), ("smoother", SHORT), ("smooth_order", SHORT),
        ("presmooth_iter", SHORT), ("postsmooth_iter", SHORT), ("relaxation", REAL),
        ("polynomial_degree", SHORT), ("coarse_solver", SHORT), ("coarse_scaling", SHORT),
        ("amli_degree", SHORT), ("amli_coef", PREAL), ("nl_amli_krylov_type", SHORT),
        ("coarsening_type", SHORT), ("aggregation_type", SHORT),
        ("aggregation_norm_type", SHORT), ("interpolation_type", SHORT),
        ("strong_threshold", REAL), ("max_row_sum", REAL), ("truncation_threshold", REAL),
        ("aggressive_level", INT), ("aggressive_path", INT), ("pair_number", INT),
        ("strong_coupled", REAL), ("max_aggregation", INT), ("tentative_smooth", REAL),
        ("smooth_filter", SHORT), ("smooth_restriction", SHORT), ("ILU_levels", SHORT),
        ("ILU_type", SHORT), ("ILU_lfil", INT), ("ILU_droptol", REAL), ("ILU_relax", REAL),
        ("ILU_permtol", REAL), ("SWZ_levels", INT), ("SWZ_mmsize", INT), ("SWZ_maxlvl", INT),
        ("SWZ_type", INT), ("SWZ_blksolver", INT), ("theta", REAL),
    ]


class Pardiso_data(C.Structure):
    _fields_ = [("pt", C.c_void_p * 64)]


class Mumps_data(C.Structure):
    _fields_ = [("job", INT)]


class ILU_data(C.Structure):
    _fields_ = [("A", C.POINTER(dCSRmat)), ("type", INT), ("row", INT), ("col", INT), ("nzlu", INT),
                ("ijlu", PINT), ("luval", PREAL), ("nb", INT), ("nwork", INT), ("work", PREAL),
                ("iperm", PINT), ("ncolors", INT), ("ic", PINT), ("icmap", PINT), ("uptr", PINT),
                ("nlevL", INT), ("nlevU", INT), ("ilevL", PINT), ("ilevU", PINT), ("jlevL", PINT),
                ("jlevU", PINT)]


class SWZ_data(C.Structure):
    _fields_ = [("A", dCSRmat), ("nblk", INT), ("iblock", PINT), ("jblock", PINT), ("rhsloc", PREAL),
                ("rhsloc1", dvector), ("xloc1", dvector), ("au", PREAL), ("al", PREAL),
                ("SWZ_type", INT), ("blk_solver", INT), ("memt", INT), ("mask", PINT),
                ("maxbs", INT), ("maxa", PINT), ("blk_data", C.POINTER(dCSRmat)),
                ("mumps", C.c_void_p), ("swzparam", C.c_void_p)]


class AMG_data(C.Structure):
    _fields_ = [("max_levels", SHORT), ("num_levels", SHORT), ("A", dCSRmat), ("R", dCSRmat),
                ("P", dCSRmat), ("b", dvector), ("x", dvector), ("Numeric", C.c_void_p),
                ("pdata", Pardiso_data), ("cfmark", ivector), ("ILU_levels", INT), ("LU", ILU_data),
                ("near_kernel_dim", INT), ("near_kernel_basis", C.c_void_p), ("SWZ_levels", INT),
                ("Schwarz", SWZ_data), ("w", dvector), ("mumps", Mumps_data), ("cycle_type", INT),
                ("ic", PINT), ("icmap", PINT), ("colors", INT), ("weight", REAL)]


class AMG_data_bsr(C.Structure):
    _fields_ = [("max_levels", INT), ("num_levels", INT), ("A", dBSRmat), ("R", dBSRmat),
                ("P", dBSRmat), ("b", dvector), ("x", dvector), ("diaginv", dvector), ("Ac", dCSRmat),
                ("Numeric", C.c_void_p), ("pdata", Pardiso_data), ("PP", dCSRmat),
                ("mglP", C.c_void_p), ("TT", dCSRmat), ("mglT", C.c_void_p), ("PT", dBSRmat),
                ("pw", PREAL), ("SS", dBSRmat), ("sw", PREAL), ("diaginv_SS", dvector),
                ("PP_LU", ILU_data), ("cfmark", ivector), ("ILU_levels", INT), ("LU", ILU_data),
                ("near_kernel_dim", INT), ("near_kernel_basis", C.c_void_p), ("A_nk", C.c_void_p),
                ("P_nk", C.c_void_p), ("R_nk", C.c_void_p), ("w", dvector), ("mumps", Mumps_data)]


class precond_data(C.Structure):
    _fields_ = [("AMG_type", SHORT), ("print_level", SHORT), ("maxit", INT), ("max_levels", SHORT),
                ("tol", REAL), ("cycle_type", SHORT), ("smoother", SHORT), ("smooth_order", SHORT),
                ("presmooth_iter", SHORT), ("postsmooth_iter", SHORT), ("relaxation", REAL),
                ("polynomial_degree", SHORT), ("coarsening_type", SHORT), ("coarse_solver", SHORT),
                ("coarse_scaling", SHORT), ("amli_degree", SHORT), ("nl_amli_krylov_type", SHORT),
                ("tentative_smooth", REAL), ("amli_coef", PREAL), ("mgl_data", C.POINTER(AMG_data)),
                ("LU", C.c_void_p), ("A", C.POINTER(dCSRmat)), ("A_nk", C.c_void_p),
                ("P_nk", C.c_void_p), ("R_nk", C.c_void_p), ("r", dvector), ("w", PREAL)]


PRECOND_FCT = C.CFUNCTYPE(None, PREAL, PREAL, C.c_void_p)


class precond(C.Structure):
    _fields_ = [("data", C.c_void_p), ("fct", PRECOND_FCT)]


MXV_FCT = C.CFUNCTYPE(None, C.c_void_p, PREAL, PREAL)


class mxv_matfree(C.Structure):
    """fasp.h:1109-1117"""
    _fields_ = [("data", C.c_void_p), ("fct", MXV_FCT)]


MAT_CSR, MAT_BSR = 1, 2


class fasp_cuda_slab_level(C.Structure):
    """include/fasp_cuda.h: one row-partitioned level handed to fasp_cuda_dist_krylov_amg_create_slabs"""
    _fields_ = [("A", dCSRmat), ("P", dCSRmat), ("R", dCSRmat), ("row_off", PINT), ("n_pext", INT), ("n_rext", INT)]


# ---- numpy <-> struct helpers ------------------------------------------------------------
class CSR:
    """A CSR matrix held as numpy arrays (int32 offsets/columns, float64 values) plus the
    dCSRmat struct that points into them (the arrays keep the memory alive)."""

    def __init__(self, n_rows, n_cols, ia, ja, val):
        self.ia = np.ascontiguousarray(ia, dtype=np.int32)
        self.ja = np.ascontiguousarray(ja, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float64)
        self.shape = (int(n_rows), int(n_cols))
        assert self.ia.size == n_rows + 1 and self.ja.size == self.val.size == self.ia[-1]
        self.struct = dCSRmat(int(n_rows), int(n_cols), int(self.ia[-1]),
                              self.ia.ctypes.data_as(PINT), self.ja.ctypes.data_as(PINT),
                              self.val.ctypes.data_as(PREAL))

    @property
    def nnz(self):
        return int(self.ia[-1])

    def ptr(self):
        p = C.pointer(self.struct)
        p._owner = self   # the pointer keeps the numpy arrays alive (safe with temporaries)
        return p

    def to_scipy(self):
        import scipy.sparse as sp
        # copies: scipy sorts indices / merges duplicates IN PLACE on some operations
        return sp.csr_matrix((self.val.copy(), self.ja.copy(), self.ia.copy()), shape=self.shape)

    @staticmethod
    def from_scipy(m):
        m = m.tocsr()
        return CSR(m.shape[0], m.shape[1], m.indptr, m.indices, m.data)

    @staticmethod
    def from_struct(s: dCSRmat):
        """Copy a dCSRmat owned by C code into numpy arrays."""
        ia = np.ctypeslib.as_array(s.IA, shape=(s.row + 1,)).copy()
        nnz = int(ia[-1])
        ja = np.ctypeslib.as_array(s.JA, shape=(nnz,)).copy() if nnz else np.zeros(0, np.int32)
        val = (np.ctypeslib.as_array(s.val, shape=(nnz,)).copy() if (nnz and s.val)
               else np.ones(nnz))
        return CSR(s.row, s.col, ia, ja, val)


class BSR:
    """Block CSR (nb x nb row-major blocks) as numpy arrays + the dBSRmat struct."""

    def __init__(self, ROW, COL, nb, ia, ja, val):
        self.ia = np.ascontiguousarray(ia, dtype=np.int32)
        self.ja = np.ascontiguousarray(ja, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float64).reshape(-1)
        self.ROW, self.COL, self.nb = int(ROW), int(COL), int(nb)
        NNZ = int(self.ia[-1])
        assert self.ja.size == NNZ and self.val.size == NNZ * nb * nb
        self.struct = dBSRmat(self.ROW, self.COL, NNZ, self.nb, 0, self.val.ctypes.data_as(PREAL),
                              self.ia.ctypes.data_as(PINT), self.ja.ctypes.data_as(PINT))

    @property
    def NNZ(self):
        return int(self.ia[-1])

    def ptr(self):
        p = C.pointer(self.struct)
        p._owner = self   # the pointer keeps the numpy arrays alive (safe with temporaries)
        return p

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.bsr_matrix((self.val.reshape(-1, self.nb, self.nb).copy(), self.ja.copy(), self.ia.copy()),
                             shape=(self.ROW * self.nb, self.COL * self.nb))


class Vec:
    def __init__(self, arr):
        self.a = np.ascontiguousarray(arr, dtype=np.float64)
        self.struct = dvector(int(self.a.size), self.a.ctypes.data_as(PREAL))

    def ptr(self):
        p = C.pointer(self.struct)
        p._owner = self   # the pointer keeps the numpy arrays alive (safe with temporaries)
        return p


def as_preal(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(PREAL)
