"""Python host-side mirror of the reference interface over the C-ABI of libfasp_cuda.

Function names, argument order and return conventions follow FASP's C API (the reference is
compiled code; this module exists so that tests and benchmarks read like the reference's own
drivers). Everything numerical happens in libfasp_cuda.so; there is no Python or CPU
fallback: if the library or a CUDA device is missing the calls raise / return FASP error
codes.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from . import fasp_types as T
from .fasp_types import (AMG_data, AMG_data_bsr, AMG_param, BSR, CSR, ILU_param, ITS_param, Vec,
                         dBSRmat, dCSRmat, dvector, precond, INT, REAL, SHORT, PREAL)

_PKG = Path(__file__).resolve().parent
_LIB_PATH = _PKG / "lib" / "libfasp_cuda.so"
_lib = None


class FaspCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libfasp_cuda status %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load libfasp_cuda.so (built in-tree by faspsolver_b200.build). Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise FileNotFoundError(
            "%s is missing: run `python -m faspsolver_b200.build` (there is no fallback path)" % _LIB_PATH)
    L = C.CDLL(str(_LIB_PATH), mode=C.RTLD_GLOBAL)
    P = C.POINTER
    vp = C.c_void_p
    sig = {
        "fasp_cuda_abi_check": (INT, [C.c_size_t, C.c_size_t, C.c_size_t]),
        "fasp_cuda_last_error": (C.c_char_p, []),
        "fasp_cuda_init": (INT, [C.c_int]),
        "fasp_cuda_launch_count": (C.c_longlong, []),
        "fasp_cuda_launch_count_reset": (None, []),
        "fasp_cuda_set_option": (INT, [C.c_char_p, C.c_double]),
        "fasp_cuda_profile_dump": (C.c_longlong, [C.c_char_p, C.c_longlong]),
        "fasp_cuda_get_option": (C.c_double, [C.c_char_p]),
        "fasp_cuda_blas_dcsr_mxv": (INT, [P(dCSRmat), PREAL, PREAL]),
        "fasp_cuda_blas_dcsr_aAxpy": (INT, [REAL, P(dCSRmat), PREAL, PREAL]),
        "fasp_cuda_blas_dcsr_vmv": (REAL, [P(dCSRmat), PREAL, PREAL]),
        "fasp_cuda_blas_dcsr_mxv_agg": (INT, [P(dCSRmat), PREAL, PREAL]),
        "fasp_cuda_blas_dcsr_aAxpy_agg": (INT, [REAL, P(dCSRmat), PREAL, PREAL]),
        "fasp_cuda_blas_dbsr_mxv": (INT, [P(dBSRmat), PREAL, PREAL]),
        "fasp_cuda_blas_dbsr_aAxpy": (INT, [REAL, P(dBSRmat), PREAL, PREAL]),
        "fasp_cuda_blas_darray_ax": (INT, [INT, REAL, PREAL]),
        "fasp_cuda_blas_darray_axpy": (INT, [INT, REAL, PREAL, PREAL]),
        "fasp_cuda_blas_darray_axpby": (INT, [INT, REAL, PREAL, REAL, PREAL]),
        "fasp_cuda_blas_darray_dotprod": (REAL, [INT, PREAL, PREAL]),
        "fasp_cuda_blas_darray_norm2": (REAL, [INT, PREAL]),
        "fasp_cuda_blas_darray_norm1": (REAL, [INT, PREAL]),
        "fasp_cuda_blas_darray_norminf": (REAL, [INT, PREAL]),
        "fasp_cuda_solver_matfree_init": (INT, [INT, P(T.mxv_matfree), vp]),
        "fasp_cuda_dense_inverse": (INT, [INT, PREAL, PREAL]),
        "fasp_cuda_dcsr_trans": (INT, [P(dCSRmat), P(dCSRmat)]),
        "fasp_cuda_blas_dcsr_rap": (INT, [P(dCSRmat), P(dCSRmat), P(dCSRmat), P(dCSRmat)]),
        "fasp_cuda_blas_mxv_csr": (None, [vp, PREAL, PREAL]),
        "fasp_cuda_blas_mxv_bsr": (None, [vp, PREAL, PREAL]),
        "fasp_cuda_smoother_dcsr_jacobi": (INT, [P(dvector), INT, INT, INT, P(dCSRmat), P(dvector), INT, REAL]),
        "fasp_cuda_smoother_dcsr_L1diag": (INT, [P(dvector), INT, INT, INT, P(dCSRmat), P(dvector), INT]),
        "fasp_cuda_smoother_dcsr_poly": (INT, [P(dCSRmat), P(dvector), P(dvector), INT, INT, INT]),
        "fasp_cuda_smoother_dcsr_gs_multicolor": (INT, [P(dvector), P(dCSRmat), P(dvector), INT, INT]),
        "fasp_cuda_multicolor_host": (INT, [INT, T.PINT, T.PINT, T.PINT, T.PINT]),
        "fasp_cuda_smoother_dbsr_jacobi1": (INT, [P(dBSRmat), P(dvector), P(dvector), PREAL]),
        "fasp_cuda_dcsr_upload": (vp, [P(dCSRmat)]),
        "fasp_cuda_dcsr_free": (None, [vp]),
        "fasp_cuda_dbsr_upload": (vp, [P(dBSRmat)]),
        "fasp_cuda_dbsr_free": (None, [vp]),
        "fasp_cuda_dvec_alloc": (vp, [C.c_size_t]),
        "fasp_cuda_dvec_free": (None, [vp]),
        "fasp_cuda_dvec_h2d": (INT, [vp, PREAL, C.c_size_t]),
        "fasp_cuda_dvec_d2h": (INT, [PREAL, vp, C.c_size_t]),
        "fasp_cuda_sync": (INT, []),
        "fasp_cuda_dcsr_spmv_dev": (INT, [vp, C.c_int, REAL, vp, vp, vp]),
        "fasp_cuda_dbsr_spmv_dev": (INT, [vp, C.c_int, REAL, vp, vp, vp]),
        "fasp_cuda_dcsr_smooth_dev": (INT, [vp, C.c_int, REAL, vp, vp, vp]),
        "fasp_cuda_dcsr_time_kernel": (C.c_double, [vp, C.c_int, C.c_int, C.c_int, C.c_int]),
        "fasp_cuda_dbsr_time_kernel": (C.c_double, [vp, C.c_int, C.c_int, C.c_int, C.c_int]),
        "fasp_cuda_amg_upload": (vp, [P(AMG_data), P(AMG_param)]),
        "fasp_cuda_amg_free": (None, [vp]),
        "fasp_cuda_amg_bytes": (C.c_size_t, [vp]),
        "fasp_cuda_amg_levels": (INT, [vp]),
        "fasp_cuda_bamg_upload": (vp, [P(AMG_data_bsr), P(AMG_param)]),
        "fasp_cuda_bamg_free": (None, [vp]),
        "fasp_cuda_solver_mgcycle": (INT, [P(AMG_data), P(AMG_param)]),
        "fasp_cuda_solver_mgcycle_bsr": (INT, [P(AMG_data_bsr), P(AMG_param)]),
        "fasp_cuda_amg_cycle_dev": (INT, [vp, vp, vp]),
        "fasp_cuda_amg_cycle_host": (INT, [vp, PREAL, PREAL]),
        "fasp_cuda_precond_amg": (None, [PREAL, PREAL, vp]),
        "fasp_cuda_precond_setup": (P(precond), [SHORT, P(AMG_param), P(ILU_param), P(dCSRmat)]),
        "fasp_cuda_precond_free": (None, [P(precond)]),
        "fasp_cuda_precond_from_mgl": (P(precond), [P(AMG_data), P(AMG_param)]),
        "fasp_cuda_solver_dcsr_pcg": (INT, [P(dCSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT]),
        "fasp_cuda_solver_dcsr_pvgmres": (INT, [P(dCSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT, SHORT]),
        "fasp_cuda_solver_dcsr_pgmres": (INT, [P(dCSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT, SHORT]),
        "fasp_cuda_solver_dcsr_pvfgmres": (INT, [P(dCSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT, SHORT]),
        "fasp_cuda_solver_dbsr_pvfgmres": (INT, [P(dBSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT, SHORT]),
        "fasp_cuda_solver_dbsr_pcg": (INT, [P(dBSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT]),
        "fasp_cuda_solver_dbsr_pvgmres": (INT, [P(dBSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT, SHORT]),
        "fasp_cuda_solver_dbsr_pgmres": (INT, [P(dBSRmat), P(dvector), P(dvector), P(precond), REAL, REAL, INT, SHORT, SHORT, SHORT]),
        "fasp_cuda_host_pin": (INT, [vp, C.c_size_t]),
        "fasp_cuda_host_unpin": (INT, [vp]),
        "fasp_cuda_solver_dcsr_itsolver": (INT, [P(dCSRmat), P(dvector), P(dvector), P(precond), P(ITS_param)]),
        "fasp_cuda_solver_dbsr_itsolver": (INT, [P(dBSRmat), P(dvector), P(dvector), P(precond), P(ITS_param)]),
        "fasp_cuda_solver_dcsr_krylov_amg": (INT, [P(dCSRmat), P(dvector), P(dvector), P(ITS_param), P(AMG_param)]),
        "fasp_cuda_solver_dbsr_krylov_amg": (INT, [P(dBSRmat), P(dvector), P(dvector), P(ITS_param), P(AMG_param)]),
        "fasp_cuda_krylov_amg_create": (vp, [P(AMG_data), P(AMG_param)]),
        "fasp_cuda_krylov_bamg_create": (vp, [P(AMG_data_bsr), P(AMG_param)]),
        "fasp_cuda_krylov_amg_solve": (INT, [vp, P(dvector), P(dvector), P(ITS_param)]),
        "fasp_cuda_krylov_amg_solve_dev": (INT, [vp, vp, vp, P(ITS_param)]),
        "fasp_cuda_krylov_amg_destroy": (None, [vp]),
        "fasp_cuda_solver_stat": (C.c_double, [vp, C.c_int]),
        "fasp_cuda_solver_history": (INT, [vp, PREAL, INT]),
        "fasp_cuda_amg_solve": (INT, [P(AMG_data), P(AMG_param)]),
        "fasp_cuda_solver_amg": (INT, [P(dCSRmat), P(dvector), P(dvector), P(AMG_param)]),
        "fasp_cuda_comm_unique_id": (INT, [vp]),
        "fasp_cuda_comm_init": (INT, [vp, C.c_int, C.c_int]),
        "fasp_cuda_comm_finalize": (INT, []),
        "fasp_cuda_comm_rank": (C.c_int, []),
        "fasp_cuda_comm_size": (C.c_int, []),
        "fasp_cuda_comm_peer_memory": (C.c_int, []),
        "fasp_cuda_dist_krylov_amg_create": (vp, [P(AMG_data), P(AMG_param), INT]),
        "fasp_cuda_dist_krylov_amg_create_slabs": (vp, [INT, P(T.fasp_cuda_slab_level), T.PINT, P(AMG_data), P(AMG_param)]),
        "fasp_cuda_dist_row_range": (INT, [vp, T.PINT, T.PINT]),
        "fasp_cuda_dist_extract_host": (INT, [P(dCSRmat), INT, INT, T.PINT, T.PINT, T.PINT, INT, T.PINT, T.PINT, INT, T.PINT]),
    }
    missing = []
    for name, (res, args) in sig.items():
        try:
            f = getattr(L, name)
        except AttributeError:  # tests/test_boundary.py::test_library_exports_every_declared_symbol fails on it
            missing.append(name)
            continue
        f.restype = res
        f.argtypes = args
    L._signatures = sig
    L._missing = missing
    _lib = L
    return L


EXPORTS = None  # filled lazily by declared_symbols()


def declared_symbols():
    """Names declared in include/fasp_cuda.h (parsed from the header)."""
    import re
    text = (_PKG.parent / "include" / "fasp_cuda.h").read_text()
    return sorted(set(re.findall(r"\b(fasp_cuda_[a-zA-Z0-9_]+)\s*\(", text)))


def last_error() -> str:
    return lib().fasp_cuda_last_error().decode()


def check(status: int) -> int:
    if status < 0:
        raise FaspCudaError(status, last_error())
    return status


def pin_host(a: np.ndarray) -> None:
    """Page-lock a long-lived application array (b / x of repeated solves): the host-pointer solve
    then copies it by DMA. Call unpin_host before the array is released."""
    check(lib().fasp_cuda_host_pin(a.ctypes.data, a.nbytes))


def unpin_host(a: np.ndarray) -> None:
    check(lib().fasp_cuda_host_unpin(a.ctypes.data))


# ---------------------------------------------------------------------------------------
# The host application's FASP: parameter initialisers and the AMG setup phase. FASP's own
# host code builds the hierarchy (north star); this class only binds it.
# ---------------------------------------------------------------------------------------
class HostFasp:
    def __init__(self, path: str | os.PathLike | None = None):
        path = path or os.environ.get("FASP_CUDA_HOST_LIBFASP")
        if not path:
            raise FileNotFoundError("set FASP_CUDA_HOST_LIBFASP to the host application's libfasp.so")
        self.path = str(path)
        os.environ.setdefault("FASP_CUDA_HOST_LIBFASP", self.path)
        L = C.CDLL(self.path, mode=C.RTLD_GLOBAL)
        P = C.POINTER
        L.fasp_param_amg_init.argtypes = [P(AMG_param)]
        L.fasp_param_amg_init.restype = None
        L.fasp_param_solver_init.argtypes = [P(ITS_param)]
        L.fasp_param_solver_init.restype = None
        L.fasp_amg_data_create.argtypes = [SHORT]
        L.fasp_amg_data_create.restype = P(AMG_data)
        L.fasp_amg_data_free.argtypes = [P(AMG_data), P(AMG_param)]
        L.fasp_amg_data_free.restype = None
        for nm in ("fasp_amg_setup_rs", "fasp_amg_setup_sa", "fasp_amg_setup_ua"):
            getattr(L, nm).argtypes = [P(AMG_data), P(AMG_param)]
            getattr(L, nm).restype = SHORT
        L.fasp_dcsr_create.argtypes = [INT, INT, INT]
        L.fasp_dcsr_create.restype = dCSRmat
        L.fasp_dcsr_cp.argtypes = [P(dCSRmat), P(dCSRmat)]
        L.fasp_dcsr_cp.restype = None
        L.fasp_dvec_create.argtypes = [INT]
        L.fasp_dvec_create.restype = dvector
        L.fasp_amg_data_bsr_create.argtypes = [SHORT]
        L.fasp_amg_data_bsr_create.restype = P(AMG_data_bsr)
        L.fasp_amg_data_bsr_free.argtypes = [P(AMG_data_bsr), P(AMG_param)]
        L.fasp_amg_data_bsr_free.restype = None
        for nm in ("fasp_amg_setup_ua_bsr", "fasp_amg_setup_sa_bsr"):
            getattr(L, nm).argtypes = [P(AMG_data_bsr), P(AMG_param)]
            getattr(L, nm).restype = SHORT
        L.fasp_dbsr_create.argtypes = [INT, INT, INT, INT, INT]
        L.fasp_dbsr_create.restype = dBSRmat
        L.fasp_dbsr_cp.argtypes = [P(dBSRmat), P(dBSRmat)]
        L.fasp_dbsr_cp.restype = None
        self.L = L

    # -- parameters with FASP's own defaults (AuxParam.c:431-489, 572-583)
    def amg_param(self, **kw) -> AMG_param:
        p = AMG_param()
        self.L.fasp_param_amg_init(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def its_param(self, **kw) -> ITS_param:
        p = ITS_param()
        self.L.fasp_param_solver_init(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    # -- setup phase (SolCSR.c:500-521)
    def amg_setup(self, A: CSR, amgparam: AMG_param):
        L = self.L
        mgl = L.fasp_amg_data_create(amgparam.max_levels)
        mgl[0].A = L.fasp_dcsr_create(A.shape[0], A.shape[1], A.nnz)
        L.fasp_dcsr_cp(A.ptr(), C.byref(mgl[0].A))
        mgl[0].b = L.fasp_dvec_create(A.shape[1])
        mgl[0].x = L.fasp_dvec_create(A.shape[1])
        fn = {T.SA_AMG: L.fasp_amg_setup_sa, T.UA_AMG: L.fasp_amg_setup_ua}.get(
            amgparam.AMG_type, L.fasp_amg_setup_rs)
        st = fn(mgl, C.byref(amgparam))
        if st < 0:
            L.fasp_amg_data_free(mgl, C.byref(amgparam))
            raise RuntimeError("host AMG setup failed: %d" % st)
        return mgl

    def amg_free(self, mgl, amgparam):
        self.L.fasp_amg_data_free(mgl, C.byref(amgparam))

    def bamg_setup(self, A: BSR, amgparam: AMG_param):
        L = self.L
        mgl = L.fasp_amg_data_bsr_create(amgparam.max_levels)
        mgl[0].A = L.fasp_dbsr_create(A.ROW, A.COL, A.NNZ, A.nb, 0)
        L.fasp_dbsr_cp(A.ptr(), C.byref(mgl[0].A))
        mgl[0].b = L.fasp_dvec_create(A.ROW * A.nb)
        mgl[0].x = L.fasp_dvec_create(A.ROW * A.nb)
        fn = L.fasp_amg_setup_sa_bsr if amgparam.AMG_type == T.SA_AMG else L.fasp_amg_setup_ua_bsr
        st = fn(mgl, C.byref(amgparam))
        if st < 0:
            L.fasp_amg_data_bsr_free(mgl, C.byref(amgparam))
            raise RuntimeError("host BSR AMG setup failed: %d" % st)
        return mgl

    def bamg_free(self, mgl, amgparam):
        self.L.fasp_amg_data_bsr_free(mgl, C.byref(amgparam))


def hierarchy_info(mgl):
    """(rows, nnz(A), nnz(P)) per level of a host hierarchy."""
    nl = mgl[0].num_levels
    return [(mgl[l].A.row, mgl[l].A.nnz, mgl[l].P.nnz if l < nl - 1 else 0) for l in range(nl)]


# ---------------------------------------------------------------------------------------
# Thin mirrors of the reference functions (same names with the fasp_cuda_ prefix)
# ---------------------------------------------------------------------------------------
def fasp_cuda_blas_dcsr_mxv(A: CSR, x: np.ndarray) -> np.ndarray:
    y = np.empty(A.shape[0])
    check(lib().fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(np.ascontiguousarray(x)), T.as_preal(y)))
    return y


def fasp_cuda_blas_dcsr_aAxpy(alpha: float, A: CSR, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    y = np.array(y, dtype=np.float64, copy=True)
    check(lib().fasp_cuda_blas_dcsr_aAxpy(alpha, A.ptr(), T.as_preal(np.ascontiguousarray(x)), T.as_preal(y)))
    return y


def fasp_cuda_solver_dcsr_krylov_amg(A: CSR, b: np.ndarray, x: np.ndarray, itparam: ITS_param,
                                     amgparam: AMG_param):
    """Drop-in for fasp_solver_dcsr_krylov_amg (SolCSR.c:476). Returns (status, x)."""
    vb, vx = Vec(b), Vec(np.array(x, dtype=np.float64, copy=True))
    st = lib().fasp_cuda_solver_dcsr_krylov_amg(A.ptr(), vb.ptr(), vx.ptr(), C.byref(itparam),
                                                C.byref(amgparam))
    return st, vx.a


class KrylovAmgSolver:
    """Hierarchy uploaded once, many solves (fasp_cuda_krylov_amg_create/solve/destroy)."""

    def __init__(self, mgl, amgparam: AMG_param, bsr: bool = False):
        L = lib()
        self.h = (L.fasp_cuda_krylov_bamg_create if bsr else L.fasp_cuda_krylov_amg_create)(
            mgl, C.byref(amgparam))
        if not self.h:
            raise FaspCudaError(-1, last_error())

    def solve(self, b: np.ndarray, x0: np.ndarray, itparam: ITS_param, out: np.ndarray | None = None):
        """x0 is not modified; the solution is returned in `out` when given (an application that solves
        many systems re-uses its arrays, which lets the library keep them page-locked)."""
        if out is None:
            out = np.array(x0, dtype=np.float64, copy=True)
        else:
            out[:] = x0
        vb, vx = Vec(b), Vec(out)
        st = lib().fasp_cuda_krylov_amg_solve(self.h, vb.ptr(), vx.ptr(), C.byref(itparam))
        return st, vx.a

    def solve_dev(self, b_dev, x_dev, itparam: ITS_param):
        return lib().fasp_cuda_krylov_amg_solve_dev(self.h, b_dev, x_dev, C.byref(itparam))

    def stat(self, what: int) -> float:
        return lib().fasp_cuda_solver_stat(self.h, what)

    def history(self, n=4096) -> np.ndarray:
        buf = np.zeros(n)
        k = lib().fasp_cuda_solver_history(self.h, T.as_preal(buf), n)
        return buf[:k]

    def close(self):
        if self.h:
            lib().fasp_cuda_krylov_amg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
