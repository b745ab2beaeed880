"""Smoother sweeps vs the sequential reference (bar: 1e-12; exact where rows are short)."""
import numpy as np
import pytest

from faspsolver_b200 import fasp_types as T
from faspsolver_b200 import problems as PB

pytestmark = pytest.mark.gpu


def _mats(data):
    return [("FD", data["FD"]), ("FE", data["FE"]), ("p7_16", PB.poisson7(16)), ("cd7_12", PB.convdiff7(12))]


@pytest.mark.parametrize("w,nsw", [(1.0, 1), (0.67, 3)])
def test_jacobi(gpu, ref, data, w, nsw):
    rng = np.random.default_rng(21)
    for name, A in _mats(data):
        n = A.shape[0]
        b, u0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        u, ur = T.Vec(u0.copy()), T.Vec(u0.copy())
        assert gpu.fasp_cuda_smoother_dcsr_jacobi(u.ptr(), 0, n - 1, 1, A.ptr(), T.Vec(b).ptr(), nsw, w) == 0, \
            gpu.fasp_cuda_last_error()
        ref.L.fasp_smoother_dcsr_jacobi(ur.ptr(), 0, n - 1, 1, A.ptr(), T.Vec(b).ptr(), nsw, w)
        assert np.allclose(u.a, ur.a, rtol=1e-12, atol=1e-12 * np.abs(ur.a).max()), name
        assert np.array_equal(u.a, ur.a), name   # short rows: same order, same rounding
        # backward direction gives the same Jacobi result (ItrSmootherCSR.c:172-228)
        u2 = T.Vec(u0.copy())
        assert gpu.fasp_cuda_smoother_dcsr_jacobi(u2.ptr(), n - 1, 0, -1, A.ptr(), T.Vec(b).ptr(), nsw, w) == 0
        assert np.array_equal(u2.a, u.a)


@pytest.mark.parametrize("nsw", [1, 2])
def test_l1diag(gpu, ref, data, nsw):
    rng = np.random.default_rng(22)
    for name, A in _mats(data):
        n = A.shape[0]
        b, u0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        u, ur = T.Vec(u0.copy()), T.Vec(u0.copy())
        assert gpu.fasp_cuda_smoother_dcsr_L1diag(u.ptr(), 0, n - 1, 1, A.ptr(), T.Vec(b).ptr(), nsw) == 0
        ref.L.fasp_smoother_dcsr_L1diag(ur.ptr(), 0, n - 1, 1, A.ptr(), T.Vec(b).ptr(), nsw)
        assert np.array_equal(u.a, ur.a), name


@pytest.mark.parametrize("ndeg,nsw", [(3, 1), (2, 2), (5, 1)])
def test_poly(gpu, ref, data, ndeg, nsw):
    rng = np.random.default_rng(23)
    for name, A in _mats(data):
        n = A.shape[0]
        b, u0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        u, ur = T.Vec(u0.copy()), T.Vec(u0.copy())
        assert gpu.fasp_cuda_smoother_dcsr_poly(A.ptr(), T.Vec(b).ptr(), u.ptr(), n, ndeg, nsw) == 0, \
            gpu.fasp_cuda_last_error()
        ref.L.fasp_smoother_dcsr_poly(A.ptr(), T.Vec(b).ptr(), ur.ptr(), n, ndeg, nsw)
        scale = np.abs(ur.a).max()
        assert np.allclose(u.a, ur.a, rtol=0, atol=1e-12 * scale), (name, np.abs(u.a - ur.a).max() / scale)


def test_golden_smoother_vectors(gpu, data, golden_vectors):
    A, g, n = data["FE"], golden_vectors, data["FE"].shape[0]
    b = data["FE_b"]
    u = T.Vec(g["x"].copy())
    assert gpu.fasp_cuda_smoother_dcsr_jacobi(u.ptr(), 0, n - 1, 1, A.ptr(), T.Vec(b).ptr(), 2, 0.67) == 0
    assert np.array_equal(u.a, g["smooth_jacobi067"])
    u = T.Vec(g["x"].copy())
    assert gpu.fasp_cuda_smoother_dcsr_L1diag(u.ptr(), 0, n - 1, 1, A.ptr(), T.Vec(b).ptr(), 2) == 0
    assert np.array_equal(u.a, g["smooth_l1diag"])
    u = T.Vec(g["x"].copy())
    assert gpu.fasp_cuda_smoother_dcsr_poly(A.ptr(), T.Vec(b).ptr(), u.ptr(), n, 3, 2) == 0
    assert np.allclose(u.a, g["smooth_poly3"], rtol=0, atol=1e-12 * np.abs(g["smooth_poly3"]).max())
