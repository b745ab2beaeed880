"""Size-independent properties at (near) BASELINE sizes, where the sequential oracle would take
minutes: linearity of the resident SpMV, agreement of the fused residual with y = Ax, row sums of
the Poisson stencil, smoother fixed point, and a full 128^3 AMG-PCG solve checked by its true
residual and by the reference's known iteration count (12)."""
import ctypes as C

import numpy as np
import pytest

from faspsolver_b200 import api
from faspsolver_b200 import fasp_types as T
from faspsolver_b200 import problems as PB

pytestmark = pytest.mark.gpu


class Dev:
    def __init__(self, L, n):
        self.L, self.n = L, n
        self.p = L.fasp_cuda_dvec_alloc(n)
        assert self.p

    def put(self, a):
        api.check(self.L.fasp_cuda_dvec_h2d(self.p, T.as_preal(np.ascontiguousarray(a, dtype=np.float64)), self.n))
        return self

    def get(self):
        out = np.empty(self.n)
        api.check(self.L.fasp_cuda_dvec_d2h(T.as_preal(out), self.p, self.n))
        return out

    def free(self):
        self.L.fasp_cuda_dvec_free(self.p)


@pytest.mark.parametrize("gen,n", [(PB.poisson7, 160), (PB.poisson27, 96)])
def test_resident_spmv_properties_large(gpu, gen, n):
    A = gen(n)
    N = A.shape[0]
    h = gpu.fasp_cuda_dcsr_upload(A.ptr())
    assert h, gpu.fasp_cuda_last_error()
    rng = np.random.default_rng(51)
    x1, x2 = rng.uniform(-1, 1, N), rng.uniform(-1, 1, N)
    dx, dy, db = Dev(gpu, N), Dev(gpu, N), Dev(gpu, N)
    def mxv(x):
        dx.put(x)
        api.check(gpu.fasp_cuda_dcsr_spmv_dev(h, 0, 1.0, dx.p, None, dy.p))
        return dy.get()
    y1, y2, y12 = mxv(x1), mxv(x2), mxv(2.0 * x1 - 0.5 * x2)
    scale = np.abs(y1).max() + np.abs(y2).max()
    assert np.abs(y12 - (2.0 * y1 - 0.5 * y2)).max() <= 1e-13 * scale          # linearity
    ones = mxv(np.ones(N))                                                       # row sums
    if gen is PB.poisson27:
        assert ones.min() >= -1e-12 and np.isclose(ones.max(), 26 - 7)           # corner row: 7 neighbours
    else:
        s = float((n + 1) ** 2)
        assert np.allclose(np.unique(np.round(ones / s)), [0, 1, 2, 3])           # interior / face / edge / corner
    b = rng.uniform(-1, 1, N)                                                    # fused residual == b - y
    dx.put(x1), db.put(b)
    api.check(gpu.fasp_cuda_dcsr_spmv_dev(h, 2, 1.0, dx.p, db.p, dy.p))
    assert np.array_equal(dy.get(), b - y1)
    dy.put(b)                                                                    # y -= A x  == residual
    api.check(gpu.fasp_cuda_dcsr_spmv_dev(h, 1, -1.0, dx.p, None, dy.p))
    assert np.array_equal(dy.get(), b - y1)
    # L1-Jacobi sweep: the exact solution of A u = b is a fixed point
    u = rng.uniform(-1, 1, N)
    bu = mxv(u)
    dx.put(u), db.put(bu)
    api.check(gpu.fasp_cuda_dcsr_smooth_dev(h, T.SMOOTHER_L1DIAG, 1.0, db.p, dx.p, dy.p))
    assert np.abs(dy.get() - u).max() <= 1e-13
    for d in (dx, dy, db):
        d.free()
    gpu.fasp_cuda_dcsr_free(h)


def test_amg_pcg_128_iterations_and_residual(gpu, ref):
    """7-pt Poisson 128^3 (2.1 M rows): the sequential reference needs 12 iterations and ends at
    relres 1.3856055505e-09; the device solve must agree."""
    A = PB.poisson7(128)
    b = np.ones(A.shape[0])
    amg = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
    it = ref.its_param(itsolver_type=T.SOLVER_CG, tol=1e-8, maxit=500, print_level=0)
    mgl = ref.amg_setup(A, amg)
    try:
        s = api.KrylovAmgSolver(mgl, amg)
        st, x = s.solve(b, np.zeros_like(b), it)
        relres_reported = s.stat(1)
        hist = s.history()
        s.close()
    finally:
        ref.amg_free(mgl, amg)
    assert st == 12, st
    assert abs(relres_reported - 1.3856055505e-09) / 1.3856055505e-09 < 1e-6
    assert hist.size == 13 and np.all(np.diff(hist[1:]) < 0) and abs(hist[-1] - relres_reported) < 1e-5 * relres_reported
    assert np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b) <= 1e-8


def test_repeated_host_solves_with_fresh_and_pinned_buffers(gpu, ref):
    """The host-pointer solve must not depend on the identity of the caller's arrays: every call gets
    newly allocated (mmap-sized) numpy arrays that are released afterwards, so a remembered
    page-lock registration would go stale. Then the same with arrays the application pinned itself
    through fasp_cuda_host_pin (DMA path), and after unpinning them again."""
    A = PB.poisson7(48)
    n = A.shape[0]
    S = A.to_scipy()
    amg = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
    it = ref.its_param(itsolver_type=T.SOLVER_CG, tol=1e-8, maxit=100, print_level=0)
    mgl = ref.amg_setup(A, amg)
    try:
        s = api.KrylovAmgSolver(mgl, amg)
        rng = np.random.default_rng(5)
        counts = []
        for k in range(6):
            b = rng.uniform(0.5, 1.5, n)
            junk = [np.empty(n) for _ in range(k % 3)]      # perturb the allocator between calls
            st, x = s.solve(b, np.zeros(n), it)
            assert st > 0, (k, st, api.last_error())
            assert np.linalg.norm(b - S @ x) / np.linalg.norm(b) <= 1e-8 * 1.001, k
            counts.append(st)
            del b, x, junk
        assert max(counts) - min(counts) <= 1
        b, out = rng.uniform(0.5, 1.5, n), np.zeros(n)
        api.pin_host(b); api.pin_host(out)
        api.pin_host(out)                                    # pinning twice is not an error
        for _ in range(2):
            st, x = s.solve(b, np.zeros(n), it, out=out)
            assert st > 0 and x is out
            assert np.linalg.norm(b - S @ out) / np.linalg.norm(b) <= 1e-8 * 1.001
        api.unpin_host(b); api.unpin_host(out)
        st, x = s.solve(b, np.zeros(n), it, out=out)
        assert st > 0 and np.linalg.norm(b - S @ out) / np.linalg.norm(b) <= 1e-8 * 1.001
        with pytest.raises(api.FaspCudaError):
            api.unpin_host(out)                              # not pinned any more
        s.close()
    finally:
        ref.amg_free(mgl, amg)


def test_amg_pcg_27pt_64_matches_sequential_reference(gpu, ref):
    """BASELINE configs[2]'s operator (27-point, diag 26 / off -1) at a size the sequential oracle finishes in
    seconds: same iteration count (+-1), true residual <= tol, solution within 1e-8 of the reference's."""
    A = PB.poisson27(64)
    n = A.shape[0]
    b = np.ones(n)
    it = ref.its_param(itsolver_type=T.SOLVER_CG, tol=1e-8, maxit=100, print_level=0)
    amg = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
    st, x = api.fasp_cuda_solver_dcsr_krylov_amg(A, b, np.zeros(n), it, amg)
    amg_r = ref.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
    st_ref, x_ref = ref.krylov_amg(A, b, np.zeros(n), it, amg_r)
    assert st > 0 and abs(st - st_ref) <= 1, (st, st_ref, api.last_error())
    assert np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b) <= 1e-8
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) <= 1e-8
