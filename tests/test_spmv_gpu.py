"""CSR SpMV parity: libfasp_cuda (C ABI) vs the sequential reference, same inputs.
Bar (north star): |y - y_ref|_i <= 1e-14 * (|A||x|)_i ; rows handled by one thread are bit-exact."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from faspsolver_b200 import api
from faspsolver_b200 import fasp_types as T
from faspsolver_b200 import problems as PB
from faspsolver_b200.fasp_types import CSR

pytestmark = pytest.mark.gpu


def _bound(A, x):
    return np.abs(A.to_scipy()) @ np.abs(x)


def _cases(data):
    rng = np.random.default_rng(3)
    out = [("FD", data["FD"]), ("FE", data["FE"]), ("p7_24", PB.poisson7(24)), ("p27_12", PB.poisson27(12)),
           ("cd7_16", PB.convdiff7(16))]
    # irregular rows incl. empty rows, rows of 40..300 and one row longer than any CTA tile
    m = sp.random(700, 900, density=0.02, format="lil", random_state=5)
    m[10, :] = rng.uniform(-1, 1, 900)
    m[11, :300] = rng.uniform(-1, 1, 300)
    m[500:520, :] = 0
    out.append(("irregular", CSR.from_scipy(m.tocsr())))
    big = sp.random(3, 9000, density=0.9, format="csr", random_state=6)
    out.append(("longrow", CSR.from_scipy(big)))
    dense_rows = sp.random(64, 2000, density=0.5, format="csr", random_state=8)
    out.append(("rows1000", CSR.from_scipy(dense_rows)))
    return out


def test_mxv_matches_reference(gpu, ref, data):
    rng = np.random.default_rng(11)
    for name, A in _cases(data):
        x = rng.uniform(-1, 1, A.shape[1])
        y = np.empty(A.shape[0])
        st = gpu.fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(x), T.as_preal(y))
        assert st == 0, gpu.fasp_cuda_last_error()
        yr = ref.mxv(A, x)
        err = np.abs(y - yr)
        assert np.all(err <= 1e-14 * _bound(A, x) + 1e-300), (name, err.max())


def test_mxv_short_rows_bit_exact(gpu, ref, data):
    """Rows reduced by a single thread follow the CPU order and rounding exactly."""
    rng = np.random.default_rng(12)
    # 27-point rows (22.7 / 24.4 entries on average): still one thread per row, 16 gathers in flight
    for name, A in (("FE", data["FE"]), ("p7_24", PB.poisson7(24)), ("cd7_16", PB.convdiff7(16)),
                    ("p27_12", PB.poisson27(12)), ("p27_20", PB.poisson27(20))):
        x = rng.uniform(-1, 1, A.shape[1])
        y = np.empty(A.shape[0])
        assert gpu.fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(x), T.as_preal(y)) == 0
        assert np.array_equal(y, ref.mxv(A, x)), name
        y0 = rng.uniform(-1, 1, A.shape[0])
        y = y0.copy()
        assert gpu.fasp_cuda_blas_dcsr_aAxpy(-1.0, A.ptr(), T.as_preal(x), T.as_preal(y)) == 0
        assert np.array_equal(y, ref.aAxpy(-1.0, A, x, y0)), name


def test_strict_mode_bit_exact_everywhere(gpu, ref, data):
    rng = np.random.default_rng(13)
    gpu.fasp_cuda_set_option(b"strict", 1.0)
    try:
        for name, A in _cases(data):
            x = rng.uniform(-1, 1, A.shape[1])
            y = np.empty(A.shape[0])
            assert gpu.fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(x), T.as_preal(y)) == 0
            assert np.array_equal(y, ref.mxv(A, x)), name
    finally:
        gpu.fasp_cuda_set_option(b"strict", 0.0)


@pytest.mark.parametrize("alpha", [1.0, -1.0, 0.37])
def test_aAxpy_matches_reference(gpu, ref, data, alpha):
    rng = np.random.default_rng(14)
    for name, A in _cases(data):
        x = rng.uniform(-1, 1, A.shape[1])
        y0 = rng.uniform(-1, 1, A.shape[0])
        y = y0.copy()
        assert gpu.fasp_cuda_blas_dcsr_aAxpy(alpha, A.ptr(), T.as_preal(x), T.as_preal(y)) == 0
        yr = ref.aAxpy(alpha, A, x, y0)
        err = np.abs(y - yr)
        assert np.all(err <= 1e-14 * (abs(alpha) * _bound(A, x) + np.abs(y0)) + 1e-300), (name, err.max())


def test_pattern_only_variants(gpu, ref, data):
    rng = np.random.default_rng(15)
    A = data["FE"]
    x = rng.uniform(-1, 1, A.shape[1])
    y, yr = np.empty(A.shape[0]), np.empty(A.shape[0])
    assert gpu.fasp_cuda_blas_dcsr_mxv_agg(A.ptr(), T.as_preal(x), T.as_preal(y)) == 0
    ref.L.fasp_blas_dcsr_mxv_agg(A.ptr(), T.as_preal(x), T.as_preal(yr))
    assert np.array_equal(y, yr)
    y0 = rng.uniform(-1, 1, A.shape[0])
    y, yr = y0.copy(), y0.copy()
    assert gpu.fasp_cuda_blas_dcsr_aAxpy_agg(-0.5, A.ptr(), T.as_preal(x), T.as_preal(y)) == 0
    ref.L.fasp_blas_dcsr_aAxpy_agg(-0.5, A.ptr(), T.as_preal(x), T.as_preal(yr))
    assert np.allclose(y, yr, rtol=0, atol=1e-13)


def test_golden_vectors(gpu, data, golden_vectors):
    """Against the committed reference outputs (no reference library needed)."""
    A, g = data["FE"], golden_vectors
    y = np.empty(A.shape[0])
    assert gpu.fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(np.ascontiguousarray(g["x"])), T.as_preal(y)) == 0
    assert np.array_equal(y, g["mxv"]), (int(np.sum(y != g["mxv"])), float(np.abs(y - g["mxv"]).max()),
                                         np.nonzero(y != g["mxv"])[0][:10])
    y = data["FE_b"].copy()
    assert gpu.fasp_cuda_blas_dcsr_aAxpy(-1.0, A.ptr(), T.as_preal(np.ascontiguousarray(g["x"])), T.as_preal(y)) == 0
    assert np.array_equal(y, g["aAxpy_m1"])


def test_empty_and_tiny(gpu):
    A = CSR(3, 3, [0, 0, 1, 1], [2], [2.5])
    y = np.full(3, 9.0)
    x = np.array([1.0, 2.0, 3.0])
    assert gpu.fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(x), T.as_preal(y)) == 0
    assert np.array_equal(y, [0.0, 7.5, 0.0])


@pytest.mark.parametrize("lpr", [32, 64, 128, 256])
def test_wide_row_kernels(gpu, ref, data, lpr):
    """Rows spread over one warp (32 lanes) or 2 / 4 / 8 warps (csr_wide_kernel): SpMV, aAxpy and an
    L1 sweep on matrices with 1000 ... 8000 nonzeros per row, same bound as every other kernel."""
    rng = np.random.default_rng(21)
    sq = sp.random(300, 300, density=0.6, format="csr", random_state=9) + sp.eye(300) * 40.0
    mats = [A for nm, A in _cases(data) if nm in ("longrow", "rows1000", "irregular")] + [CSR.from_scipy(sq.tocsr())]
    gpu.fasp_cuda_set_option(b"vec_min_avg", 1.0)
    gpu.fasp_cuda_set_option(b"vec_lpr", float(lpr))
    try:
        for A in mats:
            x = rng.uniform(-1, 1, A.shape[1])
            y = np.empty(A.shape[0])
            assert gpu.fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(x), T.as_preal(y)) == 0
            assert np.all(np.abs(y - ref.mxv(A, x)) <= 1e-14 * _bound(A, x) + 1e-300)
            y0 = rng.uniform(-1, 1, A.shape[0])
            y = y0.copy()
            assert gpu.fasp_cuda_blas_dcsr_aAxpy(-1.0, A.ptr(), T.as_preal(x), T.as_preal(y)) == 0
            assert np.all(np.abs(y - ref.aAxpy(-1.0, A, x, y0)) <= 1e-14 * (_bound(A, x) + np.abs(y0)) + 1e-300)
        A = mats[-1]
        n = A.shape[0]
        b, u0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        vb, u, ur = T.Vec(b), T.Vec(u0.copy()), T.Vec(u0.copy())
        assert gpu.fasp_cuda_smoother_dcsr_L1diag(u.ptr(), 0, n - 1, 1, A.ptr(), vb.ptr(), 2) == 0
        ref.L.fasp_smoother_dcsr_L1diag(ur.ptr(), 0, n - 1, 1, A.ptr(), vb.ptr(), 2)
        assert np.abs(u.a - ur.a).max() <= 1e-12 * max(1.0, np.abs(ur.a).max())
    finally:
        gpu.fasp_cuda_set_option(b"vec_lpr", 0.0)
        gpu.fasp_cuda_set_option(b"vec_min_avg", 48.0)


# ---- entry points added last in round 1 (kept at the end of the last GPU test module) -------------

def test_vmv_matches_reference(gpu, ref, data):
    """y'Ax in one fused pass (fasp_blas_dcsr_vmv, BlaSpmvCSR.c:839; used by coarse_scaling)."""
    rng = np.random.default_rng(31)
    for name, A in _cases(data):
        if A.shape[0] != A.shape[1]:
            continue
        x, y = rng.uniform(-1, 1, A.shape[1]), rng.uniform(-1, 1, A.shape[0])
        v = gpu.fasp_cuda_blas_dcsr_vmv(A.ptr(), T.as_preal(x), T.as_preal(y))
        vr = ref.L.fasp_blas_dcsr_vmv(A.ptr(), T.as_preal(x), T.as_preal(y))
        bound = float(np.abs(y) @ _bound(A, x))
        assert abs(v - vr) <= 1e-12 * bound + 1e-300, (name, v, vr)   # reduction order differs


def test_solver_amg_drop_in(gpu, ref, data, golden_answers):
    """fasp_cuda_solver_amg (SolAMG.c:49): setup by the host FASP + device cycles; reg.gcc:412 pins the
    L1_DIAG recipe on the FE problem at 19 iterations / 8.612004e-11."""
    g = golden_answers["reg_gcc"]["FE_amg_solver_L1DIAG_tol1e-10"]
    A, b = data["FE"], data["FE_b"]
    n = A.shape[0]
    amg = ref.amg_param(print_level=0, maxit=500, tol=1e-10, smoother=T.SMOOTHER_L1DIAG)
    vb, vx = T.Vec(b), T.Vec(np.zeros(n))
    st = gpu.fasp_cuda_solver_amg(A.ptr(), vb.ptr(), vx.ptr(), C.byref(amg))
    assert st == g["iters"], (st, gpu.fasp_cuda_last_error())
    relres = np.linalg.norm(b - A.to_scipy() @ vx.a) / np.linalg.norm(b)
    assert abs(relres - g["relres"]) / g["relres"] < 1e-3, relres
    assert np.abs(vx.a - data["FE_sol"]).max() < 1e-4


def test_c_application_three_usages(gpu, c_example_exe):
    """examples/poisson_amg_cuda.c on the GPU: drop-in solve, preconditioner plug-in, one hierarchy with
    several right-hand sides (page-locked application arrays). The program checks the true residual
    of every solve itself and returns 0 only if all meet the tolerance."""
    import subprocess
    for mode in ("0", "1", "2"):
        r = subprocess.run([str(c_example_exe), "24", mode], capture_output=True, text=True)
        assert r.returncode == 0, (mode, r.stdout[-1500:], r.stderr[-1500:])
        assert "iterations, true relative residual" in r.stdout


def test_gs_multicolor_drop_in_matches_oracle(gpu, data):
    """fasp_cuda_smoother_dcsr_gs_multicolor (BlaSparseCSR.c:2123) against the plain-C restatement, which the
    CPU suite pins bit for bit to the OpenMP reference build (tests/test_oracle.py): two sweeps, both
    colour orders."""
    from oracle.port import Oracle, _pd, _pi
    orc = Oracle()
    rng = np.random.default_rng(41)
    for A in (data["FE"], PB.poisson7(12), PB.poisson27(8)):
        n = A.shape[0]
        b, u0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        ic, icmap = np.zeros(n + 2, dtype=np.int32), np.zeros(n, dtype=np.int32)
        ncol = orc.L.oracle_multicolor(n, _pi(A.ia), _pi(A.ja), _pi(ic), _pi(icmap))
        for order in (1, -1):
            want = u0.copy()
            orc.L.oracle_gs_multicolor(n, _pi(A.ia), _pi(A.ja), _pd(A.val), _pd(b), _pd(want), 2, order, ncol,
                                       _pi(ic), _pi(icmap))
            vu, vb = T.Vec(u0.copy()), T.Vec(b)
            assert gpu.fasp_cuda_smoother_dcsr_gs_multicolor(vu.ptr(), A.ptr(), vb.ptr(), 2, order) == 0
            assert np.abs(vu.a - want).max() <= 1e-13 * max(1.0, np.abs(want).max()), order


def test_kernel_scheduling_options_do_not_change_results(gpu, ref, data):
    """two_phase_mask (gather rounds in two enforced phases, per mode) and vec_u (loads in flight per lane of the
    vector kernel) only change instruction scheduling: y = Ax, y += aAx and the smoother sweeps must be bit-identical
    for every setting, and equal to the reference where the reference order is kept (one thread per row)."""
    rng = np.random.default_rng(16)
    mats = [("p7_24", PB.poisson7(24)), ("p27_12", PB.poisson27(12)), ("FE", data["FE"])]
    # a few long rows -> vector kernel
    mats.append(("dense_rows", T.CSR.from_scipy(sp.random(400, 3000, density=0.2, format="csr", random_state=3))))
    settings = [(b"two_phase_mask", 0.0), (b"two_phase_mask", 255.0), (b"two_phase_mask", 6.0), (b"vec_u", 8.0), (b"vec_u", 4.0)]
    try:
        for name, A in mats:
            x = rng.uniform(-1, 1, A.shape[1])
            y0 = rng.uniform(-1, 1, A.shape[0])
            outs = []
            for key, val in settings:
                api.check(gpu.fasp_cuda_set_option(key, val))
                y = np.empty(A.shape[0])
                assert gpu.fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(x), T.as_preal(y)) == 0
                z = y0.copy()
                assert gpu.fasp_cuda_blas_dcsr_aAxpy(-1.0, A.ptr(), T.as_preal(x), T.as_preal(z)) == 0
                outs.append((y, z))
            for y, z in outs[1:]:
                assert np.array_equal(y, outs[0][0]) and np.array_equal(z, outs[0][1]), name
            if name != "dense_rows":
                assert np.array_equal(outs[0][0], ref.mxv(A, x)), name
        # the L1 sweep through the smoother drop-in (square matrices)
        for name, A in mats[:3]:
            b = rng.uniform(-1, 1, A.shape[0])
            res = []
            for key, val in settings[:3]:
                api.check(gpu.fasp_cuda_set_option(key, val))
                u = T.Vec(np.zeros(A.shape[0]))
                vb = T.Vec(b)
                assert gpu.fasp_cuda_smoother_dcsr_L1diag(u.ptr(), 0, A.shape[0] - 1, 1, A.ptr(), vb.ptr(), 2) == 0
                res.append(u.a.copy())
            assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2]), name
    finally:
        gpu.fasp_cuda_set_option(b"two_phase_mask", 6.0)
        gpu.fasp_cuda_set_option(b"vec_u", 0.0)
