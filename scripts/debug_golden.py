import sys; sys.path.insert(0, '.')
import numpy as np
from faspsolver_b200 import api, fasp_types as T
from faspsolver_b200.fasp_types import CSR
from oracle.ref import RefFasp
L = api.lib(); L.fasp_cuda_init(0)
z = np.load('tests/golden/fasp_data.npz'); g = np.load('tests/golden/oracle_vectors.npz')
n = z['FE_ia'].size - 1
A = CSR(n, n, z['FE_ia'], z['FE_ja'], z['FE_val'])
x = np.ascontiguousarray(g['x'])
y = np.empty(n)
L.fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(x), T.as_preal(y))
ref = RefFasp()
yr = ref.mxv(A, x)
print("gpu vs golden mismatches:", np.sum(y != g['mxv']), "gpu vs live ref:", np.sum(y != yr), "live ref vs golden:", np.sum(yr != g['mxv']))
bad = np.nonzero(y != yr)[0]
print(bad[:20], (y - yr)[bad[:10]])
rowlen = np.diff(A.ia)
print("rowlen stats", rowlen.min(), rowlen.max(), np.bincount(rowlen))
# --- replicate pytest order
import scipy.sparse as sp
from faspsolver_b200 import problems as PB
rng = np.random.default_rng(3)
for rep in range(3):
    for A2 in (PB.poisson7(24), PB.poisson27(12), CSR.from_scipy(sp.random(3, 9000, density=0.9, format="csr", random_state=6))):
        x2 = rng.uniform(-1, 1, A2.shape[1]); y2 = np.empty(A2.shape[0])
        L.fasp_cuda_blas_dcsr_mxv(A2.ptr(), T.as_preal(x2), T.as_preal(y2))
    y = np.empty(n)
    L.fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(x), T.as_preal(y))
    print("rep", rep, "mismatch vs golden", np.sum(y != g['mxv']), np.abs(y - g['mxv']).max())
L.fasp_cuda_set_option(b"strict", 1.0); L.fasp_cuda_set_option(b"strict", 0.0)
y = np.empty(n); L.fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(x), T.as_preal(y))
print("after strict toggle", np.sum(y != g['mxv']))
yy = np.empty(n); L.fasp_cuda_blas_dcsr_mxv_agg(A.ptr(), T.as_preal(x), T.as_preal(yy))
y = np.empty(n); L.fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(x), T.as_preal(y))
print("after agg", np.sum(y != g['mxv']))
y0 = rng.uniform(-1, 1, n); yy = y0.copy(); L.fasp_cuda_blas_dcsr_aAxpy_agg(-0.5, A.ptr(), T.as_preal(x), T.as_preal(yy))
y = np.empty(n); L.fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(x), T.as_preal(y))
print("after aAxpy_agg", np.sum(y != g['mxv']), np.abs(y - g['mxv']).max())
