// dense.cu — coarsest-level direct solve: a dense inverse built once on the device by
// Gauss-Jordan elimination with partial pivoting, applied per cycle as one GEMV.
//
// Replaces fasp_coarse_itsolver (PreMGUtil.inl:37-58: safeguarded CG to 1e-10, ~20-25
// iterations of ~10 tiny dependent kernels each) on the path; the coarsest matrices FASP
// produces are small and 25-30 % dense (SURVEY.md finding 8). The result differs from the
// reference's iterative coarse solve by its tolerance (<= 1e-10 relative).
#include "amg.cuh"
#include "reduce.cuh"

namespace fc {

__global__ void k_csr_to_dense(int n, const int* ia, const int* ja, const double* val, double* a)
{
    const int i = blockIdx.x;
    for (int k = ia[i] + threadIdx.x; k < ia[i + 1]; k += blockDim.x)
        atomicAdd(&a[(size_t)i * n + ja[k]], val ? val[k] : 1.0);   // duplicates add up
}

// pivot search in column k, rows k..n-1 (one CTA)
__global__ void k_gj_pivot(int n, int k, const double* a, int* piv, double* pivval)
{
    __shared__ double sv[256];
    __shared__ int    si[256];
    double best = -1.0;
    int    bi   = k;
    for (int i = k + threadIdx.x; i < n; i += blockDim.x) {
        const double v = fabs(a[(size_t)i * n + k]);
        if (v > best) {
            best = v;
            bi   = i;
        }
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            const double o = sv[threadIdx.x + off];
            const int    oi = si[threadIdx.x + off];
            if (o > sv[threadIdx.x] || (o == sv[threadIdx.x] && oi < si[threadIdx.x])) {
                sv[threadIdx.x] = o;
                si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        piv[k]  = si[0];
        *pivval = a[(size_t)si[0] * n + k];
    }
}

// swap rows k and piv[k]; scale row k by 1/pivot (pivot position becomes 1/pivot);
// save the scaled row in rowk
__global__ void k_gj_swap_scale(int n, int k, double* a, const int* piv, const double* pivval,
                                double* rowk)
{
    const int p = piv[k];
    const double pv = *pivval;
    const double ip = 1.0 / pv;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        double vk = a[(size_t)k * n + j];
        double vp = a[(size_t)p * n + j];
        if (p != k) a[(size_t)p * n + j] = vk;
        double r = (j == k) ? ip : vp * ip;
        rowk[j]  = r;
    }
}

// column k factors are read before they are overwritten: colk[i] = a[i][k] (i != k)
__global__ void k_gj_col(int n, int k, const double* a, double* colk)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) colk[i] = (i == k) ? 0.0 : a[(size_t)i * n + k];
}

// a[i][j] = (j==k ? 0 : a[i][j]) - colk[i]*rowk[j] for i != k ; row k = rowk
__global__ void k_gj_update(int n, int k, double* a, const double* rowk, const double* colk)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double rj = rowk[j];
    const int i0 = blockIdx.y * 16;
#pragma unroll 4
    for (int ii = 0; ii < 16; ++ii) {
        const int i = i0 + ii;
        if (i >= n) break;
        double* p = a + (size_t)i * n + j;
        if (i == k) {
            *p = rj;
        } else {
            const double base = (j == k) ? 0.0 : *p;
            *p                = base - colk[i] * rj;
        }
    }
}

// undo the row interchanges as column interchanges, in reverse order
__global__ void k_gj_unpermute(int n, double* a, const int* piv)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* row = a + (size_t)i * n;
    for (int k = n - 1; k >= 0; --k) {
        const int p = piv[k];
        if (p != k) {
            const double t = row[k];
            row[k]         = row[p];
            row[p]         = t;
        }
    }
}

static void gauss_jordan(int n, double* a)
{
    int*    piv  = dalloc<int>(n);
    double* rowk = dalloc<double>(n);
    double* colk = dalloc<double>(n);
    double* pivval = dalloc<double>(1);
    const int tb = 256;
    dim3 ug((n + tb - 1) / tb, (n + 15) / 16);
    for (int k = 0; k < n; ++k) {
        FC_LAUNCH(k_gj_pivot, 1, 256, 0, n, k, a, piv, pivval);
        // after this kernel row piv[k] holds the old row k; row k itself is rewritten from
        // rowk by k_gj_update, so its stale contents are never read again
        FC_LAUNCH(k_gj_swap_scale, (n + tb - 1) / tb, tb, 0, n, k, a, piv, pivval, rowk);
        FC_LAUNCH(k_gj_col, (n + tb - 1) / tb, tb, 0, n, k, a, colk);
        FC_LAUNCH(k_gj_update, ug, tb, 0, n, k, a, rowk, colk);
    }
    FC_LAUNCH(k_gj_unpermute, (n + tb - 1) / tb, tb, 0, n, a, piv);
    FC_CUDA(cudaStreamSynchronize(ctx().stream));
    dfree(piv);
    dfree(rowk);
    dfree(colk);
    dfree(pivval);
}

void dense_invert_csr(DenseInv& D, const DevCSR& A)
{
    dense_free(D);
    const int n = A.rows;
    if (A.cols != n) fail(ERROR_MAT_SIZE, "coarsest matrix is not square");
    D.n    = n;
    D.ainv = dalloc<double>((size_t)n * n);
    FC_CUDA(cudaMemsetAsync(D.ainv, 0, sizeof(double) * (size_t)n * n, ctx().stream));
    FC_LAUNCH(k_csr_to_dense, n, 128, 0, n, A.ia, A.ja, A.val, D.ainv);
    gauss_jordan(n, D.ainv);
}

void dense_invert_bsr(DenseInv& D, const DevBSR& A)
{
    dense_free(D);
    if (A.ROW != A.COL) fail(ERROR_MAT_SIZE, "coarsest BSR matrix is not square");
    const int n = A.ROW * A.nb;
    D.n         = n;
    D.ainv      = dalloc<double>((size_t)n * n);
    FC_CUDA(cudaMemsetAsync(D.ainv, 0, sizeof(double) * (size_t)n * n, ctx().stream));
    bsr_to_dense(A, D.ainv);
    gauss_jordan(n, D.ainv);
}

void dense_invert_host(DenseInv& D, int n, const std::vector<double>& a)
{
    dense_free(D);
    D.n    = n;
    D.ainv = dalloc<double>((size_t)n * n);
    FC_CUDA(cudaMemcpyAsync(D.ainv, a.data(), sizeof(double) * (size_t)n * n,
                            cudaMemcpyHostToDevice, ctx().stream));
    FC_CUDA(cudaStreamSynchronize(ctx().stream));
    gauss_jordan(n, D.ainv);
}

// x = Ainv b : one warp per row, coalesced along the row
__global__ void __launch_bounds__(256)
k_dense_gemv(int n, const double* __restrict__ a, const double* __restrict__ b,
             double* __restrict__ x, const int* done)
{
    if (done && *done) return;
    const int lane = threadIdx.x & 31;
    const int row  = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const double* ar = a + (size_t)row * n;
    double        s  = 0.0;
    for (int j = lane; j < n; j += 32) s += ar[j] * __ldg(b + j);
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) x[row] = s;
}

void dense_apply(const DenseInv& D, const double* b, double* x, const int* done)
{
    if (D.n == 0) return;
    ProfScope prof(100, D.n, (long long)D.n * D.n, 8.0 * D.n * D.n + 16.0 * D.n);
    FC_LAUNCH(k_dense_gemv, (D.n + 7) / 8, 256, 0, D.n, D.ainv, b, x, done);
}

void dense_free(DenseInv& D)
{
    dfree(D.ainv);
    D = DenseInv();
}

} // namespace fc
