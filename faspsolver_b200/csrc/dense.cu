// dense.cu — coarsest-level direct solve: a dense inverse built once on the device by blocked
// Gauss-Jordan elimination with partial pivoting (one cooperative kernel), applied per cycle as one GEMV.
//
// Replaces fasp_coarse_itsolver (PreMGUtil.inl:37-58: safeguarded CG to 1e-10, ~20-25
// iterations of ~10 tiny dependent kernels each) on the path; the coarsest matrices FASP
// produces are small and 25-30 % dense (SURVEY.md finding 8). The result differs from the
// reference's iterative coarse solve by its tolerance (<= 1e-10 relative).
#include "amg.cuh"
#include "reduce.cuh"
#include <cooperative_groups.h>

namespace fc {

__global__ void k_csr_to_dense(int n, const int* ia, const int* ja, const double* val, double* a)
{
    const int i = blockIdx.x;
    for (int k = ia[i] + threadIdx.x; k < ia[i + 1]; k += blockDim.x)
        atomicAdd(&a[(size_t)i * n + ja[k]], val ? val[k] : 1.0);   // duplicates add up
}

// ------------------------------------------------------------------------------------
// Blocked in-place Gauss-Jordan inversion with partial pivoting, ONE cooperative kernel.
//
// Panels of GJ_W columns. For a panel K = [k0, k1):
//   panel phase   the GJ column steps k = k0..k1-1 restricted to the panel's columns (all n rows): the
//                 multipliers of a GJ step come from column k only, so the panel can run ahead of the
//                 other columns. Per column: grid-wide pivot search, exchange of panel rows k <-> p,
//                 rank-1 update of the panel. Afterwards the panel holds G[:, K] of the panel
//                 transformation G (rows K: the inverted pivot block, other rows: -C_R C_K^-1).
//   swap phase    the panel's row exchanges applied to every other column; T = A[K, other] saved,
//                 A[K, other] = 0
//   update phase  A[:, other] += G[:, K] * T : a tiled FP64 product, n x (n - w) x w, the only phase
//                 that touches the whole matrix (n / GJ_W passes instead of n)
// Finally the row exchanges are undone as column exchanges in reverse order. The unblocked version
// needed 4 launches and a pass over the whole matrix per column (20 k launches, 25 M-entry passes for the
// 5041-row coarsest level of the 256^3 hierarchy); this one is a single launch and n / 32 passes.
// Grid-wide steps are separated by cooperative-groups grid barriers (the kernel is launched with
// cudaLaunchCooperativeKernel, all CTAs co-resident).
// ------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int GJ_W  = 32;    // panel width = lanes of a warp: lane j owns panel column j
constexpr int GJ_T  = 256;   // threads per CTA
constexpr int GJ_TM = 64;    // update phase: output tile of a CTA (GJ_TM x GJ_TM)

struct GjStatus {
    double min_pivot;   // smallest |pivot| met
    double max_entry;   // largest |a_ij| of the input
    int    singular;    // a zero / non-finite pivot was met
};

__global__ void __launch_bounds__(GJ_T)
k_gj_blocked(const int n, double* __restrict__ a, int* __restrict__ piv, double* __restrict__ tbuf,
             double* __restrict__ cand_val, int* __restrict__ cand_idx, GjStatus* status)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_rowk[GJ_W], s_rowp[GJ_W];
    __shared__ double s_cv[GJ_T / 32];
    __shared__ int    s_ci[GJ_T / 32];
    __shared__ int    s_p;
    __shared__ double s_A[GJ_TM][GJ_W + 1];   // update phase: G[rows, K] tile
    __shared__ double s_B[GJ_W][GJ_TM + 1];   // update phase: T[K, cols] tile
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = gridDim.x, nwarps = G * (GJ_T / 32), gw = blockIdx.x * (GJ_T / 32) + wid;
    const size_t ld = (size_t)n;

    // largest entry of the input (scale of the singularity test)
    {
        double m = 0.0;
        for (size_t i = (size_t)blockIdx.x * GJ_T + tid; i < ld * ld; i += (size_t)G * GJ_T) m = fmax(m, fabs(a[i]));
        for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
        if (lane == 0) s_cv[wid] = m;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < GJ_T / 32; ++w) m = fmax(m, s_cv[w]);
            cand_val[blockIdx.x] = m;
        }
        grid.sync();
        if (blockIdx.x == 0 && tid == 0) {
            double mm = 0.0;
            for (int b = 0; b < G; ++b) mm = fmax(mm, cand_val[b]);
            status->max_entry = mm;
            status->min_pivot = 1e300;
            status->singular  = 0;
        }
        grid.sync();
    }

    // this CTA's pivot candidate for column k among rows >= k: the rows warp gw + m * nwarps
    auto publish_candidate = [&](int k) {
        double best = -1.0;
        int    bi   = n;
        for (int i = gw; i < n; i += nwarps) {      // warp-uniform row; lane 0 reads the entry
            if (i < k) continue;
            const double v = fabs(__ldcg(a + (size_t)i * ld + k));
            if (v > best || (v == best && i < bi)) best = v, bi = i;
        }
        if (lane == 0) s_cv[wid] = best, s_ci[wid] = bi;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < GJ_T / 32; ++w)
                if (s_cv[w] > best || (s_cv[w] == best && s_ci[w] < bi)) best = s_cv[w], bi = s_ci[w];
            cand_val[blockIdx.x] = best;
            cand_idx[blockIdx.x] = bi;
        }
        __syncthreads();
    };

    for (int k0 = 0; k0 < n; k0 += GJ_W) {
        const int k1 = (k0 + GJ_W < n) ? k0 + GJ_W : n;
        const int w  = k1 - k0;
        // ---------------- panel phase ----------------
        publish_candidate(k0);
        for (int k = k0; k < k1; ++k) {
            grid.sync();   // (A) candidates of column k visible; the updates of column k-1 are complete
            // every CTA picks the same pivot row: largest |a_ik|, ties to the smaller row index
            if (wid == 0) {
                double best = -1.0;
                int    bi   = n;
                for (int b = lane; b < G; b += 32) {
                    const double v  = __ldcg(cand_val + b);
                    const int    ix = __ldcg(cand_idx + b);
                    if (v > best || (v == best && ix < bi)) best = v, bi = ix;
                }
                for (int off = 16; off > 0; off >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, off);
                    const int    oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
                }
                if (lane == 0) s_p = (bi < n) ? bi : k;
            }
            __syncthreads();
            const int p = s_p;
            // old panel rows k and p, read by everybody before their owners overwrite them
            if (tid < w) {
                s_rowk[tid] = __ldcg(a + (size_t)k * ld + k0 + tid);
                s_rowp[tid] = __ldcg(a + (size_t)p * ld + k0 + tid);
            }
            __syncthreads();
            grid.sync();   // (B)
            const int    kc   = k - k0;
            const double pv   = s_rowp[kc];            // the pivot a[p][k]
            const double pinv = 1.0 / pv;
            if (blockIdx.x == 0 && tid == 0) {
                piv[k] = p;
                const double apv = fabs(pv);
                if (!(apv > 0.0) || !isfinite(pinv)) status->singular = 1;
                if (apv < status->min_pivot) status->min_pivot = apv;
            }
            // new pivot row (lane j owns panel column j)
            const double rk = (lane < w) ? ((lane == kc) ? pinv : s_rowp[lane] * pinv) : 0.0;
            // rank-1 update of the panel, one row per warp; next column's candidates on the way
            double nbest = -1.0;
            int    nbi   = n;
            for (int i = gw; i < n; i += nwarps) {
                double* row = a + (size_t)i * ld + k0;
                double  out;
                if (i == k) {
                    out = rk;
                } else {
                    // row p receives the old row k (the exchange), every other row keeps its own
                    const double cur = (lane < w) ? ((i == p) ? s_rowk[lane] : __ldcg(row + lane)) : 0.0;
                    const double f   = __shfl_sync(0xffffffffu, cur, kc);
                    out              = ((lane == kc) ? 0.0 : cur) - f * rk;
                }
                if (lane < w) row[lane] = out;
                if (k + 1 < k1 && i > k) {   // candidate for column k+1 (lane kc+1 holds it)
                    const double v = fabs(__shfl_sync(0xffffffffu, out, kc + 1));
                    if (v > nbest || (v == nbest && i < nbi)) nbest = v, nbi = i;
                }
            }
            if (k + 1 < k1) {
                if (lane == 0) s_cv[wid] = nbest, s_ci[wid] = nbi;
                __syncthreads();
                if (tid == 0) {
                    for (int ww = 1; ww < GJ_T / 32; ++ww)
                        if (s_cv[ww] > nbest || (s_cv[ww] == nbest && s_ci[ww] < nbi)) nbest = s_cv[ww], nbi = s_ci[ww];
                    cand_val[blockIdx.x] = nbest;
                    cand_idx[blockIdx.x] = nbi;
                }
                __syncthreads();
            }
        }
        grid.sync();   // panel complete, piv[k0..k1) visible
        // ---------------- swap phase: one thread per column outside the panel ----------------
        for (int j = blockIdx.x * GJ_T + tid; j < n; j += G * GJ_T) {
            if (j >= k0 && j < k1) continue;
            for (int k = k0; k < k1; ++k) {
                const int p = __ldcg(piv + k);
                if (p != k) {
                    const double t        = __ldcg(a + (size_t)k * ld + j);
                    a[(size_t)k * ld + j] = __ldcg(a + (size_t)p * ld + j);
                    a[(size_t)p * ld + j] = t;
                }
            }
            for (int k = k0; k < k1; ++k) {
                tbuf[(size_t)(k - k0) * ld + j] = __ldcg(a + (size_t)k * ld + j);
                a[(size_t)k * ld + j]           = 0.0;
            }
        }
        grid.sync();
        // ---------------- update phase: A[:, other] += G[:, K] * T ----------------
        {
            const int tiles  = (n + GJ_TM - 1) / GJ_TM;
            const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads, 4 x 4 outputs each
            for (int t = blockIdx.x; t < tiles * tiles; t += G) {
                const int r0 = (t / tiles) * GJ_TM, c0 = (t % tiles) * GJ_TM;
                __syncthreads();
                for (int e = tid; e < GJ_TM * GJ_W; e += GJ_T) {
                    const int r = e / GJ_W, c = e % GJ_W;
                    s_A[r][c] = (r0 + r < n && c < w) ? __ldcg(a + (size_t)(r0 + r) * ld + k0 + c) : 0.0;
                }
                for (int e = tid; e < GJ_W * GJ_TM; e += GJ_T) {
                    const int r = e / GJ_TM, c = e % GJ_TM;
                    s_B[r][c] = (r < w && c0 + c < n) ? __ldcg(tbuf + (size_t)r * ld + c0 + c) : 0.0;
                }
                __syncthreads();
                double acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
                for (int kk = 0; kk < GJ_W; ++kk) {
                    double av[4], bv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) av[i] = s_A[ty + 16 * i][kk];
#pragma unroll
                    for (int j = 0; j < 4; ++j) bv[j] = s_B[kk][tx + 16 * j];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = r0 + ty + 16 * i;
                    if (r >= n) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = c0 + tx + 16 * j;
                        if (c >= n || (c >= k0 && c < k1)) continue;   // the panel's own columns hold G
                        a[(size_t)r * ld + c] = __ldcg(a + (size_t)r * ld + c) + acc[i][j];
                    }
                }
            }
        }
        grid.sync();
    }
    // undo the row exchanges as column exchanges, in reverse order: one thread per row
    for (int i = blockIdx.x * GJ_T + tid; i < n; i += G * GJ_T) {
        double* row = a + (size_t)i * ld;
        for (int k = n - 1; k >= 0; --k) {
            const int p = __ldcg(piv + k);
            if (p != k) {
                const double t = __ldcg(row + k);
                row[k]         = __ldcg(row + p);
                row[p]         = t;
            }
        }
    }
}

// Inverts the n x n row-major matrix `a` in place. Returns false (and leaves `a` undefined) when a pivot
// is zero or below 1e-14 times the largest entry: a singular / semi-definite coarsest operator (pure
// Neumann, periodic problems) must not turn into an inf/NaN preconditioner silently.
static bool gauss_jordan(int n, double* a)
{
    Ctx& c = ctx();
    if (n == 0) return true;
    int per_sm = 0;
    FC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gj_blocked, GJ_T, 0));
    if (per_sm < 1) fail(ERROR_SOLVER_MISC, "dense inverse: the cooperative kernel does not fit on an SM");
    if (per_sm > 2) per_sm = 2;
    int grid = c.sm_count * per_sm;
    const int want = (n + 7) / 8;   // a warp per row in the panel phase
    if (grid > want) grid = want < 1 ? 1 : want;
    int*      piv  = dalloc<int>(n);
    double*   tbuf = dalloc<double>((size_t)GJ_W * n);
    double*   cval = dalloc<double>(grid);
    int*      cidx = dalloc<int>(grid);
    GjStatus* st   = dalloc<GjStatus>(1);
    GjStatus  hs;
    try {
        int   nn     = n;
        void* args[] = {&nn, &a, &piv, &tbuf, &cval, &cidx, &st};
        FC_CUDA(cudaLaunchCooperativeKernel((void*)k_gj_blocked, dim3(grid), dim3(GJ_T), args, 0, c.stream));
        c.launches++;
        FC_CUDA(cudaMemcpyAsync(&hs, st, sizeof(hs), cudaMemcpyDeviceToHost, c.stream));
        FC_CUDA(cudaStreamSynchronize(c.stream));
    } catch (...) {
        dfree(piv), dfree(tbuf), dfree(cval), dfree(cidx), dfree(st);
        throw;
    }
    dfree(piv), dfree(tbuf), dfree(cval), dfree(cidx), dfree(st);
    return !(hs.singular || !(hs.min_pivot > 1e-14 * hs.max_entry));
}

bool dense_invert_csr(DenseInv& D, const DevCSR& A)
{
    dense_free(D);
    const int n = A.rows;
    if (A.cols != n) fail(ERROR_MAT_SIZE, "coarsest matrix is not square");
    D.n    = n;
    D.ainv = dalloc<double>((size_t)n * n);
    FC_CUDA(cudaMemsetAsync(D.ainv, 0, sizeof(double) * (size_t)n * n, ctx().stream));
    if (n > 0) FC_LAUNCH(k_csr_to_dense, n, 128, 0, n, A.ia, A.ja, A.val, D.ainv);
    if (gauss_jordan(n, D.ainv)) return true;
    dense_free(D);
    return false;
}

bool dense_invert_bsr(DenseInv& D, const DevBSR& A)
{
    dense_free(D);
    if (A.ROW != A.COL) fail(ERROR_MAT_SIZE, "coarsest BSR matrix is not square");
    const int n = A.ROW * A.nb;
    D.n         = n;
    D.ainv      = dalloc<double>((size_t)n * n);
    FC_CUDA(cudaMemsetAsync(D.ainv, 0, sizeof(double) * (size_t)n * n, ctx().stream));
    bsr_to_dense(A, D.ainv);
    if (gauss_jordan(n, D.ainv)) return true;
    dense_free(D);
    return false;
}

bool dense_invert_host(DenseInv& D, int n, const std::vector<double>& a)
{
    dense_free(D);
    D.n    = n;
    D.ainv = dalloc<double>((size_t)n * n);
    FC_CUDA(cudaMemcpyAsync(D.ainv, a.data(), sizeof(double) * (size_t)n * n,
                            cudaMemcpyHostToDevice, ctx().stream));
    FC_CUDA(cudaStreamSynchronize(ctx().stream));
    if (gauss_jordan(n, D.ainv)) return true;
    dense_free(D);
    return false;
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
