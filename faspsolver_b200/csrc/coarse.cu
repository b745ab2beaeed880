// coarse.cu — iterative coarsest-level solve on the device: conjugate gradients to a relative
// residual of `tol` inside ONE persistent cooperative kernel.
//
// Replaces fasp_coarse_itsolver (PreMGUtil.inl:37-58) -> fasp_solver_dcsr_spcg (KrySPcg.c:60-370) for
// hierarchies whose coarsest level is too large for the dense inverse (option coarse_dense_max) or
// whose coarsest operator is singular / semi-definite (pure Neumann, periodic problems: the dense
// factorisation reports a vanishing pivot, CG still solves the consistent system). Ported literally,
// the reference's loop is ~10 tiny kernels and 7 host-visible scalars per CG iteration, 20-25
// iterations per cycle; here the whole loop lives in one kernel: grid-wide
// steps are separated by cooperative-groups grid barriers, every CTA adds the per-CTA partials of a
// dot product in CTA order (deterministic, identical on all CTAs, so all of them take the same
// branches) and no scalar ever leaves the device. The launch is captured in the V-cycle's CUDA graph
// like any other kernel.
//
// Kept from the reference loop: x0 = 0 (PreMGCycle.c:151), relative residual against ||b||
// (StopType 1), tolerance tol = param->tol * 1e-4 (PreMGCycle.c:56), maxit = max(250, min(n^2, 1000))
// (PreMGUtil.inl:44), the division guard (KrySPcg.c:158-165) and the false-convergence re-check with
// the true residual, restarting with p = 0 at most MAX_RESTART times (KrySPcg.c:290-340). Not kept:
// the stagnation restart and the best-iterate bookkeeping (never triggered by CG on the SPD / M-matrix
// coarse operators an AMG setup produces); the GMRES safety net is not on the device.
#include "amg.cuh"
#include <cooperative_groups.h>

namespace fc {

namespace cg = cooperative_groups;
constexpr int CC_T = 256;

// sum of the per-CTA partials in CTA order; every thread of every CTA gets the same value
__device__ __forceinline__ double cc_total(const double* partials, int G, double* s_bcast)
{
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = 0.0;
        // fixed order: lane l adds partials l, l+32, ... ; lanes combined by the xor tree
        for (int b = threadIdx.x; b < G; b += 32) v += __ldcg(partials + b);
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (threadIdx.x == 0) *s_bcast = v;
    }
    __syncthreads();
    return *s_bcast;
}

__device__ __forceinline__ void cc_publish(double v, double* partials, double* s_w)
{
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = s_w[0];
        for (int w = 1; w < CC_T / 32; ++w) t += s_w[w];
        partials[blockIdx.x] = t;
    }
}

// y_i = sum_k a_ik x_k for the rows of this CTA's warps (one warp per row); returns this thread's share of
// sum_i y_i * d_i when d != nullptr (lane 0 of the row's warp carries it)
__device__ __forceinline__ double cc_spmv(int n, const int* __restrict__ ia, const int* __restrict__ ja,
                                          const double* __restrict__ val, const double* x, double* y,
                                          const double* b, const double* d, bool residual)
{
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (CC_T / 32) + (threadIdx.x >> 5), nw = gridDim.x * (CC_T / 32);
    double    part = 0.0;
    for (int i = gw; i < n; i += nw) {
        double s = 0.0;
        for (int k = ia[i] + lane; k < ia[i + 1]; k += 32) s += (val ? val[k] : 1.0) * __ldcg(x + ja[k]);
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) {
            const double out = residual ? b[i] - s : s;
            y[i]             = out;
            part += d ? out * __ldcg(d + i) : out * out;
        }
    }
    return part;
}

__global__ void __launch_bounds__(CC_T)
k_coarse_cg(const int n, const int* __restrict__ ia, const int* __restrict__ ja, const double* __restrict__ val,
            const double* __restrict__ b, double* x, double* p, double* r, double* t, double* partials,
            const int maxit, const double tol, int* iters_out, const int* done)
{
    if (done != nullptr && *done != 0) return;   // uniform over the grid: nobody reaches a barrier
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_w[CC_T / 32];
    __shared__ double s_b;
    const int G = gridDim.x;
    double*   P0 = partials, *P1 = partials + G, *P2 = partials + 2 * G;
    const int gt = blockIdx.x * CC_T + threadIdx.x, nt = G * CC_T;

    // x = 0 ; r = p = b ; rr = ||b||^2
    double v = 0.0;
    for (int i = gt; i < n; i += nt) {
        const double bi = b[i];
        x[i] = 0.0, r[i] = bi, p[i] = bi;
        v += bi * bi;
    }
    cc_publish(v, P0, s_w);
    grid.sync();
    double       rr     = cc_total(P0, G, &s_b);
    const double absr0  = sqrt(rr);
    const double normr0 = fmax(SMALLREAL, absr0);
    int          it = 0, more_step = 1;
    if (absr0 / normr0 < tol) {   // b == 0 (to SMALLREAL): x = 0 is the answer
        if (gt == 0 && iters_out) *iters_out = 0;
        return;
    }
    while (it < maxit) {
        ++it;
        // t = A p ; tp = (t, p)
        v = cc_spmv(n, ia, ja, val, p, t, nullptr, p, false);
        cc_publish(v, P1, s_w);
        grid.sync();
        const double tp = cc_total(P1, G, &s_b);
        if (!(fabs(tp) > SMALLREAL2)) break;   // possible breakdown (KrySPcg.c:158)
        const double alpha = rr / tp;
        // x += alpha p ; r -= alpha t ; rr_new = ||r||^2
        v = 0.0;
        for (int i = gt; i < n; i += nt) {
            x[i] += alpha * p[i];
            const double ri = __ldcg(r + i) - alpha * __ldcg(t + i);   // written by other CTAs: read through L2
            r[i] = ri;
            v += ri * ri;
        }
        cc_publish(v, P2, s_w);
        grid.sync();
        double rr_new = cc_total(P2, G, &s_b);
        bool   zero_p = false;
        if (sqrt(rr_new) / normr0 < tol) {
            // false-convergence guard: true residual r = b - A x (KrySPcg.c:290-340)
            v = cc_spmv(n, ia, ja, val, x, r, b, nullptr, true);
            cc_publish(v, P0, s_w);
            grid.sync();
            rr_new = cc_total(P0, G, &s_b);
            if (sqrt(rr_new) / normr0 < tol) break;
            if (more_step >= MAX_RESTART) break;
            ++more_step;
            zero_p = true;
        }
        // p = r + beta p
        const double beta = rr_new / rr;
        rr                = rr_new;
        for (int i = gt; i < n; i += nt) p[i] = __ldcg(r + i) + (zero_p ? 0.0 : beta * p[i]);
        grid.sync();
    }
    if (gt == 0 && iters_out) *iters_out = it;
}

void coarse_cg_setup(CoarseCG& C, const DevCSR& A, double tol)
{
    coarse_cg_free(C);
    Ctx& c = ctx();
    C.n    = A.rows;
    if (A.cols != A.rows) fail(ERROR_MAT_SIZE, "coarsest matrix is not square");
    int per_sm = 0;
    FC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_coarse_cg, CC_T, 0));
    if (per_sm < 1) fail(ERROR_SOLVER_MISC, "coarse CG: the cooperative kernel does not fit on an SM");
    if (per_sm > 4) per_sm = 4;
    long long grid = (long long)c.sm_count * per_sm;
    const long long want = ((long long)C.n + 7) / 8;   // a warp per row
    if (grid > want) grid = want < 1 ? 1 : want;
    C.grid  = (int)grid;
    C.work  = dalloc<double>(3 * (size_t)C.n + 3 * (size_t)C.grid + 8);
    C.iters = dalloc<int>(1);
    FC_CUDA(cudaMemsetAsync(C.iters, 0, sizeof(int), c.stream));
    const long long n2 = (long long)C.n * C.n;
    C.maxit = (int)std::max<long long>(250, std::min<long long>(n2, 1000));   // PreMGUtil.inl:44
    C.tol   = tol;
}

void coarse_cg_free(CoarseCG& C)
{
    dfree(C.work);
    dfree(C.iters);
    C = CoarseCG();
}

void coarse_cg_apply(const CoarseCG& C, const DevCSR& A, const double* b, double* x, const int* done)
{
    if (C.n == 0) return;
    Ctx&      c = ctx();
    ProfScope prof(101, C.n, A.nnz, 0.0);
    int       n = C.n, maxit = C.maxit;
    double    tol = C.tol;
    const int *ia = A.ia, *ja = A.ja;
    const double* val = A.val;
    double *p = C.work, *r = p + C.n, *t = r + C.n, *partials = t + C.n;
    int*    iters = C.iters;
    void*   args[] = {&n, &ia, &ja, &val, &b, &x, &p, &r, &t, &partials, &maxit, &tol, &iters, &done};
    FC_CUDA(cudaLaunchCooperativeKernel((void*)k_coarse_cg, dim3(C.grid), dim3(CC_T), args, 0, c.stream));
    if (c.capturing) c.captured++; else c.launches++;
}

int coarse_cg_last_iters(const CoarseCG& C)
{
    int h = 0;
    if (!C.iters) return 0;
    FC_CUDA(cudaMemcpyAsync(&h, C.iters, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
    FC_CUDA(cudaStreamSynchronize(ctx().stream));
    return h;
}

} // namespace fc
