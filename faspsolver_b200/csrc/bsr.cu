// bsr.cu — block-CSR kernels for sm_100a: SpMV / aAxpy / residual / block-Jacobi sweep.
//
// Replaces fasp_blas_dbsr_mxv (BlaSpmvBSR.c:1055), fasp_blas_dbsr_aAxpy (:514) and
// fasp_smoother_dbsr_jacobi1 (ItrSmootherBSR.c:263) on the BSR solve path.
//
// Same structure as the pipelined CSR kernel (spmv.cu): persistent CTAs, the val / ja / ia
// slices of the next row block are staged into shared memory by TMA bulk copies while the
// current one is consumed. A CTA owns RB = 32 block rows and nb threads per block row: thread
// (I, i) produces scalar row i of block row I. It walks the blocks of its row in storage order;
// for each block it reads its nb entries from the staged slice, gathers x_J (the nb threads of
// a block row read the same addresses: one transaction) and evaluates
//     y_i += (((A_i0 x_0 + A_i1 x_1) + A_i2 x_2) + ...)
// with separate multiply/add roundings, i.e. the expression of fasp_blas_smat_ypAx_nc*
// (BlaSmallMat.c:673-681): the result is bit-identical to the sequential CPU code.
#include "bsr.cuh"
#include "reduce.cuh"
#include "pipe.cuh"

namespace fc {

constexpr int B_MAX_STAGES = 4;

struct BsrView {
    const int*    ia;
    const int*    ja;
    const double* val;
    const int2*   blkdesc;
    int           cap;   // blocks per stage
    int           ROW;
};

struct BMeta {
    int r0, nrows, k0, n;
};

// B_RB block rows per CTA (NB threads each), U blocks of a row in flight per thread
template <int MODE, int NB, int B_RB, int U>
__global__ void __launch_bounds__(B_RB* NB)
bsr_pipe_kernel(const BsrView A, const BsrArgs a, const int nblk, const int nstages,
                double* partials, unsigned int* ticket)
{
    constexpr int T   = B_RB * NB;
    constexpr int NB2 = NB * NB;
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) unsigned long long s_bar[B_MAX_STAGES];
    __shared__ BMeta  s_meta[B_MAX_STAGES];
    __shared__ double s_tmp[T];

    if (a.done != nullptr && *a.done != 0) return;

    const int    tid         = threadIdx.x;
    const int    cap         = A.cap;
    const int    G           = gridDim.x;
    const size_t val_doubles = (size_t)cap * NB2 + 4;
    const size_t stage_bytes = (val_doubles * 8 + (size_t)(cap + 8) * 4 + (size_t)(B_RB + 8) * 4 + 127) & ~(size_t)127;
    auto st_val = [&](int s) { return reinterpret_cast<double*>(s_raw + (size_t)s * stage_bytes); };
    auto st_ja  = [&](int s) { return reinterpret_cast<int*>(st_val(s) + val_doubles); };
    auto st_ia  = [&](int s) { return st_ja(s) + cap + 8; };
    const double* __restrict__ x = a.x;
    const int zero = (int)gridDim.y - 1;   // 0 at run time, opaque to the compiler (gather rounds)

    auto issue = [&](int blk, int s) {
        const int2 d0 = A.blkdesc[blk], d1 = A.blkdesc[blk + 1];
        BMeta      m{d0.x, d1.x - d0.x, d0.y, d1.y - d0.y};
        s_meta[s] = m;
        if (m.n <= cap) {
            const int          k0a = m.k0 & ~3;
            const unsigned int na  = (unsigned int)(((m.k0 + m.n + 3) & ~3) - k0a);
            const long long    v0  = ((long long)m.k0 * NB2) & ~1LL;
            const unsigned int nv  = (unsigned int)((((long long)(m.k0 + m.n) * NB2 + 1) & ~1LL) - v0);
            const int          r0a = m.r0 & ~3;
            const unsigned int nra = (unsigned int)(((m.r0 + m.nrows + 1 + 3) & ~3) - r0a);
            mbar_expect_tx(&s_bar[s], na * 4u + nv * 8u + nra * 4u);
            if (na) bulk_g2s(st_ja(s), A.ja + k0a, na * 4u, &s_bar[s]);
            if (nv) bulk_g2s(st_val(s), A.val + v0, nv * 8u, &s_bar[s]);
            bulk_g2s(st_ia(s), A.ia + r0a, nra * 4u, &s_bar[s]);
        } else {
            mbar_expect_tx(&s_bar[s], 0u);
        }
    };

    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < nstages; ++s) {
            const long long blk = (long long)blockIdx.x + (long long)s * G;
            if (blk < nblk) issue((int)blk, s);
        }
    }
    __syncthreads();

    double     red_dot = 0.0, red_n2 = 0.0;
    const bool want_dot = a.red.dot_out != nullptr;
    const bool want_n2  = a.red.nrm2_out != nullptr;
    const int  il = tid / NB;         // local block row
    const int  i  = tid - il * NB;    // scalar row inside the block

    int s = 0, ph = 0;
    for (long long blk = blockIdx.x; blk < nblk; blk += G) {
        mbar_wait(&s_bar[s], (unsigned int)ph);
        const BMeta m      = s_meta[s];
        const bool  staged = m.n <= cap;
        const bool  valid  = il < m.nrows;
        const int   I      = m.r0 + il;
        double      acc    = 0.0;
        double      drow[MODE == BSR_JACOBI ? NB : 1];
        if (valid) {
            // where this block row's blocks live: staged slice or (oversized row) global memory
            const double* vbase;
            const int*    jbase;
            int           ka, kb;
            if (staged) {
                const int* sia = st_ia(s) + (m.r0 & 3);
                ka    = sia[il] - m.k0;
                kb    = sia[il + 1] - m.k0;
                vbase = st_val(s) + (((long long)m.k0 * NB2) & 1);
                jbase = st_ja(s) + (m.k0 & 3);
            } else {
                ka    = A.ia[I];
                kb    = A.ia[I + 1];
                vbase = A.val;
                jbase = A.ja;
            }
            const size_t row = (size_t)I * NB + i;
            if (MODE == BSR_JACOBI) {   // this thread's row of Dinv_I, fetched now: its latency hides behind the gathers
                const double* D = a.diaginv + (size_t)I * NB2 + i * NB;
#pragma unroll
                for (int j = 0; j < NB; ++j) drow[j] = __ldg(D + j);
            }
            if (MODE == BSR_MXV) acc = 0.0;
            else if (MODE == BSR_AXPY) {
                const double y0 = a.y[row];
                acc = (a.alpha == 1.0) ? y0 : __dmul_rn(1.0 / a.alpha, y0);   // fasp_blas_darray_ax
            } else acc = a.b[row];

            for (int kk = ka; kk < kb; kk += U) {
                int    col[U];
                double xv[U][NB];
#pragma unroll
                for (int u = 0; u < U; ++u) col[u] = (kk + u < kb) ? jbase[kk + u] : -1;
                // Two phases (all column indices out of shared memory, then all gathers), enforced by a term that is
                // zero at run time but depends on every index of the round: ptxas otherwise interleaves the LDS and
                // LDG and can serialise the gathers on a shared scoreboard (spmv.cu, profiles/r02_l1_sweep_scheduling.txt).
                // 272^3 3x3 blocks, 64 block rows x 8 blocks in flight: Jacobi sweep 3.06 -> 2.25 ms, y = A x 2.79 -> 2.62,
                // r = b - A x 2.33 -> 2.32 (profiles/r02_bsr272_kernel_shapes.txt)
                int any = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) any |= col[u];
                const int dep = any & zero;
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < NB; ++j)
                        xv[u][j] = (col[u] >= 0) ? __ldg(x + (size_t)(col[u] + dep) * NB + j) : 0.0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (col[u] < 0) continue;
                    if (MODE == BSR_JACOBI && col[u] == I) continue;   // j != i (ItrSmootherBSR.c:373)
                    const double* Ar = vbase + (size_t)(kk + u) * NB2 + i * NB;
                    if (NB <= 7) {
                        double e = __dmul_rn(Ar[0], xv[u][0]);
#pragma unroll
                        for (int j = 1; j < NB; ++j) e = __dadd_rn(e, __dmul_rn(Ar[j], xv[u][j]));
                        acc = (MODE == BSR_MXV || MODE == BSR_AXPY) ? __dadd_rn(acc, e) : __dsub_rn(acc, e);
                    } else {   // generic loop of fasp_blas_smat_ypAx / ymAx: term by term
#pragma unroll
                        for (int j = 0; j < NB; ++j) {
                            const double p = __dmul_rn(Ar[j], xv[u][j]);
                            acc = (MODE == BSR_MXV || MODE == BSR_AXPY) ? __dadd_rn(acc, p) : __dsub_rn(acc, p);
                        }
                    }
                }
            }
        }
        double out = acc;
        if (MODE == BSR_AXPY) {
            if (a.alpha != 1.0) out = __dmul_rn(a.alpha, acc);
        }
        if (MODE == BSR_JACOBI) {
            // u_I = Dinv_I * b_tmp_I  (fasp_blas_smat_mxv, BlaSmallMat.c:238)
            s_tmp[tid] = acc;
            __syncthreads();
            if (valid) {
                const double* t = s_tmp + il * NB;
                if (NB <= 7) {
                    double e = __dmul_rn(drow[0], t[0]);
#pragma unroll
                    for (int j = 1; j < NB; ++j) e = __dadd_rn(e, __dmul_rn(drow[j], t[j]));
                    out = e;
                } else {
                    double e = 0.0;
#pragma unroll
                    for (int j = 0; j < NB; ++j) e = __dadd_rn(e, __dmul_rn(drow[j], t[j]));
                    out = e;
                }
            }
        }
        if (valid) {
            const size_t row = (size_t)I * NB + i;
            a.y[row]         = out;
            if (want_dot) red_dot += out * a.red.dot_with[row];
            if (want_n2) red_n2 += out * out;
        }
        __syncthreads();
        if (tid == 0) {
            const long long refill = blk + (long long)nstages * G;
            if (refill < nblk) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue((int)refill, s);
            }
        }
        if (++s == nstages) s = 0, ph ^= 1;
    }

    if (want_dot || want_n2) {
        double v[2] = {red_dot, red_n2};
        grid_reduce<2, 0>(v, partials, ticket, [&](const double* t) {
            if (want_dot) *a.red.dot_out = t[0];
            if (want_n2) *a.red.nrm2_out = t[1];
        });
    }
}

// ------------------------------------------------------------------------------------
// identity-block operators (UA-AMG P / R): y_I (+)= sum_{k in row I} x_{ja[k]} per scalar component.
// The CPU code multiplies by the stored identity blocks (fasp_blas_dbsr_mxv/aAxpy, 76 B per block);
// 1*x_i + 0*x_j + 0*x_k == x_i, so summing the gathered entries in k order is bit-identical.
// ------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256)
bsr_ident_kernel(const int ROW, const int nb, const int* __restrict__ ia, const int* __restrict__ ja,
                 const BsrArgs a, double* partials, unsigned int* ticket)
{
    if (a.done != nullptr && *a.done != 0) return;
    const long long n  = (long long)ROW * nb;
    double     red_dot = 0.0, red_n2 = 0.0;
    const bool want_dot = a.red.dot_out != nullptr, want_n2 = a.red.nrm2_out != nullptr;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long long)gridDim.x * 256) {
        const int I = (int)(t / nb), i = (int)(t - (long long)I * nb);
        double acc;
        if (MODE == BSR_MXV) acc = 0.0;
        else if (MODE == BSR_AXPY) acc = (a.alpha == 1.0) ? a.y[t] : __dmul_rn(1.0 / a.alpha, a.y[t]);
        else acc = a.b[t];
        for (int k = ia[I]; k < ia[I + 1]; ++k) {
            const double xv = __ldg(a.x + (size_t)ja[k] * nb + i);
            acc = (MODE == BSR_RESID) ? __dsub_rn(acc, xv) : __dadd_rn(acc, xv);
        }
        if (MODE == BSR_AXPY && a.alpha != 1.0) acc = __dmul_rn(a.alpha, acc);
        a.y[t] = acc;
        if (want_dot) red_dot += acc * a.red.dot_with[t];
        if (want_n2) red_n2 += acc * acc;
    }
    if (want_dot || want_n2) {
        double v[2] = {red_dot, red_n2};
        grid_reduce<2, 0>(v, partials, ticket, [&](const double* t) {
            if (want_dot) *a.red.dot_out = t[0];
            if (want_n2) *a.red.nrm2_out = t[1];
        });
    }
}

// y_I = Dinv_I b_I (fasp_blas_smat_mxv, BlaSmallMat.c:238): thread per scalar row
__global__ void __launch_bounds__(256)
bsr_dinv_kernel(const int ROW, const int nb, const BsrArgs a, double* partials, unsigned int* ticket)
{
    if (a.done != nullptr && *a.done != 0) return;
    const long long n  = (long long)ROW * nb;
    double     red_dot = 0.0, red_n2 = 0.0;
    const bool want_dot = a.red.dot_out != nullptr, want_n2 = a.red.nrm2_out != nullptr;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long long)gridDim.x * 256) {
        const int     I = (int)(t / nb), i = (int)(t - (long long)I * nb);
        const double* D = a.diaginv + ((size_t)I * nb + i) * nb;
        const double* bb = a.b + (size_t)I * nb;
        double        e;
        if (nb <= 7) {   // the unrolled nc2/nc3/nc5/nc7 (and 4, 6) expressions start from the first product
            e = __dmul_rn(D[0], bb[0]);
            for (int j = 1; j < nb; ++j) e = __dadd_rn(e, __dmul_rn(D[j], bb[j]));
        } else {
            e = 0.0;
            for (int j = 0; j < nb; ++j) e = __dadd_rn(e, __dmul_rn(D[j], bb[j]));
        }
        a.y[t] = e;
        if (want_dot) red_dot += e * a.red.dot_with[t];
        if (want_n2) red_n2 += e * e;
    }
    if (want_dot || want_n2) {
        double v[2] = {red_dot, red_n2};
        grid_reduce<2, 0>(v, partials, ticket, [&](const double* t) {
            if (want_dot) *a.red.dot_out = t[0];
            if (want_n2) *a.red.nrm2_out = t[1];
        });
    }
}

static int flat_grid(long long n)
{
    long long g = (n + 1023) / 1024, cap = (long long)ctx().sm_count * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

template <int MODE>
static void launch_ident(const DevBSR& A, const BsrArgs& a)
{
    const int g = flat_grid((long long)A.ROW * A.nb);
    FC_LAUNCH((bsr_ident_kernel<MODE>), g, 256, 0, A.ROW, A.nb, A.ia, A.ja, a, red_partials(g), red_ticket());
}

template <int MODE, int NB, int B_RB, int U>
static void launch_var(const DevBSR& A, const BsrView& v, const BsrArgs& a)
{
    Ctx&         c    = ctx();
    const size_t vald = (size_t)A.blk_cap * NB * NB + 4;
    const size_t stage = (vald * 8 + (size_t)(A.blk_cap + 8) * 4 + (size_t)(B_RB + 8) * 4 + 127) & ~(size_t)127;
    int          nst   = c.opt.bsr_stages;
    if (nst < 2) nst = 2;
    if (nst > B_MAX_STAGES) nst = B_MAX_STAGES;
    const size_t smem  = stage * nst;
    static bool  attr_set = false;
    if (!attr_set) {
        FC_CUDA(cudaFuncSetAttribute(bsr_pipe_kernel<MODE, NB, B_RB, U>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     200 * 1024));
        attr_set = true;
    }
    if (smem > 200 * 1024) fail(ERROR_INPUT_PAR, "BSR stage ring of %zu bytes does not fit in shared memory", smem);
    int per_sm = (int)((size_t)(220 * 1024) / (smem + 4096));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2048 / (B_RB * NB)) per_sm = 2048 / (B_RB * NB);
    if (per_sm > 12) per_sm = 12;
    long long grid = (long long)c.sm_count * per_sm;
    if (grid > A.nblk) grid = A.nblk;
    double*       part = nullptr;
    unsigned int* tick = nullptr;
    if (a.red.dot_out || a.red.nrm2_out) {
        part = red_partials((size_t)grid);
        tick = red_ticket();
    }
    FC_LAUNCH((bsr_pipe_kernel<MODE, NB, B_RB, U>), (int)grid, B_RB * NB, smem, v, a, A.nblk, nst, part, tick);
}

template <int MODE, int NB>
static void launch_nb(const DevBSR& A, const BsrView& v, const BsrArgs& a)
{
    if constexpr (NB <= 4) {   // the small blocks of black-oil / elasticity systems: tunable shape
        const bool u8 = ctx().opt.bsr_u >= 8;
        if (A.blk_rb == 64) {
            if (u8) launch_var<MODE, NB, 64, 8>(A, v, a);
            else launch_var<MODE, NB, 64, 4>(A, v, a);
        } else {
            if (u8) launch_var<MODE, NB, 32, 8>(A, v, a);
            else launch_var<MODE, NB, 32, 4>(A, v, a);
        }
    } else {
        launch_var<MODE, NB, 32, 2>(A, v, a);
    }
}

template <int MODE>
static void launch_mode(const DevBSR& A, const BsrView& v, const BsrArgs& a)
{
    switch (A.nb) {
        case 1: launch_nb<MODE, 1>(A, v, a); break;
        case 2: launch_nb<MODE, 2>(A, v, a); break;
        case 3: launch_nb<MODE, 3>(A, v, a); break;
        case 4: launch_nb<MODE, 4>(A, v, a); break;
        case 5: launch_nb<MODE, 5>(A, v, a); break;
        case 6: launch_nb<MODE, 6>(A, v, a); break;
        case 7: launch_nb<MODE, 7>(A, v, a); break;
        case 8: launch_nb<MODE, 8>(A, v, a); break;
        default: fail(ERROR_INPUT_PAR, "BSR block size nb = %d is not supported on the device (1..8)", A.nb);
    }
}

void bsr_launch(const DevBSR& A, const BsrArgs& a)
{
    if (A.ROW == 0) return;
    double pbytes = bsr_spmv_bytes(A, a.mode != BSR_MXV);
    if (a.mode == BSR_JACOBI) pbytes += 8.0 * A.nb * A.nb * A.ROW;
    if (a.mode == BSR_DINV) pbytes = 8.0 * A.nb * A.ROW * (2.0 + A.nb);
    ProfScope prof(200 + a.mode + (a.conditional ? 50 : 0), A.ROW, A.NNZ, pbytes);
    if (a.mode == BSR_DINV) {
        if (!a.diaginv) fail(ERROR_DATA_STRUCTURE, "block Jacobi sweep without diaginv");
        const int g = flat_grid((long long)A.ROW * A.nb);
        FC_LAUNCH(bsr_dinv_kernel, g, 256, 0, A.ROW, A.nb, a, red_partials(g), red_ticket());
        return;
    }
    if (A.ident) {
        switch (a.mode) {
            case BSR_MXV: launch_ident<BSR_MXV>(A, a); return;
            case BSR_AXPY:
                if (a.alpha == 0.0) return;
                launch_ident<BSR_AXPY>(A, a);
                return;
            case BSR_RESID: launch_ident<BSR_RESID>(A, a); return;
            default: fail(ERROR_INPUT_PAR, "identity-block operator: mode %d is not supported", a.mode);
        }
    }
    BsrView   v{A.ia, A.ja, A.val, A.blkdesc, A.blk_cap, A.ROW};
    switch (a.mode) {
        case BSR_MXV: launch_mode<BSR_MXV>(A, v, a); break;
        case BSR_AXPY:
            if (a.alpha == 0.0) return;   // BlaSpmvBSR.c:541-543: nothing to compute
            if (a.alpha == -1.0) {        // -((-y) + sum) == y - sum, term by term
                BsrArgs r = a;
                r.mode    = BSR_RESID;
                r.b       = a.y;
                launch_mode<BSR_RESID>(A, v, r);
            } else
                launch_mode<BSR_AXPY>(A, v, a);
            break;
        case BSR_RESID: launch_mode<BSR_RESID>(A, v, a); break;
        case BSR_JACOBI:
            if (!a.diaginv) fail(ERROR_DATA_STRUCTURE, "block Jacobi sweep without diaginv");
            launch_mode<BSR_JACOBI>(A, v, a);
            break;
        default: fail(ERROR_INPUT_PAR, "bsr_launch: unknown mode %d", a.mode);
    }
}

// ------------------------------------------------------------------------------------
void bsr_upload(DevBSR& d, int ROW, int COL, long long NNZ, int nb, const int* ia, const int* ja,
                const double* val, bool detect_identity)
{
    ensure_init();
    Ctx& c = ctx();
    bsr_free(d);
    if (nb < 1 || nb > 8) fail(ERROR_INPUT_PAR, "BSR block size nb = %d is not supported on the device (1..8)", nb);
    if (NNZ >= 2147483647LL / (nb * nb)) fail(ERROR_MAT_SIZE, "bsr_upload: matrix exceeds 32-bit offsets");
    d.ROW = ROW, d.COL = COL, d.NNZ = NNZ, d.nb = nb;
    const size_t pad = 8, nb2 = (size_t)nb * nb;
    if (detect_identity && NNZ > 0) {
        bool ident = true;
        for (long long k = 0; k < NNZ && ident; ++k)
            for (int e = 0; e < nb * nb; ++e)
                if (val[(size_t)k * nb2 + e] != ((e / nb == e % nb) ? 1.0 : 0.0)) {
                    ident = false;
                    break;
                }
        d.ident = ident;
    }
    d.ia  = dalloc<int>((size_t)ROW + 1 + pad);
    d.ja  = dalloc<int>((size_t)NNZ + pad);
    if (d.ident) {
        FC_CUDA(cudaMemcpyAsync(d.ia, ia, sizeof(int) * ((size_t)ROW + 1), cudaMemcpyHostToDevice, c.stream));
        FC_CUDA(cudaMemsetAsync(d.ia + ROW + 1, 0, sizeof(int) * pad, c.stream));
        FC_CUDA(cudaMemcpyAsync(d.ja, ja, sizeof(int) * (size_t)NNZ, cudaMemcpyHostToDevice, c.stream));
        FC_CUDA(cudaMemsetAsync(d.ja + NNZ, 0, sizeof(int) * pad, c.stream));
        d.bytes = sizeof(int) * ((size_t)ROW + 1 + NNZ + 2 * pad);
        FC_CUDA(cudaStreamSynchronize(c.stream));
        red_partials((size_t)c.sm_count * 16);
        return;
    }
    d.val = dalloc<double>((size_t)NNZ * nb2 + pad);
    FC_CUDA(cudaMemcpyAsync(d.ia, ia, sizeof(int) * ((size_t)ROW + 1), cudaMemcpyHostToDevice, c.stream));
    FC_CUDA(cudaMemsetAsync(d.ia + ROW + 1, 0, sizeof(int) * pad, c.stream));
    FC_CUDA(cudaMemcpyAsync(d.ja, ja, sizeof(int) * (size_t)NNZ, cudaMemcpyHostToDevice, c.stream));
    FC_CUDA(cudaMemsetAsync(d.ja + NNZ, 0, sizeof(int) * pad, c.stream));
    FC_CUDA(cudaMemcpyAsync(d.val, val, sizeof(double) * (size_t)NNZ * nb2, cudaMemcpyHostToDevice, c.stream));
    FC_CUDA(cudaMemsetAsync(d.val + (size_t)NNZ * nb2, 0, sizeof(double) * pad, c.stream));
    d.bytes = sizeof(int) * ((size_t)ROW + 1 + NNZ + 2 * pad) + sizeof(double) * ((size_t)NNZ * nb2 + pad);
    // row blocks: B_RB block rows, at most cap blocks (stage <= ~40 KB with 32 block rows, ~80 KB with 64)
    const int B_RB = (nb <= 4 && c.opt.bsr_rb == 64) ? 64 : 32;
    d.blk_rb       = B_RB;
    const double avg = ROW > 0 ? (double)NNZ / ROW : 1.0;
    long long    cap = (long long)(avg * B_RB + 31) / 32 * 32;
    const long long cap_max = (long long)(40 * 1024 * (B_RB / 32)) / (long long)(nb2 * 8 + 4) / 32 * 32;
    if (cap < 64) cap = 64;
    if (cap > cap_max) cap = cap_max;
    if (cap < 32) cap = 32;
    d.blk_cap = (int)cap;
    std::vector<int2> bd;
    int r = 0;
    while (r < ROW) {
        bd.push_back(make_int2(r, ia[r]));
        const long long base = ia[r];
        int             e    = r + 1;
        while (e < ROW && e - r < B_RB && (long long)ia[e + 1] - base <= cap) ++e;
        r = e;
    }
    bd.push_back(make_int2(ROW, ia[ROW]));
    d.nblk    = (int)bd.size() - 1;
    d.blkdesc = dalloc<int2>(bd.size());
    FC_CUDA(cudaMemcpyAsync(d.blkdesc, bd.data(), sizeof(int2) * bd.size(), cudaMemcpyHostToDevice, c.stream));
    d.bytes += sizeof(int2) * bd.size();
    FC_CUDA(cudaStreamSynchronize(c.stream));
    red_partials((size_t)c.sm_count * 16);
}

void bsr_free(DevBSR& d)
{
    dfree(d.ia);
    dfree(d.ja);
    dfree(d.val);
    dfree(d.blkdesc);
    d = DevBSR();
}

__global__ void k_bsr_to_dense(int ROW, int COL, int nb, const int* ia, const int* ja, const double* val,
                               double* dense)
{
    const int    I  = blockIdx.x;
    const size_t ld = (size_t)COL * nb;
    for (int k = ia[I]; k < ia[I + 1]; ++k)
        for (int e = threadIdx.x; e < nb * nb; e += blockDim.x)
            atomicAdd(&dense[((size_t)I * nb + e / nb) * ld + (size_t)ja[k] * nb + e % nb],
                      val[(size_t)k * nb * nb + e]);
}

void bsr_to_dense(const DevBSR& A, double* dense_dev)
{
    FC_LAUNCH(k_bsr_to_dense, A.ROW, 64, 0, A.ROW, A.COL, A.nb, A.ia, A.ja, A.val, dense_dev);
}

} // namespace fc
