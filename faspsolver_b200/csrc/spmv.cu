// spmv.cu — CSR row-block kernels for sm_100a: SpMV / residual / Jacobi / L1-Jacobi /
// polynomial-smoother steps, all instances of one "stream the nonzeros, reduce the rows"
// kernel with different per-row epilogues.
//
// Replaces the CPU loops
//   fasp_blas_dcsr_mxv        BlaSpmvCSR.c:242      fasp_blas_dcsr_aAxpy     BlaSpmvCSR.c:494
//   fasp_blas_dcsr_mxv_agg    BlaSpmvCSR.c:438      fasp_blas_dcsr_aAxpy_agg BlaSpmvCSR.c:727
//   fasp_smoother_dcsr_jacobi ItrSmootherCSR.c:98   fasp_smoother_dcsr_L1diag ItrSmootherCSR.c:1509
//   Rr / residual of fasp_smoother_dcsr_poly        ItrSmootherCSRpoly.c:551, :116
//
// Design: the rows are cut once, at upload, into row blocks of <= 256
// rows and <= cap nonzeros. One CTA owns one row block. Phase 1 streams the block's
// contiguous slice of val/ja with fully coalesced loads, gathers x through the read-only
// path (x re-use is served by L1/L2: a 7-point stencil touches 8 B/row of new x), and
// parks val*x in shared memory. Phase 2 gives every row to one thread (short rows) or to a
// 2..32 lane group (long rows) that adds the products. With one thread per row the products
// are added left to right with separate multiply and add roundings, i.e. in exactly the
// order and precision of the sequential CPU loops, so short-row levels reproduce the
// reference bit for bit; lane groups differ by summation order only (<= 1e-14 relative).
// A row longer than cap is its own block and is reduced by the whole CTA.
#include "common.cuh"
#include "reduce.cuh"
#include "pipe.cuh"
#include <algorithm>
#include <thread>

namespace fc {

constexpr int TPB     = 256;
constexpr int CAP_MAX = 2048;   // largest row-block capacity (products per CTA)

// A launch covers the logical units j = 0 .. count-1 (row blocks of the pipelined kernel, rows of the others);
// unit j is lo + j, plus `jump` once j >= split: one contiguous range (the interior of a multi-GPU slab) or two
// (the boundary rows at both ends of the slab), without a second launch.
struct UnitRange {
    int lo, split, jump, count;
    __host__ __device__ __forceinline__ long long map(long long j) const { return lo + j + (j >= split ? jump : 0); }
};

struct CsrView {
    const int*    ia;
    const int*    ja;
    const double* val;
    const int*    rowblk;
    const int2*   blkdesc;
    const double* diag;
    const int*    dpos;
    const double* l1;
    const double* dinv;
    int           cap;
    int           rowwise_max;
};

__device__ __forceinline__ int ld_stream_i32(const int* p)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream_f64(const double* p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// ---- per-row epilogue ---------------------------------------------------------------
// `acc` is sum_k a_ik x_k for the SpMV-like modes. For the smoother modes with exact==true it
// is already t_i = b_i - sum (accumulated by subtraction, as the CPU does); with
// exact==false it is the plain sum and t_i is formed here.
// The vector operands of the epilogue, fetched BEFORE the gather rounds of the row: read where they are used,
// they are one more dependent DRAM access at the end of every row (ncu, L1 sweep: 18 % of all stall samples sat
// on the test of the freshly loaded l1 entry, profiles/r02_l1_sweep_scheduling.txt).
struct EpiPre {
    double p0, p1, p2;
};
template <int MODE>
__device__ __forceinline__ EpiPre epi_prefetch(const CsrView& A, const CsrArgs& a, int row, bool exact)
{
    EpiPre e{0.0, 0.0, 0.0};
    if (MODE == CSR_MXV_DIV) e.p0 = a.div_d[row];
    else if (MODE == CSR_AXPY) e.p0 = a.y[row];
    else if (MODE == CSR_RESID) e.p0 = a.b[row];
    else if (MODE == CSR_RESID_DINV) e.p0 = a.b[row], e.p1 = A.dinv[row];
    else if (MODE == CSR_JACOBI) e.p0 = exact ? 0.0 : a.b[row], e.p1 = A.diag[row], e.p2 = a.x[row];
    else if (MODE == CSR_L1) e.p0 = exact ? 0.0 : a.b[row], e.p1 = A.l1[row], e.p2 = a.x[row];
    else if (MODE == CSR_POLY1) e.p0 = a.x[row], e.p1 = A.dinv[row];
    else if (MODE == CSR_POLYJ) e.p0 = a.b[row], e.p1 = A.dinv[row], e.p2 = a.x[row];
    return e;
}

template <int MODE>
__device__ __forceinline__ double row_epilogue(const CsrView& A, const CsrArgs& a, int row,
                                               double acc, bool exact, const EpiPre& pre)
{
    double out;
    if (MODE == CSR_MXV) {
        out = acc;
    } else if (MODE == CSR_MXV_DIV) {
        out = acc;   // + x_i = s * b_i / d_i of the next level's zero-guess sweep (k_scale_div), same roundings
        const double num = (a.div_s == 1.0) ? out : __dmul_rn(a.div_s, out);
        a.div_out[row]   = (fabs(pre.p0) > SMALLREAL) ? __ddiv_rn(num, pre.p0) : 0.0;
    } else if (MODE == CSR_AXPY) {
        // BlaSpmvCSR.c:509-590: alpha == 1 / -1 / general (temp*alpha added last)
        const double y0 = pre.p0;
        const double al = a.alpha_dev ? *a.alpha_dev : a.alpha;
        if (al == 1.0) out = __dadd_rn(y0, acc);
        else if (al == -1.0) out = __dsub_rn(y0, acc);
        else out = __dadd_rn(y0, __dmul_rn(acc, al));
    } else if (MODE == CSR_RESID || MODE == CSR_RESID_DINV) {
        out = __dsub_rn(pre.p0, acc);
        if (MODE == CSR_RESID_DINV) a.v0_out[row] = __dmul_rn(pre.p1, out);
    } else if (MODE == CSR_JACOBI) {
        // ItrSmootherCSR.c:148-170
        const double t = exact ? acc : __dsub_rn(pre.p0, acc);
        const double d = pre.p1;
        const double u = pre.p2;
        const double w = a.alpha;
        out = (fabs(d) > SMALLREAL)
                  ? __dadd_rn(__dmul_rn(1.0 - w, u), __ddiv_rn(__dmul_rn(w, t), d))
                  : u;
    } else if (MODE == CSR_L1) {
        // ItrSmootherCSR.c:1560-1574
        const double t = exact ? acc : __dsub_rn(pre.p0, acc);
        const double d = pre.p1;
        const double u = pre.p2;
        out = (fabs(d) > SMALLREAL) ? __dadd_rn(u, __ddiv_rn(t, d)) : u;
    } else if (MODE == CSR_POLY1) {
        // ItrSmootherCSRpoly.c:572-583 : x = rbar ; v0 = k1 rbar ; v1 = k2 rbar - k3 Dinv (A rbar)
        const double rb = pre.p0;
        const double av = __dmul_rn(pre.p1, acc);
        a.v0_out[row]   = __dmul_rn(a.k1, rb);
        out             = __dsub_rn(__dmul_rn(a.k2, rb), __dmul_rn(a.k3, av));
        if (a.u_acc) a.u_acc[row] = __dadd_rn(a.u_acc[row], out);   // never used by FASP (ndeg>=2)
    } else { // CSR_POLYJ, ItrSmootherCSRpoly.c:588-608 : x = v1
        const double v1 = pre.p2;
        const double rb = __dmul_rn(__dsub_rn(pre.p0, acc), pre.p1);
        out = __dadd_rn(__dadd_rn(v1, __dmul_rn(a.k5, __dsub_rn(v1, a.v0[row]))),
                        __dmul_rn(a.k4, rb));
        if (a.u_acc) a.u_acc[row] = __dadd_rn(a.u_acc[row], out);
    }
    a.y[row] = out;
    return out;
}

template <int MODE> struct ModeTraits {
    static constexpr bool smoother = (MODE == CSR_JACOBI || MODE == CSR_L1);
    static constexpr bool skipdiag = (MODE == CSR_JACOBI);
};

// 16-byte asynchronous global->shared copy (LDGSTS), L2-only caching: the streamed matrix
// slice goes straight to shared memory without occupying registers, so a CTA has its whole
// slice in flight at once (memory-level parallelism independent of the register budget)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------
// kernel P ("pipelined stream"): persistent CTAs, TMA bulk copies (cp.async.bulk +
// mbarrier) stage the ja / val / ia slices of the NEXT row blocks into a ring of shared-
// memory stages while the current block is gathered and reduced. The streamed bytes in
// flight per SM are (stages - 1) x slice x CTAs/SM, independent of registers and occupancy;
// the only latency left on a block's critical path is the x gather (L1/L2).
// One thread per short row: exact CPU summation order.
// ------------------------------------------------------------------------------------
constexpr int P_MAX_STAGES = 8;
constexpr int P_EPT = 8;            // entries per thread and stage: cap <= P_EPT * threads

struct PipeMeta {
    int r0, nrows, k0, n;
};

template <int MODE, bool PATTERN, int T, int UG, bool TP>
__global__ void __launch_bounds__(T)
csr_pipe_kernel(const CsrView A, const CsrArgs a, const UnitRange ur, const int nstages,
                const int strict, double* partials, unsigned int* ticket)
{
    const int nblk = ur.count;   // logical row blocks of this launch
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) unsigned long long s_bar[P_MAX_STAGES];
    __shared__ PipeMeta s_meta[P_MAX_STAGES];
    __shared__ double   s_red[T / 32];
    constexpr int EPT = P_EPT;   // gathers in flight per thread in the entry-wise path

    if (a.done != nullptr && *a.done != 0) return;

    const int    tid         = threadIdx.x;
    const int    cap         = A.cap;
    const int    G           = gridDim.x;
    const size_t stage_bytes = ((size_t)(cap + 8) * 12 + (size_t)(T + 8) * 4 + 127) & ~(size_t)127;
    auto st_val = [&](int s) { return reinterpret_cast<double*>(s_raw + (size_t)s * stage_bytes); };
    auto st_ja  = [&](int s) { return reinterpret_cast<int*>(st_val(s) + cap + 8); };
    auto st_ia  = [&](int s) { return st_ja(s) + cap + 8; };
    const int2* __restrict__ desc = A.blkdesc;   // desc[b] = {first row, ia[first row]}
    const double* __restrict__ x  = a.x;
    auto gx = [&](int col) { return __ldg(x + col); };
    const int zero = (int)gridDim.y - 1;   // 0 (the grid is one-dimensional), but neither nvcc nor ptxas can fold it

    // thread 0 is the producer: one mbarrier arrival (+ the bulk-copy byte count) per stage use
    auto issue = [&](int blk, int s) {
        const int2 d0 = desc[blk], d1 = desc[blk + 1];
        PipeMeta   m{d0.x, d1.x - d0.x, d0.y, d1.y - d0.y};
        s_meta[s] = m;
        if (m.n <= cap) {
            const int          k0a = m.k0 & ~3;
            const unsigned int na  = (unsigned int)(((m.k0 + m.n + 3) & ~3) - k0a);
            const int          r0a = m.r0 & ~3;
            const unsigned int nra = (unsigned int)(((m.r0 + m.nrows + 1 + 3) & ~3) - r0a);
            mbar_expect_tx(&s_bar[s], na * 4u + (PATTERN ? 0u : na * 8u) + nra * 4u);
            if (na) bulk_g2s(st_ja(s), A.ja + k0a, na * 4u, &s_bar[s]);
            if (!PATTERN && na) bulk_g2s(st_val(s), A.val + k0a, na * 8u, &s_bar[s]);
            bulk_g2s(st_ia(s), A.ia + r0a, nra * 4u, &s_bar[s]);
        } else {
            mbar_expect_tx(&s_bar[s], 0u);   // long row: handled from global memory
        }
    };

    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < nstages; ++s) {
            const long long blk = (long long)blockIdx.x + (long long)s * G;
            if (blk < nblk) issue((int)ur.map(blk), s);
        }
    }
    __syncthreads();

    double     red_dot = 0.0, red_n2 = 0.0;
    const bool want_dot = a.red.dot_out != nullptr;
    const bool want_n2  = a.red.nrm2_out != nullptr;

    int s = 0, ph = 0;   // stage of the current block and the parity of its ring round
    for (long long blk = blockIdx.x; blk < nblk; blk += G) {
        mbar_wait(&s_bar[s], (unsigned int)ph);
        const PipeMeta m = s_meta[s];
        const int r0 = m.r0, nrows = m.nrows, k0 = m.k0, n = m.n;
        if (n <= cap) {
            double*    sp  = st_val(s) + (k0 & 3);
            const int* sj  = st_ja(s) + (k0 & 3);
            const int* sia = st_ia(s) + (r0 & 3);   // sia[i] = ia[r0 + i]
            int lpr = 1;
            if (!strict)
                while (lpr < 32 && nrows * lpr * 2 <= T && n > A.rowwise_max * lpr * nrows) lpr <<= 1;
            if (lpr == 1) {
                // ---- one thread per row, straight from the staged slice: for a fixed position
                // in the row the lanes of a warp gather x at neighbouring columns (consecutive
                // rows of a stencil-like matrix), so the gathers coalesce into few sectors; the
                // products are added left to right with separate roundings = the CPU loop.
                if (tid < nrows) {
                    const int row  = r0 + tid;
                    const int ka   = sia[tid] - k0;
                    const int kb   = sia[tid + 1] - k0;
                    const int skip = ModeTraits<MODE>::skipdiag ? ka + A.dpos[row] : -1;
                    double    acc  = ModeTraits<MODE>::smoother ? a.b[row] : 0.0;
                    const EpiPre pre = epi_prefetch<MODE>(A, a, row, true);
                    constexpr int U = UG;   // gathers in flight per row thread
                    for (int kk = ka; kk < kb; kk += U) {
                        int    col[U];
                        double xv[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) col[u] = (kk + u < kb) ? sj[kk + u] : -1;
                        // TP ("two phases"): all column indices out of shared memory, THEN the gathers back to back.
                        // Left alone, ptxas interleaves the LDS and LDG of a round and, with six scoreboards for 32
                        // loads, may make the sign test of one column index wait for an earlier gather: the L1 sweep
                        // lost 10-17 % in one build with unchanged source (ncu: long-scoreboard 4.3 -> 6.3 per issue,
                        // profiles/r02_l1_sweep_scheduling.txt). With TP every gather address carries a term that is
                        // zero at run time but depends on every column index of the round. It costs the modes whose
                        // natural schedule is good 3-7 %, so it is chosen per mode (option two_phase_mask).
                        int any = 0;
#pragma unroll
                        for (int u = 0; u < U; ++u) any |= col[u];
                        const int dep = TP ? (any & zero) : 0;
#pragma unroll
                        for (int u = 0; u < U; ++u) xv[u] = (col[u] >= 0) ? gx(col[u] + dep) : 0.0;
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int k = kk + u;
                            if (k < kb && k != skip) {
                                const double p = PATTERN ? xv[u] : __dmul_rn(sp[k], xv[u]);
                                acc = ModeTraits<MODE>::smoother ? __dsub_rn(acc, p) : __dadd_rn(acc, p);
                            }
                        }
                    }
                    const double out = row_epilogue<MODE>(A, a, row, acc, true, pre);
                    if (want_dot) red_dot += out * a.red.dot_with[row];
                    if (want_n2) red_n2 += out * out;
                }
            } else {
                // ---- longer rows: entry-wise gather, products parked in the stage, then a
                // group of lpr lanes per row adds them
                for (int base = 0; base < n; base += EPT * T) {
                    int    col[EPT];
                    double xv[EPT];
#pragma unroll
                    for (int e = 0; e < EPT; ++e) {
                        const int k = base + tid + e * T;
                        col[e]      = (k < n) ? sj[k] : -1;
                    }
                    int any = 0;   // two phases, as in the row-wise path
#pragma unroll
                    for (int e = 0; e < EPT; ++e) any |= col[e];
                    const int dep = TP ? (any & zero) : 0;
#pragma unroll
                    for (int e = 0; e < EPT; ++e) xv[e] = (col[e] >= 0) ? gx(col[e] + dep) : 0.0;
#pragma unroll
                    for (int e = 0; e < EPT; ++e) {
                        const int k = base + tid + e * T;
                        if (k < n) sp[k] = PATTERN ? xv[e] : __dmul_rn(sp[k], xv[e]);
                    }
                }
                __syncthreads();
                const int  g     = tid / lpr;
                const int  gl    = tid - g * lpr;
                const bool valid = g < nrows;
                double     part  = 0.0;
                int        row   = r0;
                EpiPre pre{0.0, 0.0, 0.0};
                if (valid) {
                    row            = r0 + g;
                    if (gl == 0) pre = epi_prefetch<MODE>(A, a, row, false);
                    const int ka   = sia[g] - k0;
                    const int kb   = sia[g + 1] - k0;
                    const int skip = ModeTraits<MODE>::skipdiag ? ka + A.dpos[row] : -1;
                    for (int k = ka + gl; k < kb; k += lpr)
                        if (k != skip) part += sp[k];
                }
                for (int off = lpr >> 1; off > 0; off >>= 1)
                    part += __shfl_xor_sync(0xffffffffu, part, off);
                if (valid && gl == 0) {
                    const double out = row_epilogue<MODE>(A, a, row, part, false, pre);
                    if (want_dot) red_dot += out * a.red.dot_with[row];
                    if (want_n2) red_n2 += out * out;
                }
            }
        } else {
            // ---- one row longer than a stage: whole CTA on it, straight from global memory
            const int row  = r0;
            const int skip = ModeTraits<MODE>::skipdiag ? k0 + A.dpos[row] : -1;
            double    part = 0.0;
            if (strict) {
                if (tid == 0) {
                    double acc = ModeTraits<MODE>::smoother ? a.b[row] : 0.0;
                    for (int k = k0; k < k0 + n; ++k) {
                        if (k == skip) continue;
                        const double p = PATTERN ? x[A.ja[k]] : __dmul_rn(A.val[k], x[A.ja[k]]);
                        acc = ModeTraits<MODE>::smoother ? __dsub_rn(acc, p) : __dadd_rn(acc, p);
                    }
                    const double out = row_epilogue<MODE>(A, a, row, acc, true, epi_prefetch<MODE>(A, a, row, true));
                    if (want_dot) red_dot += out * a.red.dot_with[row];
                    if (want_n2) red_n2 += out * out;
                }
            } else {
#pragma unroll 4
                for (int k = k0 + tid; k < k0 + n; k += T) {
                    const double xk = __ldg(x + ld_stream_i32(A.ja + k));
                    const double p  = PATTERN ? xk : ld_stream_f64(A.val + k) * xk;
                    if (k != skip) part += p;
                }
                for (int off = 16; off > 0; off >>= 1)
                    part += __shfl_xor_sync(0xffffffffu, part, off);
                if ((tid & 31) == 0) s_red[tid >> 5] = part;
                __syncthreads();
                if (tid == 0) {
                    double acc = 0.0;
                    for (int w = 0; w < T / 32; ++w) acc += s_red[w];
                    const double out = row_epilogue<MODE>(A, a, row, acc, false, epi_prefetch<MODE>(A, a, row, false));
                    if (want_dot) red_dot += out * a.red.dot_with[row];
                    if (want_n2) red_n2 += out * out;
                }
            }
        }
        __syncthreads();   // stage s fully consumed
        if (tid == 0) {
            const long long refill = blk + (long long)nstages * G;
            if (refill < nblk) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue((int)ur.map(refill), s);
            }
        }
        if (++s == nstages) s = 0, ph ^= 1;
    }

    if (want_dot || want_n2) {
        double v[2] = {red_dot, red_n2};
        grid_reduce<2, 0>(v, partials, ticket, [&](const double* t) {
            if (want_dot) *a.red.dot_out = t[0] + (a.red_add ? a.red_add[0] : 0.0);
            if (want_n2) *a.red.nrm2_out = t[1] + (a.red_add ? a.red_add[1] : 0.0);
        });
    }
}

// ------------------------------------------------------------------------------------
// kernel V ("vector"): LPR lanes per row, no shared memory (the whole L1 serves the x
// gathers). Used for rows that average more than ~12 nonzeros: the coarse levels of a
// classical-AMG hierarchy (19 ... 1400 nonzeros per row). Lanes of a group read consecutive
// entries (coalesced), partial sums are combined with warp shuffles.
// ------------------------------------------------------------------------------------
template <int MODE, bool PATTERN, int LPR, int U>
__global__ void __launch_bounds__(TPB)
csr_vector_kernel(const CsrView A, const CsrArgs a, const UnitRange ur, double* partials,
                  unsigned int* ticket)
{
    if (a.done != nullptr && *a.done != 0) return;
    const long long gt   = (long long)blockIdx.x * TPB + threadIdx.x;
    const int       lane = (int)(gt & (LPR - 1));
    const long long rowl = gt / LPR;
    const bool      valid = rowl < ur.count;
    const int       row   = valid ? (int)ur.map(rowl) : 0;
    const double* __restrict__ x   = a.x;
    const int*    __restrict__ ja  = A.ja;
    const double* __restrict__ val = A.val;
    double part = 0.0;
    EpiPre pre{0.0, 0.0, 0.0};
    if (valid) {
        const int ka   = A.ia[row];
        const int kb   = A.ia[row + 1];
        const int skip = ModeTraits<MODE>::skipdiag ? ka + A.dpos[row] : -1;
        if (lane == 0) pre = epi_prefetch<MODE>(A, a, row, false);   // in flight beside the row pointers
        // U predicated loads per lane are issued together (indices, values, then gathers):
        // a plain unrolled loop would fall into its serial remainder for short trip counts.
        // U = 4, or 8 when a lane owns at least 12 entries of an average row (option vec_u)
        for (int kk = ka + lane; kk < kb; kk += U * LPR) {
            int    col[U];
            double v[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = kk + u * LPR;
                col[u]      = (k < kb) ? ld_stream_i32(ja + k) : -1;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = kk + u * LPR;
                v[u]        = (PATTERN || k >= kb) ? 1.0 : ld_stream_f64(val + k);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = (col[u] >= 0) ? __ldg(x + col[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = kk + u * LPR;
                if (k != skip) part += v[u] * xv[u];
            }
        }
    }
#pragma unroll
    for (int off = LPR >> 1; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    double red_dot = 0.0, red_n2 = 0.0;
    const bool want_dot = a.red.dot_out != nullptr;
    const bool want_n2  = a.red.nrm2_out != nullptr;
    if (valid && lane == 0) {
        const double out = row_epilogue<MODE>(A, a, row, part, false, pre);
        if (want_dot) red_dot = out * a.red.dot_with[row];
        if (want_n2) red_n2 = out * out;
    }
    if (want_dot || want_n2) {
        double v[2] = {red_dot, red_n2};
        grid_reduce<2, 0>(v, partials, ticket, [&](const double* t) {
            if (want_dot) *a.red.dot_out = t[0] + (a.red_add ? a.red_add[0] : 0.0);
            if (want_n2) *a.red.nrm2_out = t[1] + (a.red_add ? a.red_add[1] : 0.0);
        });
    }
}

// ------------------------------------------------------------------------------------
// kernel W ("wide rows"): WPR warps per row (64 / 128 / 256 lanes). The coarse operators of a
// classical-AMG hierarchy have 400 ... 1400 nonzeros per row and only a few thousand rows (a few
// hundred per GPU in the multi-GPU solve): with one warp per row the kernel time is a chain of
// ~10 dependent index->gather rounds; spreading the row over the CTA makes it one or two.
// Warp partials are combined through shared memory in warp order (deterministic).
// ------------------------------------------------------------------------------------
template <int MODE, bool PATTERN, int WPR>
__global__ void __launch_bounds__(TPB)
csr_wide_kernel(const CsrView A, const CsrArgs a, const UnitRange ur, double* partials, unsigned int* ticket)
{
    constexpr int NW  = TPB / 32;
    constexpr int RPC = NW / WPR;   // rows per CTA
    __shared__ double s_part[NW];
    if (a.done != nullptr && *a.done != 0) return;
    const int       warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int       rl   = warp / WPR, wl = warp - rl * WPR;
    const long long rowl = (long long)blockIdx.x * RPC + rl;
    const bool      valid = rowl < ur.count;
    const int       row   = valid ? (int)ur.map(rowl) : 0;
    const double* __restrict__ x   = a.x;
    const int*    __restrict__ ja  = A.ja;
    const double* __restrict__ val = A.val;
    double part = 0.0;
    EpiPre pre{0.0, 0.0, 0.0};
    if (valid) {
        const int ka   = A.ia[row];
        const int kb   = A.ia[row + 1];
        const int skip = ModeTraits<MODE>::skipdiag ? ka + A.dpos[row] : -1;
        if (wl == 0 && lane == 0) pre = epi_prefetch<MODE>(A, a, row, false);
        constexpr int U = 4, STRIDE = 32 * WPR;
        for (int kk = ka + wl * 32 + lane; kk < kb; kk += U * STRIDE) {
            int    col[U];
            double v[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = kk + u * STRIDE;
                col[u]      = (k < kb) ? ld_stream_i32(ja + k) : -1;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = kk + u * STRIDE;
                v[u]        = (PATTERN || k >= kb) ? 1.0 : ld_stream_f64(val + k);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = (col[u] >= 0) ? __ldg(x + col[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = kk + u * STRIDE;
                if (k != skip) part += v[u] * xv[u];
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
    double red_dot = 0.0, red_n2 = 0.0;
    const bool want_dot = a.red.dot_out != nullptr;
    const bool want_n2  = a.red.nrm2_out != nullptr;
    if (valid && wl == 0 && lane == 0) {
        double acc = s_part[rl * WPR];
#pragma unroll
        for (int w = 1; w < WPR; ++w) acc += s_part[rl * WPR + w];
        const double out = row_epilogue<MODE>(A, a, row, acc, false, pre);
        if (want_dot) red_dot = out * a.red.dot_with[row];
        if (want_n2) red_n2 = out * out;
    }
    if (want_dot || want_n2) {
        double v[2] = {red_dot, red_n2};
        grid_reduce<2, 0>(v, partials, ticket, [&](const double* t) {
            if (want_dot) *a.red.dot_out = t[0] + (a.red_add ? a.red_add[0] : 0.0);
            if (want_n2) *a.red.nrm2_out = t[1] + (a.red_add ? a.red_add[1] : 0.0);
        });
    }
}

template <int MODE, bool PATTERN, int WPR>
static void launch_wide(const DevCSR& A, const CsrView& v, const CsrArgs& a, const UnitRange& ur)
{
    constexpr int RPC  = (TPB / 32) / WPR;
    const int     grid = (ur.count + RPC - 1) / RPC;
    double*       part = nullptr;
    unsigned int* tick = nullptr;
    if (a.red.dot_out || a.red.nrm2_out) {
        part = red_partials((size_t)grid);
        tick = red_ticket();
    }
    FC_LAUNCH((csr_wide_kernel<MODE, PATTERN, WPR>), grid, TPB, 0, v, a, ur, part, tick);
}

template <int MODE, bool PATTERN, int LPR, int U>
static void launch_vector_u(const DevCSR& A, const CsrView& v, const CsrArgs& a, const UnitRange& ur);

template <int MODE, bool PATTERN, int LPR>
static void launch_vector(const DevCSR& A, const CsrView& v, const CsrArgs& a, const UnitRange& ur)
{
    const Ctx& c = ctx();
    // 8 loads in flight per lane when the lanes own >= 12 entries of an average row (values only: the pattern-only
    // transfer operators of UA-AMG have 1-2 entries per row)
    const bool deep = !PATTERN && (c.opt.vec_u == 8 || (c.opt.vec_u == 0 && A.rows > 0 && (double)A.nnz >= 12.0 * LPR * A.rows));
    if (deep) launch_vector_u<MODE, PATTERN, LPR, PATTERN ? 4 : 8>(A, v, a, ur);
    else launch_vector_u<MODE, PATTERN, LPR, 4>(A, v, a, ur);
}

template <int MODE, bool PATTERN, int LPR, int U>
static void launch_vector_u(const DevCSR& A, const CsrView& v, const CsrArgs& a, const UnitRange& ur)
{
    const long long threads = (long long)ur.count * LPR;
    const int       grid    = (int)((threads + TPB - 1) / TPB);
    double*         part    = nullptr;
    unsigned int*   tick    = nullptr;
    if (a.red.dot_out || a.red.nrm2_out) {
        part = red_partials((size_t)grid);
        tick = red_ticket();
    }
    FC_LAUNCH((csr_vector_kernel<MODE, PATTERN, LPR, U>), grid, TPB, 0, v, a, ur, part, tick);
}

// persistent grid: as many CTAs per SM as the stage rings allow
template <int MODE, bool PATTERN, int T, int UG = 8, bool TP = false>
static void launch_pipe(const DevCSR& A, const CsrView& v, const CsrArgs& a, const UnitRange& ur, double* part,
                        unsigned int* tick)
{
    Ctx&         c     = ctx();
    const size_t stage = ((size_t)(A.blk_cap + 8) * 12 + (size_t)(T + 8) * 4 + 127) & ~(size_t)127;
    int          nst   = c.opt.pipe_stages;
    if (nst < 2) nst = 2;
    if (nst > P_MAX_STAGES) nst = P_MAX_STAGES;
    const size_t smem = stage * nst;
    static bool  attr_set = false;
    if (!attr_set) {
        FC_CUDA(cudaFuncSetAttribute(csr_pipe_kernel<MODE, PATTERN, T, UG, TP>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    int per_sm = (int)((size_t)(226 * 1024) / (smem + 2048));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > c.opt.pipe_ctas) per_sm = c.opt.pipe_ctas;
    if (per_sm > 2048 / T) per_sm = 2048 / T;
    // Interior rows of a multi-GPU slab (side stream): the CTAs are persistent, so a full grid would hold every SM
    // until the kernel ends and the ghost push / flag barrier on the main stream could not start beside it. Leave
    // one CTA slot per SM free for them.
    if (c.launch_stream == c.side && c.side != nullptr && per_sm > 2) per_sm -= 1;
    long long grid = (long long)c.sm_count * per_sm;
    if (grid > ur.count) grid = ur.count;
    FC_LAUNCH((csr_pipe_kernel<MODE, PATTERN, T, UG, TP>), (int)grid, T, smem, v, a, ur, nst,
              c.opt.strict, part, tick);
}

// which: 0 all rows, 1 interior rows only (no ghost columns), 2 the rows around the interior
template <int MODE, bool PATTERN>
static void launch_pattern(const DevCSR& A, const CsrView& v, const CsrArgs& a, int which)
{
    Ctx& c = ctx();
    if (A.vec_lpr > 0 && !c.opt.strict) {
        UnitRange ur{0, A.rows, 0, A.rows};
        if (which == 1) ur = UnitRange{A.int_row0, A.int_row1 - A.int_row0, 0, A.int_row1 - A.int_row0};
        if (which == 2) ur = UnitRange{0, A.int_row0, A.int_row1 - A.int_row0, A.rows - (A.int_row1 - A.int_row0)};
        if (ur.count <= 0) return;
        switch (A.vec_lpr) {
            case 4: launch_vector<MODE, PATTERN, 4>(A, v, a, ur); return;
            case 8: launch_vector<MODE, PATTERN, 8>(A, v, a, ur); return;
            case 16: launch_vector<MODE, PATTERN, 16>(A, v, a, ur); return;
            case 64: launch_wide<MODE, PATTERN, 2>(A, v, a, ur); return;
            case 128: launch_wide<MODE, PATTERN, 4>(A, v, a, ur); return;
            case 256: launch_wide<MODE, PATTERN, 8>(A, v, a, ur); return;
            default: launch_vector<MODE, PATTERN, 32>(A, v, a, ur); return;
        }
    }
    UnitRange ur{0, A.nblk, 0, A.nblk};
    if (which == 1) ur = UnitRange{A.int_blk0, A.int_blk1 - A.int_blk0, 0, A.int_blk1 - A.int_blk0};
    if (which == 2) ur = UnitRange{0, A.int_blk0, A.int_blk1 - A.int_blk0, A.nblk - (A.int_blk1 - A.int_blk0)};
    if (ur.count <= 0) return;
    double*       part = nullptr;
    unsigned int* tick = nullptr;
    if (a.red.dot_out || a.red.nrm2_out) {
        part = red_partials((size_t)ur.count);
        tick = red_ticket();
    }
    switch (A.blk_tpb) {
        case 64: launch_pipe<MODE, PATTERN, 64>(A, v, a, ur, part, tick); return;
        case 128:
            // rows of 6+ entries: 16 gathers in flight per row thread (one round for a 7-point row, two
            // for the 19-entry rows of the first coarse level instead of three) at 64 registers, still
            // 8 CTAs per SM; measured +4 % (level 0) and +10 % (level 1) over the 8-deep variant, which
            // the short rows of the transfer operators keep
            // two-phase gather rounds per mode (bit MODE of the option): measured on the 7-pt 256^3 hierarchy,
            // profiles/r02_l1_sweep_scheduling.txt
            if (c.opt.gather16_min_avg > 0 && A.rows > 0 && (double)A.nnz >= (double)c.opt.gather16_min_avg * A.rows) {
                if ((c.opt.two_phase_mask >> MODE) & 1) launch_pipe<MODE, PATTERN, 128, 16, true>(A, v, a, ur, part, tick);
                else launch_pipe<MODE, PATTERN, 128, 16, false>(A, v, a, ur, part, tick);
            } else {
                if ((c.opt.two_phase_mask >> MODE) & 1) launch_pipe<MODE, PATTERN, 128, 8, true>(A, v, a, ur, part, tick);
                else launch_pipe<MODE, PATTERN, 128, 8, false>(A, v, a, ur, part, tick);
            }
            return;
        default: launch_pipe<MODE, PATTERN, 256>(A, v, a, ur, part, tick); return;
    }
}

template <int MODE>
static void launch_mode(const DevCSR& A, const CsrView& v, const CsrArgs& a, int which)
{
    if (A.val == nullptr) launch_pattern<MODE, true>(A, v, a, which);
    else launch_pattern<MODE, false>(A, v, a, which);
}

static void launch_any(const DevCSR& A, const CsrView& v, const CsrArgs& a, int which)
{
    switch (a.mode) {
        case CSR_MXV: launch_mode<CSR_MXV>(A, v, a, which); break;
        case CSR_AXPY: launch_mode<CSR_AXPY>(A, v, a, which); break;
        case CSR_RESID: launch_mode<CSR_RESID>(A, v, a, which); break;
        case CSR_JACOBI: launch_mode<CSR_JACOBI>(A, v, a, which); break;
        case CSR_L1: launch_mode<CSR_L1>(A, v, a, which); break;
        case CSR_POLY1: launch_mode<CSR_POLY1>(A, v, a, which); break;
        case CSR_POLYJ: launch_mode<CSR_POLYJ>(A, v, a, which); break;
        case CSR_RESID_DINV: launch_mode<CSR_RESID_DINV>(A, v, a, which); break;
        case CSR_MXV_DIV: launch_mode<CSR_MXV_DIV>(A, v, a, which); break;
        default: fail(ERROR_INPUT_PAR, "csr_launch: unknown mode %d", a.mode);
    }
}

void csr_launch(const DevCSR& A, const CsrArgs& a_in)
{
    if (A.rows == 0) {
        // a rank that owns no row of this operator still takes part in the collectives around the kernel
        if (A.halo) halo_exchange(*A.halo, const_cast<double*>(a_in.x), a_in.done, a_in.conditional);
        vec_reduce(a_in.y, 0, a_in.red, a_in.done);   // zero contribution (+ the all-reduce)
        return;
    }
    Ctx&       c       = ctx();
    const bool reads_y = (a_in.mode == CSR_AXPY || a_in.mode == CSR_RESID ||
                          (a_in.mode >= CSR_JACOBI && a_in.mode != CSR_MXV_DIV));
    double     pbytes  = csr_spmv_bytes(A, reads_y);
    if (a_in.mode == CSR_JACOBI || a_in.mode == CSR_L1) pbytes += 16.0 * A.rows;   // + u read, d read
    if (a_in.mode >= CSR_POLY1) pbytes += 16.0 * A.rows;   // incl. CSR_MXV_DIV: + d read, x written
    const CsrArgs& a = a_in;
    switch (a.mode) {
        case CSR_JACOBI:
            if (!A.diag) fail(ERROR_DATA_STRUCTURE, "Jacobi sweep without diagonal data");
            if (A.dup_diag) fail(ERROR_DATA_STRUCTURE, "Jacobi sweep: a row stores several diagonal entries");
            break;
        case CSR_L1:
            if (!A.l1) fail(ERROR_DATA_STRUCTURE, "L1 sweep without l1 row sums");
            break;
        default: break;
    }
    CsrView v{A.ia, A.ja, A.val, A.rowblk, A.blkdesc, A.diag, A.dpos, A.l1, A.dinv, A.blk_cap, c.opt.rowwise_max};
    // Multi-GPU: the rows of the slab that touch no ghost column (its interior: all but the planes next to the
    // neighbours) start on a second stream at once, the ghost exchange (push over NVLink + flag barrier) runs
    // beside them, and only the boundary rows wait for it. In a captured graph this is a fork / join.
    const bool use_vec = A.vec_lpr > 0 && !c.opt.strict;
    const int  n_int   = use_vec ? A.int_row1 - A.int_row0 : A.int_blk1 - A.int_blk0;
    const int  n_bnd   = (use_vec ? A.rows : A.nblk) - n_int;
    const bool exch    = A.halo != nullptr && !a.skip_halo;
    const bool split   = exch && c.opt.overlap && c.side != nullptr && n_int > 0 && n_bnd > 0 &&
                       A.int_row1 - A.int_row0 >= c.opt.overlap_min_rows;
    if (!split) {
        if (exch) halo_exchange(*A.halo, const_cast<double*>(a.x), a.done, a.conditional);
        ProfScope prof(a.conditional ? a.mode + 50 : a.mode, A.rows, A.nnz, pbytes);
        launch_any(A, v, a, 0);
        reduce_finish(a.red, a.done);
        return;
    }
    ProfScope  prof(a.conditional ? a.mode + 50 : a.mode, A.rows, A.nnz, pbytes);
    const bool has_red = a.red.dot_out != nullptr || a.red.nrm2_out != nullptr;
    FC_CUDA(cudaEventRecord(c.ev_fork, c.stream));
    FC_CUDA(cudaStreamWaitEvent(c.side, c.ev_fork, 0));
    {
        CsrArgs ai = a;
        if (a.red.dot_out) ai.red.dot_out = c.side_tot;
        if (a.red.nrm2_out) ai.red.nrm2_out = c.side_tot + 1;
        c.launch_stream = c.side;
        try {
            launch_any(A, v, ai, 1);
        } catch (...) {
            c.launch_stream = c.stream;
            throw;
        }
        c.launch_stream = c.stream;
    }
    FC_CUDA(cudaEventRecord(c.ev_join, c.side));
    halo_exchange(*A.halo, const_cast<double*>(a.x), a.done, a.conditional);
    CsrArgs ab = a;
    if (has_red) {   // the boundary kernel adds the interior's totals: it must run after the interior
        ab.red_add = c.side_tot;
        FC_CUDA(cudaStreamWaitEvent(c.stream, c.ev_join, 0));
        launch_any(A, v, ab, 2);
    } else {
        launch_any(A, v, ab, 2);
        FC_CUDA(cudaStreamWaitEvent(c.stream, c.ev_join, 0));
    }
    reduce_finish(a.red, a.done);
}

// ------------------------------------------------------------------------------------
// upload + side data
// ------------------------------------------------------------------------------------

// thread per row, storage order: diagonal entry as fasp_smoother_dcsr_jacobi sees it (the
// last stored (i,i) entry, ItrSmootherCSR.c:149-158), its offset in the row, duplicates flag
__global__ void csr_diag_kernel(int rows, const int* ia, const int* ja, const double* val,
                                double* diag, int* dpos, int* dup)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    double d = 0.0;
    int    p = -1, cnt = 0;
    for (int k = ia[i]; k < ia[i + 1]; ++k)
        if (ja[k] == i) {
            d = val ? val[k] : 1.0;
            p = k - ia[i];
            ++cnt;
        }
    diag[i] = d;
    dpos[i] = p;
    if (cnt > 1) *dup = 1;
}

// sum_k |a_ik| left to right (ItrSmootherCSR.c:1565-1569, ItrSmootherCSRpoly.c:453-456)
__global__ void csr_l1_kernel(int rows, const int* ia, const double* val, double* l1)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    double s = 0.0;
    for (int k = ia[i]; k < ia[i + 1]; ++k) s = __dadd_rn(s, val ? fabs(val[k]) : 1.0);
    l1[i] = s;
}

// 1/a_ii with the FIRST stored diagonal entry (Diaginv, ItrSmootherCSRpoly.c:392-410)
__global__ void csr_dinv_kernel(int rows, const int* ia, const int* ja, const double* val,
                                double* dinv)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    double d = 0.0;
    for (int k = ia[i]; k < ia[i + 1]; ++k)
        if (ja[k] == i) {
            d = val ? val[k] : 1.0;
            break;
        }
    dinv[i] = 1.0 / d;
}

__global__ void rowmax_kernel(int rows, const double* l1, const double* dinv, double* blockmax)
{
    __shared__ double s[256];
    double            m = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
        const double t = __dmul_rn(l1[i], dinv[i]);
        m              = (m > t) ? m : t;   // MAX(norm,temp) of ItrSmootherCSRpoly.c:462
    }
    s[threadIdx.x] = m;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off && s[threadIdx.x + off] > s[threadIdx.x])
            s[threadIdx.x] = s[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) blockmax[blockIdx.x] = s[0];
}

static void build_rowblocks(int rows, const int* ia, int cap, int maxrows, std::vector<int>& rb)
{
    rb.clear();
    rb.reserve((size_t)rows / 128 + 2);
    int r = 0;
    while (r < rows) {
        rb.push_back(r);
        const long long base = ia[r];
        int             e    = r + 1;   // a block always takes its first row, however long
        while (e < rows && e - r < maxrows && (long long)ia[e + 1] - base <= cap) ++e;
        r = e;
    }
    rb.push_back(rows);
}

void csr_upload(DevCSR& d, int rows, int cols, long long nnz, const int* ia, const int* ja,
                const double* val, bool pattern_only, int ghost_col0)
{
    ensure_init();
    Ctx& c = ctx();
    csr_free(d);
    d.rows = rows;
    d.cols = cols;
    d.nnz  = nnz;
    if (nnz >= 2147483647LL) fail(ERROR_MAT_SIZE, "csr_upload: nnz exceeds 32-bit offsets");
    const size_t pad = 8;
    // kernel choice: short rows -> pipelined stream kernel (exact CPU summation order); longer
    // rows -> vector kernel with LPR lanes per row
    const double avg0 = rows > 0 ? (double)nnz / rows : 1.0;
    d.vec_lpr         = 0;
    const int vmin    = c.opt.vec_min_avg;
    if (vmin > 0 && avg0 >= vmin)
        d.vec_lpr = c.opt.vec_lpr > 0 ? c.opt.vec_lpr : (avg0 < 48 ? 4 : (avg0 < 96 ? 8 : (avg0 < 384 ? 16 : 32)));
    // few long rows (coarse levels, above all the per-GPU slabs of a multi-GPU solve): spread a
    // row over more lanes so that the kernel is not a chain of dependent gather rounds
    if (d.vec_lpr > 0 && c.opt.vec_lpr <= 0)
        while (d.vec_lpr < 256 && (long long)rows * d.vec_lpr < c.opt.wide_threads && avg0 / d.vec_lpr >= 8.0)
            d.vec_lpr *= 2;
    // Rows that go to the vector kernel are summed by lane groups anyway (order differs from the
    // CPU loop by construction), so their entries are sorted by column at upload: consecutive lanes
    // then gather neighbouring x entries and share 32-byte sectors (the unsorted RAP output of
    // the host setup scatters them). Short rows keep FASP's storage order = the CPU summation order.
    std::vector<int>    sja;
    std::vector<double> sval;
    if (d.vec_lpr > 0 && c.opt.sort_rows && !c.opt.strict && nnz > 0) {
        sja.assign(ja, ja + nnz);
        if (!pattern_only && val) sval.assign(val, val + nnz);
        const unsigned nt = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t]() {
                std::vector<std::pair<int, double>> tmp;
                for (int i = (int)(((long long)rows * t) / nt); i < (int)(((long long)rows * (t + 1)) / nt); ++i) {
                    const int a = ia[i], b = ia[i + 1];
                    if (b - a < 2 || std::is_sorted(sja.begin() + a, sja.begin() + b)) continue;
                    tmp.resize(b - a);
                    for (int k = a; k < b; ++k) tmp[k - a] = {sja[k], sval.empty() ? 0.0 : sval[k]};
                    std::stable_sort(tmp.begin(), tmp.end(),
                                     [](const std::pair<int, double>& x, const std::pair<int, double>& y) {
                                         return x.first < y.first;
                                     });
                    for (int k = a; k < b; ++k) {
                        sja[k] = tmp[k - a].first;
                        if (!sval.empty()) sval[k] = tmp[k - a].second;
                    }
                }
            });
        for (auto& t : th) t.join();
        ja = sja.data();
        if (!sval.empty()) val = sval.data();
    }
    d.ia             = dalloc<int>((size_t)rows + 1 + pad);
    FC_CUDA(cudaMemsetAsync(d.ia + rows + 1, 0, sizeof(int) * pad, c.stream));
    d.ja             = dalloc<int>((size_t)nnz + pad);
    FC_CUDA(cudaMemcpyAsync(d.ia, ia, sizeof(int) * ((size_t)rows + 1), cudaMemcpyHostToDevice,
                            c.stream));
    FC_CUDA(cudaMemcpyAsync(d.ja, ja, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice,
                            c.stream));
    FC_CUDA(cudaMemsetAsync(d.ja + nnz, 0, sizeof(int) * pad, c.stream));
    d.bytes = sizeof(int) * ((size_t)rows + 1 + nnz + pad);
    if (!pattern_only && val != nullptr) {
        d.val = dalloc<double>((size_t)nnz + pad);
        FC_CUDA(cudaMemcpyAsync(d.val, val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice,
                                c.stream));
        FC_CUDA(cudaMemsetAsync(d.val + nnz, 0, sizeof(double) * pad, c.stream));
        d.bytes += sizeof(double) * ((size_t)nnz + pad);
    }
    FC_CUDA(cudaStreamSynchronize(c.stream));   // sorted staging copies go out of scope below
    // row blocks: at most T rows (T = threads of the pipelined kernel) and cap products, cap ~ T
    // average rows and <= 8 T (12 B of shared memory per product and stage)
    const double avg = rows > 0 ? (double)nnz / rows : 1.0;
    int T = c.opt.pipe_tpb;
    if (T != 64 && T != 128 && T != 256) T = 128;
    d.blk_tpb     = T;
    long long cap = (long long)(avg * T + 63) / 64 * 64;
    if (cap < 4 * T) cap = 4 * T;
    int mult = c.opt.pipe_cap_mult;
    if (avg >= 24.0 && mult == 16) mult = 24;   // measured: 35-entry rows want ~90 rows per block, 19-entry rows ~108
    if (mult < 4) mult = 4;
    if (mult > 32) mult = 32;
    if (cap > (long long)mult * T) cap = (long long)mult * T;
    d.blk_cap = (int)cap;
    std::vector<int> rb;
    build_rowblocks(rows, ia, d.blk_cap, d.blk_tpb, rb);
    d.nblk   = (int)rb.size() - 1;
    d.rowblk = dalloc<int>(rb.size());
    FC_CUDA(cudaMemcpyAsync(d.rowblk, rb.data(), sizeof(int) * rb.size(), cudaMemcpyHostToDevice,
                            c.stream));
    d.bytes += sizeof(int) * rb.size();
    std::vector<int2> bd(rb.size());
    for (size_t i = 0; i < rb.size(); ++i) bd[i] = make_int2(rb[i], ia[rb[i]]);
    d.blkdesc = dalloc<int2>(bd.size());
    FC_CUDA(cudaMemcpyAsync(d.blkdesc, bd.data(), sizeof(int2) * bd.size(), cudaMemcpyHostToDevice,
                            c.stream));
    d.bytes += sizeof(int2) * bd.size();
    FC_CUDA(cudaStreamSynchronize(c.stream));   // host staging vectors go out of scope
    red_partials((size_t)d.nblk);
    // multi-GPU slab: the longest run of rows that reference no ghost column (for a z-slab of a grid: everything
    // but the planes next to the neighbours), and the row blocks that lie entirely inside it
    d.int_row0 = d.int_row1 = d.int_blk0 = d.int_blk1 = 0;
    if (ghost_col0 >= 0 && rows > 0) {
        int best0 = 0, best1 = 0, run0 = 0;
        for (int i = 0; i <= rows; ++i) {
            bool ghost = (i == rows);
            if (!ghost)
                for (int k = ia[i]; k < ia[i + 1]; ++k)
                    if (ja[k] >= ghost_col0) {
                        ghost = true;
                        break;
                    }
            if (ghost) {
                if (i - run0 > best1 - best0) best0 = run0, best1 = i;
                run0 = i + 1;
            }
        }
        d.int_row0 = best0, d.int_row1 = best1;
        int b0 = 0;
        while (b0 < d.nblk && rb[b0] < best0) ++b0;
        int b1 = b0;
        while (b1 < d.nblk && rb[b1 + 1] <= best1) ++b1;
        d.int_blk0 = b0, d.int_blk1 = b1;
        // reduction scratch of the side stream, reserved before any graph capture
        c.launch_stream = c.side;
        const size_t vgrid = (size_t)rows * (size_t)(d.vec_lpr > 0 ? d.vec_lpr : 1) / TPB + 2;
        red_partials(std::max((size_t)d.nblk, vgrid));
        c.launch_stream = c.stream;
    }
}

void csr_free(DevCSR& d)
{
    dfree(d.ia);
    dfree(d.ja);
    dfree(d.val);
    dfree(d.rowblk);
    dfree(d.blkdesc);
    dfree(d.diag);
    dfree(d.dpos);
    dfree(d.l1);
    dfree(d.dinv);
    d = DevCSR();
}

void csr_ensure_diag(DevCSR& d)
{
    if (d.diag || d.rows == 0) return;
    Ctx& c = ctx();
    d.diag = dalloc<double>(d.rows);
    d.dpos = dalloc<int>(d.rows);
    int* dup = dalloc<int>(1);
    FC_CUDA(cudaMemsetAsync(dup, 0, sizeof(int), c.stream));
    FC_LAUNCH(csr_diag_kernel, (d.rows + 255) / 256, 256, 0, d.rows, d.ia, d.ja, d.val, d.diag,
              d.dpos, dup);
    int h = 0;
    FC_CUDA(cudaMemcpyAsync(&h, dup, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    FC_CUDA(cudaStreamSynchronize(c.stream));
    dfree(dup);
    d.dup_diag = (h != 0);
    d.bytes += (size_t)d.rows * 12;
}

void csr_ensure_l1(DevCSR& d)
{
    if (d.l1 || d.rows == 0) return;
    d.l1 = dalloc<double>(d.rows);
    FC_LAUNCH(csr_l1_kernel, (d.rows + 255) / 256, 256, 0, d.rows, d.ia, d.val, d.l1);
    d.bytes += (size_t)d.rows * 8;
}

void csr_ensure_dinv(DevCSR& d)
{
    if (d.dinv || d.rows == 0) return;
    d.dinv = dalloc<double>(d.rows);
    FC_LAUNCH(csr_dinv_kernel, (d.rows + 255) / 256, 256, 0, d.rows, d.ia, d.ja, d.val, d.dinv);
    d.bytes += (size_t)d.rows * 8;
}

double csr_dinv_a_norminf(DevCSR& d)
{
    csr_ensure_l1(d);
    csr_ensure_dinv(d);
    Ctx&      c  = ctx();
    const int nb = 296;
    double*   bm = dalloc<double>(nb);
    FC_LAUNCH(rowmax_kernel, nb, 256, 0, d.rows, d.l1, d.dinv, bm);
    std::vector<double> h(nb);
    FC_CUDA(cudaMemcpyAsync(h.data(), bm, sizeof(double) * nb, cudaMemcpyDeviceToHost, c.stream));
    FC_CUDA(cudaStreamSynchronize(c.stream));
    dfree(bm);
    double m = 0.0;
    for (double v : h) m = (m > v) ? m : v;
    return m;
}

} // namespace fc
