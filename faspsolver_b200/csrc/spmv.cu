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

namespace fc {

constexpr int TPB = 256;

struct CsrView {
    const int*    ia;
    const int*    ja;
    const double* val;
    const int*    rowblk;
    const double* diag;
    const int*    dpos;
    const double* l1;
    const double* dinv;
    int           cap;
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
template <int MODE>
__device__ __forceinline__ double row_epilogue(const CsrView& A, const CsrArgs& a, int row,
                                               double acc, bool exact)
{
    double out;
    if (MODE == CSR_MXV) {
        out = acc;
    } else if (MODE == CSR_AXPY) {
        // BlaSpmvCSR.c:509-590: alpha == 1 / -1 / general (temp*alpha added last)
        const double y0 = a.y[row];
        const double al = a.alpha_dev ? *a.alpha_dev : a.alpha;
        if (al == 1.0) out = __dadd_rn(y0, acc);
        else if (al == -1.0) out = __dsub_rn(y0, acc);
        else out = __dadd_rn(y0, __dmul_rn(acc, al));
    } else if (MODE == CSR_RESID || MODE == CSR_RESID_DINV) {
        out = __dsub_rn(a.b[row], acc);
        if (MODE == CSR_RESID_DINV) a.v0_out[row] = __dmul_rn(A.dinv[row], out);
    } else if (MODE == CSR_JACOBI) {
        // ItrSmootherCSR.c:148-170
        const double t = exact ? acc : __dsub_rn(a.b[row], acc);
        const double d = A.diag[row];
        const double u = a.x[row];
        const double w = a.alpha;
        out = (fabs(d) > SMALLREAL)
                  ? __dadd_rn(__dmul_rn(1.0 - w, u), __ddiv_rn(__dmul_rn(w, t), d))
                  : u;
    } else if (MODE == CSR_L1) {
        // ItrSmootherCSR.c:1560-1574
        const double t = exact ? acc : __dsub_rn(a.b[row], acc);
        const double d = A.l1[row];
        const double u = a.x[row];
        out = (fabs(d) > SMALLREAL) ? __dadd_rn(u, __ddiv_rn(t, d)) : u;
    } else if (MODE == CSR_POLY1) {
        // ItrSmootherCSRpoly.c:572-583 : x = rbar ; v0 = k1 rbar ; v1 = k2 rbar - k3 Dinv (A rbar)
        const double rb = a.x[row];
        const double av = __dmul_rn(A.dinv[row], acc);
        a.v0_out[row]   = __dmul_rn(a.k1, rb);
        out             = __dsub_rn(__dmul_rn(a.k2, rb), __dmul_rn(a.k3, av));
        if (a.u_acc) a.u_acc[row] = __dadd_rn(a.u_acc[row], out);   // never used by FASP (ndeg>=2)
    } else { // CSR_POLYJ, ItrSmootherCSRpoly.c:588-608 : x = v1
        const double v1 = a.x[row];
        const double rb = __dmul_rn(__dsub_rn(a.b[row], acc), A.dinv[row]);
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

template <int MODE, bool PATTERN>
__global__ void __launch_bounds__(TPB)
csr_rowblock_kernel(const CsrView A, const CsrArgs a, const int strict, double* partials,
                    unsigned int* ticket)
{
    extern __shared__ double s_prod[];
    __shared__ double        s_red[2][TPB / 32];

    if (a.done != nullptr && *a.done != 0) return;

    const int tid   = threadIdx.x;
    const int r0    = A.rowblk[blockIdx.x];
    const int r1    = A.rowblk[blockIdx.x + 1];
    const int nrows = r1 - r0;
    const int k0    = A.ia[r0];
    const int n     = A.ia[r1] - k0;
    const double* __restrict__ x = a.x;

    double red_dot = 0.0, red_n2 = 0.0;
    const bool want_dot = a.red.dot_out != nullptr;
    const bool want_n2  = a.red.nrm2_out != nullptr;

    if (n <= A.cap) {
        // ---- phase 1: stream the block's nonzeros, park val*x in shared memory
        const int*    __restrict__ ja  = A.ja + k0;
        const double* __restrict__ val = PATTERN ? nullptr : A.val + k0;
#pragma unroll 4
        for (int k = tid; k < n; k += TPB) {
            const int    c  = ld_stream_i32(ja + k);
            const double xv = __ldg(x + c);
            s_prod[k]       = PATTERN ? xv : __dmul_rn(ld_stream_f64(val + k), xv);
        }
        __syncthreads();

        // ---- phase 2: rows
        // lanes per row: 1 while rows average <= 32 nonzeros (exact CPU order), then one
        // lane per ~32 nonzeros as far as the CTA has threads for it
        int lpr = 1;
        if (!strict)
            while (lpr < 32 && nrows * lpr * 2 <= TPB && n > 32 * lpr * nrows) lpr <<= 1;

        if (lpr == 1) {
            if (tid < nrows) {
                const int row = r0 + tid;
                const int ka  = A.ia[row] - k0;
                const int kb  = A.ia[row + 1] - k0;
                double    acc;
                if (ModeTraits<MODE>::smoother) {
                    acc            = a.b[row];
                    const int skip = ModeTraits<MODE>::skipdiag ? ka + A.dpos[row] : -1;
                    for (int k = ka; k < kb; ++k)
                        if (k != skip) acc = __dsub_rn(acc, s_prod[k]);
                } else {
                    acc = 0.0;
                    for (int k = ka; k < kb; ++k) acc = __dadd_rn(acc, s_prod[k]);
                }
                const double out = row_epilogue<MODE>(A, a, row, acc, true);
                if (want_dot) red_dot = out * a.red.dot_with[row];
                if (want_n2) red_n2 = out * out;
            }
        } else {
            const int  g     = tid / lpr;
            const int  gl    = tid - g * lpr;
            const bool valid = g < nrows;
            double     part  = 0.0;
            int        row   = r0;
            if (valid) {
                row            = r0 + g;
                const int ka   = A.ia[row] - k0;
                const int kb   = A.ia[row + 1] - k0;
                const int skip = ModeTraits<MODE>::skipdiag ? ka + A.dpos[row] : -1;
                for (int k = ka + gl; k < kb; k += lpr)
                    if (k != skip) part += s_prod[k];
            }
            for (int off = lpr >> 1; off > 0; off >>= 1)
                part += __shfl_xor_sync(0xffffffffu, part, off);
            if (valid && gl == 0) {
                const double out = row_epilogue<MODE>(A, a, row, part, false);
                if (want_dot) red_dot = out * a.red.dot_with[row];
                if (want_n2) red_n2 = out * out;
            }
        }
    } else {
        // ---- one row longer than the shared-memory capacity: whole CTA on it
        const int row  = r0;
        const int skip = ModeTraits<MODE>::skipdiag ? k0 + A.dpos[row] : -1;
        if (strict) {
            if (tid == 0) {
                double acc = ModeTraits<MODE>::smoother ? a.b[row] : 0.0;
                for (int k = k0; k < k0 + n; ++k) {
                    if (k == skip) continue;
                    const double p = PATTERN ? x[A.ja[k]] : __dmul_rn(A.val[k], x[A.ja[k]]);
                    acc = ModeTraits<MODE>::smoother ? __dsub_rn(acc, p) : __dadd_rn(acc, p);
                }
                const double out = row_epilogue<MODE>(A, a, row, acc, true);
                if (want_dot) red_dot = out * a.red.dot_with[row];
                if (want_n2) red_n2 = out * out;
            }
        } else {
            double part = 0.0;
            for (int k = k0 + tid; k < k0 + n; k += TPB) {
                if (k == skip) continue;
                const double xv = __ldg(x + ld_stream_i32(A.ja + k));
                part += PATTERN ? xv : ld_stream_f64(A.val + k) * xv;
            }
            for (int off = 16; off > 0; off >>= 1)
                part += __shfl_xor_sync(0xffffffffu, part, off);
            if ((tid & 31) == 0) s_red[0][tid >> 5] = part;
            __syncthreads();
            if (tid == 0) {
                double acc = 0.0;
                for (int w = 0; w < TPB / 32; ++w) acc += s_red[0][w];
                const double out = row_epilogue<MODE>(A, a, row, acc, false);
                if (want_dot) red_dot = out * a.red.dot_with[row];
                if (want_n2) red_n2 = out * out;
            }
            __syncthreads();
        }
    }

    // ---- fused grid reduction (deterministic: partials are added in CTA order)
    if (want_dot || want_n2) {
        double v[2] = {red_dot, red_n2};
        grid_reduce<2, 0>(v, partials, ticket, [&](const double* t) {
            if (want_dot) *a.red.dot_out = t[0];
            if (want_n2) *a.red.nrm2_out = t[1];
        });
    }
}

template <int MODE>
static void launch_mode(const DevCSR& A, const CsrView& v, const CsrArgs& a)
{
    Ctx&          c    = ctx();
    const size_t  smem = (size_t)A.blk_cap * sizeof(double);
    double*       part = nullptr;
    unsigned int* tick = nullptr;
    if (a.red.dot_out || a.red.nrm2_out) {
        part = red_partials((size_t)A.nblk);
        tick = red_ticket();
    }
    if (A.val == nullptr)
        FC_LAUNCH((csr_rowblock_kernel<MODE, true>), A.nblk, TPB, smem, v, a, c.opt.strict, part,
                  tick);
    else
        FC_LAUNCH((csr_rowblock_kernel<MODE, false>), A.nblk, TPB, smem, v, a, c.opt.strict,
                  part, tick);
}

void csr_launch(const DevCSR& A, const CsrArgs& a)
{
    if (A.rows == 0) return;
    const bool reads_y = (a.mode == CSR_AXPY || a.mode == CSR_RESID || a.mode >= CSR_JACOBI);
    double     pbytes  = csr_spmv_bytes(A, reads_y);
    if (a.mode == CSR_JACOBI || a.mode == CSR_L1) pbytes += 16.0 * A.rows;   // + u read, d read
    if (a.mode >= CSR_POLY1) pbytes += 16.0 * A.rows;
    ProfScope  prof(a.mode, A.rows, A.nnz, pbytes);
    CsrView v{A.ia, A.ja, A.val, A.rowblk, A.diag, A.dpos, A.l1, A.dinv, A.blk_cap};
    switch (a.mode) {
        case CSR_MXV: launch_mode<CSR_MXV>(A, v, a); break;
        case CSR_AXPY: launch_mode<CSR_AXPY>(A, v, a); break;
        case CSR_RESID: launch_mode<CSR_RESID>(A, v, a); break;
        case CSR_JACOBI:
            if (!A.diag) fail(ERROR_DATA_STRUCTURE, "Jacobi sweep without diagonal data");
            if (A.dup_diag)
                fail(ERROR_DATA_STRUCTURE, "Jacobi sweep: a row stores several diagonal entries");
            launch_mode<CSR_JACOBI>(A, v, a);
            break;
        case CSR_L1:
            if (!A.l1) fail(ERROR_DATA_STRUCTURE, "L1 sweep without l1 row sums");
            launch_mode<CSR_L1>(A, v, a);
            break;
        case CSR_POLY1: launch_mode<CSR_POLY1>(A, v, a); break;
        case CSR_POLYJ: launch_mode<CSR_POLYJ>(A, v, a); break;
        case CSR_RESID_DINV: launch_mode<CSR_RESID_DINV>(A, v, a); break;
        default: fail(ERROR_INPUT_PAR, "csr_launch: unknown mode %d", a.mode);
    }
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

static void build_rowblocks(int rows, const int* ia, int cap, std::vector<int>& rb)
{
    rb.clear();
    rb.reserve((size_t)rows / 128 + 2);
    int r = 0;
    while (r < rows) {
        rb.push_back(r);
        const long long base = ia[r];
        int             e    = r + 1;   // a block always takes its first row, however long
        while (e < rows && e - r < TPB && (long long)ia[e + 1] - base <= cap) ++e;
        r = e;
    }
    rb.push_back(rows);
}

void csr_upload(DevCSR& d, int rows, int cols, long long nnz, const int* ia, const int* ja,
                const double* val, bool pattern_only)
{
    ensure_init();
    Ctx& c = ctx();
    csr_free(d);
    d.rows = rows;
    d.cols = cols;
    d.nnz  = nnz;
    if (nnz >= 2147483647LL) fail(ERROR_MAT_SIZE, "csr_upload: nnz exceeds 32-bit offsets");
    const size_t pad = 8;
    d.ia             = dalloc<int>((size_t)rows + 1);
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
    // row blocks: capacity ~ 256 average rows, between 1024 and 4096 products per CTA
    const double avg = rows > 0 ? (double)nnz / rows : 1.0;
    long long    cap = (long long)(avg * TPB + 255) / 256 * 256;
    if (cap < 1024) cap = 1024;
    if (cap > 4096) cap = 4096;
    d.blk_cap = (int)cap;
    std::vector<int> rb;
    build_rowblocks(rows, ia, d.blk_cap, rb);
    d.nblk   = (int)rb.size() - 1;
    d.rowblk = dalloc<int>(rb.size());
    FC_CUDA(cudaMemcpyAsync(d.rowblk, rb.data(), sizeof(int) * rb.size(), cudaMemcpyHostToDevice,
                            c.stream));
    d.bytes += sizeof(int) * rb.size();
    FC_CUDA(cudaStreamSynchronize(c.stream));   // host staging vectors go out of scope
    red_partials((size_t)d.nblk);
}

void csr_free(DevCSR& d)
{
    dfree(d.ia);
    dfree(d.ja);
    dfree(d.val);
    dfree(d.rowblk);
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
