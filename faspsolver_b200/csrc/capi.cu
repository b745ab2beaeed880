// capi.cu — the extern "C" boundary of libfasp_cuda (declared in include/fasp_cuda.h).
//
// Every function here takes FASP's own structs / plain pointers, translates internal
// fc::Error exceptions into FASP status codes and never calls exit(). There is no CPU
// fallback: without a CUDA device every compute entry point returns ERROR_SOLVER_MISC and
// fasp_cuda_last_error() says why.
#include "common.cuh"
#include "amg.cuh"
#include "krylov.cuh"
#include "solver.cuh"
#include <dlfcn.h>
#include <limits>
#include <new>

using namespace fc;

#define API_TRY try {
#define API_CATCH(retexpr)                                                                 \
    }                                                                                      \
    catch (const fc::Error& e)                                                             \
    {                                                                                      \
        fc::set_last_error(e.msg);                                                         \
        if (getenv("FASP_CUDA_VERBOSE")) fprintf(stderr, "### libfasp_cuda: %s\n", e.msg.c_str()); \
        int code__ = e.code;                                                               \
        (void)code__;                                                                      \
        return retexpr;                                                                    \
    }                                                                                      \
    catch (const std::bad_alloc&)                                                          \
    {                                                                                      \
        fc::set_last_error("host allocation failed");                                      \
        int code__ = ERROR_ALLOC_MEM;                                                      \
        (void)code__;                                                                      \
        return retexpr;                                                                    \
    }                                                                                      \
    catch (const std::exception& e)                                                        \
    {                                                                                      \
        fc::set_last_error(e.what());                                                      \
        int code__ = ERROR_UNKNOWN;                                                        \
        (void)code__;                                                                      \
        return retexpr;                                                                    \
    }

struct fasp_cuda_csr_s {
    DevCSR m;
};
struct fasp_cuda_amg_s {
    Amg* h;
};

// ------------------------------------------------------------------------------------
// small RAII helpers for host-pointer entry points
// ------------------------------------------------------------------------------------
namespace {
struct DVec {
    double* p = nullptr;
    size_t  n = 0;
    explicit DVec(size_t n_) : n(n_) { p = dalloc<double>(n_ ? n_ : 1); }
    DVec(const double* h, size_t n_) : n(n_)
    {
        p = dalloc<double>(n_ ? n_ : 1);
        if (n_) FC_CUDA(cudaMemcpyAsync(p, h, sizeof(double) * n_, cudaMemcpyHostToDevice, ctx().stream));
    }
    ~DVec() { dfree(p); }
    void to_host(double* h)
    {
        if (n) FC_CUDA(cudaMemcpyAsync(h, p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx().stream));
        FC_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    DVec(const DVec&)            = delete;
    DVec& operator=(const DVec&) = delete;
};
struct TmpCSR {
    DevCSR m;
    TmpCSR(const dCSRmat* A, bool pattern = false)
    {
        if (!A) fail(ERROR_INPUT_PAR, "null matrix");
        csr_upload(m, A->row, A->col, A->nnz, A->IA, A->JA, A->val, pattern);
    }
    ~TmpCSR() { csr_free(m); }
};

void check_csr(const dCSRmat* A)
{
    if (!A || A->row < 0 || A->col < 0 || A->nnz < 0 || !A->IA || (A->nnz > 0 && !A->JA))
        fail(ERROR_DATA_STRUCTURE, "invalid dCSRmat");
}

int host_spmv(const dCSRmat* A, int mode, double alpha, const double* x, double* y, bool pattern)
{
    ensure_init();
    check_csr(A);
    TmpCSR  dA(A, pattern);
    DVec    dx(x, A->col);
    DVec    dy((size_t)A->row);
    if (mode == CSR_AXPY)
        FC_CUDA(cudaMemcpyAsync(dy.p, y, sizeof(double) * A->row, cudaMemcpyHostToDevice,
                                ctx().stream));
    CsrArgs a;
    a.mode  = mode;
    a.alpha = alpha;
    a.x     = dx.p;
    a.y     = dy.p;
    csr_launch(dA.m, a);
    dy.to_host(y);
    return FASP_SUCCESS;
}
} // namespace

extern "C" {

// ------------------------------------------------------------------------------------
// library
// ------------------------------------------------------------------------------------
INT fasp_cuda_abi_check(size_t s_csr, size_t s_amgdata, size_t s_amgparam)
{
    if (s_csr == sizeof(dCSRmat) && s_amgdata == sizeof(AMG_data) && s_amgparam == sizeof(AMG_param))
        return FASP_SUCCESS;
    char buf[256];
    snprintf(buf, sizeof(buf),
             "ABI mismatch: caller dCSRmat/AMG_data/AMG_param = %zu/%zu/%zu bytes, library "
             "%zu/%zu/%zu (OpenMP vs sequential FASP build?)",
             s_csr, s_amgdata, s_amgparam, sizeof(dCSRmat), sizeof(AMG_data), sizeof(AMG_param));
    set_last_error(buf);
    return ERROR_DATA_STRUCTURE;
}

const char* fasp_cuda_last_error(void) { return fc::last_error(); }

INT fasp_cuda_init(int device)
{
    API_TRY
    if (ctx().inited && ctx().device != device)
        fail(ERROR_INPUT_PAR, "library already initialised on device %d", ctx().device);
    ctx().device = device;
    if (!getenv("FASP_CUDA_DEVICE")) {
        char b[16];
        snprintf(b, sizeof(b), "%d", device);
        setenv("FASP_CUDA_DEVICE", b, 1);
    }
    ensure_init();
    return FASP_SUCCESS;
    API_CATCH(code__)
}

long long fasp_cuda_launch_count(void) { return ctx().launches; }
void      fasp_cuda_launch_count_reset(void) { ctx().launches = 0; }

static int* option_slot(const char* key)
{
    Options& o = ctx().opt;
    if (!key) return nullptr;
    if (!strcmp(key, "strict")) return &o.strict;
    if (!strcmp(key, "coarse_dense")) return &o.coarse_dense;
    if (!strcmp(key, "coarse_dense_max")) return &o.coarse_dense_max;
    if (!strcmp(key, "graph")) return &o.graph;
    if (!strcmp(key, "zero_guess")) return &o.zero_guess;
    if (!strcmp(key, "lookahead")) return &o.lookahead;
    if (!strcmp(key, "profile")) return &o.profile;
    if (!strcmp(key, "vec_min_avg")) return &o.vec_min_avg;
    if (!strcmp(key, "pipe")) return &o.pipe;
    if (!strcmp(key, "pipe_ctas")) return &o.pipe_ctas;
    if (!strcmp(key, "pipe_stages")) return &o.pipe_stages;
    if (!strcmp(key, "pipe_tpb")) return &o.pipe_tpb;
    if (!strcmp(key, "pipe_cap_mult")) return &o.pipe_cap_mult;
    if (!strcmp(key, "sort_rows")) return &o.sort_rows;
    if (!strcmp(key, "host_register")) return &o.host_register;
    if (!strcmp(key, "rowwise_max")) return &o.rowwise_max;
    if (!strcmp(key, "gather16_min_avg")) return &o.gather16_min_avg;
    if (!strcmp(key, "two_phase_mask")) return &o.two_phase_mask;
    if (!strcmp(key, "wide_threads")) return &o.wide_threads;
    if (!strcmp(key, "vec_lpr")) return &o.vec_lpr;
    if (!strcmp(key, "vec_u")) return &o.vec_u;
    if (!strcmp(key, "fuse_restrict")) return &o.fuse_restrict;
    if (!strcmp(key, "gs_multicolor")) return &o.gs_multicolor;
    if (!strcmp(key, "ghost_redundant")) return &o.ghost_redundant;
    if (!strcmp(key, "overlap")) return &o.overlap;
    if (!strcmp(key, "overlap_min_rows")) return &o.overlap_min_rows;
    if (!strcmp(key, "bsr_rb")) return &o.bsr_rb;
    if (!strcmp(key, "bsr_u")) return &o.bsr_u;
    if (!strcmp(key, "bsr_stages")) return &o.bsr_stages;
    return nullptr;
}
INT fasp_cuda_set_option(const char* key, double value)
{
    int* s = option_slot(key);
    if (!s) {
        set_last_error(std::string("unknown option ") + (key ? key : "(null)"));
        return ERROR_INPUT_PAR;
    }
    *s = (int)value;
    ctx().opt_epoch++;
    return FASP_SUCCESS;
}
double fasp_cuda_get_option(const char* key)
{
    int* s = option_slot(key);
    return s ? (double)*s : -1.0;
}

// Profile records since the last call, one text line per launch:
//   "<kind> <rows> <nnz> <ms> <algorithmic bytes>"  (kind = CsrMode, 100 = dense GEMV, 2xx = BSR)
// Returns the number of characters written (truncated to cap-1).
long long fasp_cuda_profile_dump(char* buf, long long cap)
{
    Ctx& c = ctx();
    if (c.inited) cudaStreamSynchronize(c.stream);
    std::string out;
    char        line[160];
    for (auto& r : c.prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        snprintf(line, sizeof(line), "%d %d %lld %.6f %.0f\n", r.kind, r.rows, r.nnz, ms, r.bytes);
        out += line;
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    c.prof.clear();
    if (!buf || cap <= 0) return (long long)out.size();
    long long n = (long long)out.size() < cap - 1 ? (long long)out.size() : cap - 1;
    memcpy(buf, out.data(), (size_t)n);
    buf[n] = 0;
    return n;
}

// ------------------------------------------------------------------------------------
// level-1 drop-ins, host pointers
// ------------------------------------------------------------------------------------
INT fasp_cuda_blas_dcsr_mxv(const dCSRmat* A, const REAL* x, REAL* y)
{
    API_TRY
    return host_spmv(A, CSR_MXV, 1.0, x, y, false);
    API_CATCH(code__)
}
// y^T A x (BlaSpmvCSR.c:839): one fused pass, row sums multiplied by y_i and reduced on the device.
// Returns NaN (and sets the error string) on failure.
REAL fasp_cuda_blas_dcsr_vmv(const dCSRmat* A, const REAL* x, const REAL* y)
{
    API_TRY
    ensure_init();
    check_csr(A);
    if (A->row == 0) return 0.0;
    TmpCSR  dA(A);
    DVec    dx(x, A->col), dy(y, A->row), dw((size_t)A->row);
    DVec    out((size_t)1);
    FC_CUDA(cudaMemsetAsync(out.p, 0, sizeof(double), ctx().stream));
    CsrArgs a;
    a.mode         = CSR_MXV;
    a.x            = dx.p;
    a.y            = dw.p;
    a.red.dot_with = dy.p;
    a.red.dot_out  = out.p;
    csr_launch(dA.m, a);
    double v = 0.0;
    out.to_host(&v);
    return v;
    API_CATCH(std::numeric_limits<double>::quiet_NaN())
}
INT fasp_cuda_blas_dcsr_aAxpy(const REAL alpha, const dCSRmat* A, const REAL* x, REAL* y)
{
    API_TRY
    return host_spmv(A, CSR_AXPY, alpha, x, y, false);
    API_CATCH(code__)
}
INT fasp_cuda_blas_dcsr_mxv_agg(const dCSRmat* A, const REAL* x, REAL* y)
{
    API_TRY
    return host_spmv(A, CSR_MXV, 1.0, x, y, true);
    API_CATCH(code__)
}
INT fasp_cuda_blas_dcsr_aAxpy_agg(const REAL alpha, const dCSRmat* A, const REAL* x, REAL* y)
{
    API_TRY
    return host_spmv(A, CSR_AXPY, alpha, x, y, true);
    API_CATCH(code__)
}
void fasp_cuda_blas_mxv_csr(const void* A, const REAL* x, REAL* y)
{
    fasp_cuda_blas_dcsr_mxv(static_cast<const dCSRmat*>(A), x, y);
}
// fasp_solver_matfree_init (SolMatFree.c:201-235) for the formats on the device path
INT fasp_cuda_solver_matfree_init(INT matrix_format, mxv_matfree* mf, void* A)
{
    if (!mf) return ERROR_INPUT_PAR;
    switch (matrix_format) {
        case 1: mf->fct = fasp_cuda_blas_mxv_csr; break;   // MAT_CSR
        case 2: mf->fct = fasp_cuda_blas_mxv_bsr; break;   // MAT_BSR
        default:
            set_last_error("fasp_cuda_solver_matfree_init: only MAT_CSR (1) and MAT_BSR (2) are on the device path");
            return ERROR_DATA_STRUCTURE;
    }
    mf->data = A;
    return FASP_SUCCESS;
}
// Inverse of a dense n x n row-major matrix by the blocked Gauss-Jordan kernel that factors the coarsest AMG
// level (dense.cu). ERROR_AMG_SETUP when a pivot falls below 1e-14 of the largest entry.
INT fasp_cuda_dense_inverse(INT n, const REAL* a, REAL* ainv)
{
    API_TRY
    ensure_init();
    if (n < 0 || (n > 0 && (!a || !ainv))) fail(ERROR_INPUT_PAR, "fasp_cuda_dense_inverse: bad arguments");
    if (n == 0) return FASP_SUCCESS;
    std::vector<double> h(a, a + (size_t)n * n);
    DenseInv D;
    if (!dense_invert_host(D, n, h))
        fail(ERROR_AMG_SETUP, "matrix is numerically singular (pivot below 1e-14 of the largest entry)");
    cudaError_t e = cudaMemcpy(ainv, D.ainv, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost);
    dense_free(D);
    FC_CUDA(e);
    return FASP_SUCCESS;
    API_CATCH(code__)
}

// ---- BLAS-1 drop-ins (BlaArray.c), host pointers: same element-wise operations in the same
// order as the CPU loops (bit-identical); reductions are tree sums (<= 1e-14 relative).
INT fasp_cuda_blas_darray_ax(const INT n, const REAL a, REAL* x)
{
    API_TRY
    ensure_init();
    if (n <= 0 || a == 1.0) return FASP_SUCCESS;   // BlaArray.c:45
    DVec dx(x, (size_t)n);
    vec_ax(a, dx.p, (size_t)n);
    dx.to_host(x);
    return FASP_SUCCESS;
    API_CATCH(code__)
}
INT fasp_cuda_blas_darray_axpy(const INT n, const REAL a, const REAL* x, REAL* y)
{
    API_TRY
    ensure_init();
    if (n <= 0) return FASP_SUCCESS;
    DVec dx(x, (size_t)n), dy(y, (size_t)n);
    vec_axpy(a, dx.p, dy.p, (size_t)n);
    dy.to_host(y);
    return FASP_SUCCESS;
    API_CATCH(code__)
}
INT fasp_cuda_blas_darray_axpby(const INT n, const REAL a, const REAL* x, const REAL b, REAL* y)
{
    API_TRY
    ensure_init();
    if (n <= 0) return FASP_SUCCESS;
    DVec dx(x, (size_t)n), dy(y, (size_t)n);
    vec_axpby(a, dx.p, b, dy.p, (size_t)n);
    dy.to_host(y);
    return FASP_SUCCESS;
    API_CATCH(code__)
}
REAL fasp_cuda_blas_darray_dotprod(const INT n, const REAL* x, const REAL* y)
{
    API_TRY
    ensure_init();
    if (n <= 0) return 0.0;
    DVec dx(x, (size_t)n), dy(y, (size_t)n);
    return vec_dot_host(dx.p, dy.p, (size_t)n);
    API_CATCH(std::numeric_limits<double>::quiet_NaN())
}
REAL fasp_cuda_blas_darray_norm2(const INT n, const REAL* x)
{
    API_TRY
    ensure_init();
    if (n <= 0) return 0.0;
    DVec dx(x, (size_t)n);
    return vec_norm2_host(dx.p, (size_t)n);
    API_CATCH(std::numeric_limits<double>::quiet_NaN())
}
REAL fasp_cuda_blas_darray_norm1(const INT n, const REAL* x)
{
    API_TRY
    ensure_init();
    if (n <= 0) return 0.0;
    DVec   dx(x, (size_t)n);
    double v = 0.0;
    vec_norm1_inf_host(dx.p, (size_t)n, &v, nullptr);
    return v;
    API_CATCH(std::numeric_limits<double>::quiet_NaN())
}
REAL fasp_cuda_blas_darray_norminf(const INT n, const REAL* x)
{
    API_TRY
    ensure_init();
    if (n <= 0) return 0.0;
    DVec   dx(x, (size_t)n);
    double v = 0.0;
    vec_norm1_inf_host(dx.p, (size_t)n, nullptr, &v);
    return v;
    API_CATCH(std::numeric_limits<double>::quiet_NaN())
}

INT fasp_cuda_smoother_dcsr_jacobi(dvector* u, const INT i_1, const INT i_n, const INT s,
                                   dCSRmat* A, dvector* b, INT L, const REAL w)
{
    API_TRY
    ensure_init();
    check_csr(A);
    const int lo = i_1 < i_n ? i_1 : i_n, hi = i_1 < i_n ? i_n : i_1;
    if (lo != 0 || hi != A->row - 1 || (s != 1 && s != -1))
        fail(ERROR_INPUT_PAR, "device Jacobi sweeps all rows (i_1..i_n must span 0..row-1, |s| = 1)");
    TmpCSR dA(A);
    csr_ensure_diag(dA.m);
    DVec   db(b->val, A->row), du(u->val, A->row), dv((size_t)A->row);
    double *in = du.p, *out = dv.p;
    while (L-- > 0) {
        CsrArgs a;
        a.mode  = CSR_JACOBI;
        a.alpha = w;
        a.x     = in;
        a.b     = db.p;
        a.y     = out;
        csr_launch(dA.m, a);
        std::swap(in, out);
    }
    FC_CUDA(cudaMemcpyAsync(u->val, in, sizeof(double) * A->row, cudaMemcpyDeviceToHost,
                            ctx().stream));
    FC_CUDA(cudaStreamSynchronize(ctx().stream));
    return FASP_SUCCESS;
    API_CATCH(code__)
}

INT fasp_cuda_smoother_dcsr_L1diag(dvector* u, const INT i_1, const INT i_n, const INT s,
                                   dCSRmat* A, dvector* b, INT L)
{
    API_TRY
    ensure_init();
    check_csr(A);
    const int lo = i_1 < i_n ? i_1 : i_n, hi = i_1 < i_n ? i_n : i_1;
    if (lo != 0 || hi != A->row - 1 || (s != 1 && s != -1))
        fail(ERROR_INPUT_PAR, "device L1 sweeps all rows (i_1..i_n must span 0..row-1, |s| = 1)");
    TmpCSR dA(A);
    csr_ensure_l1(dA.m);
    DVec   db(b->val, A->row), du(u->val, A->row), dv((size_t)A->row);
    double *in = du.p, *out = dv.p;
    while (L-- > 0) {
        CsrArgs a;
        a.mode = CSR_L1;
        a.x    = in;
        a.b    = db.p;
        a.y    = out;
        csr_launch(dA.m, a);
        std::swap(in, out);
    }
    FC_CUDA(cudaMemcpyAsync(u->val, in, sizeof(double) * A->row, cudaMemcpyDeviceToHost,
                            ctx().stream));
    FC_CUDA(cudaStreamSynchronize(ctx().stream));
    return FASP_SUCCESS;
    API_CATCH(code__)
}

INT fasp_cuda_smoother_dcsr_poly(dCSRmat* Amat, dvector* brhs, dvector* usol, INT n, INT ndeg,
                                 INT L)
{
    API_TRY
    ensure_init();
    check_csr(Amat);
    if (n != Amat->row) fail(ERROR_INPUT_PAR, "poly smoother: n must equal A->row");
    // a one-level "hierarchy" drives the same code path the cycle uses
    AMG_data  mgl[2];
    AMG_param par;
    memset(mgl, 0, sizeof(mgl));
    memset(&par, 0, sizeof(par));
    // level 0 = the matrix, level 1 = a dummy 1x1 coarse level (never visited)
    INT    cia[2] = {0, 1}, cja[1] = {0};
    REAL   cva[1] = {1.0};
    std::vector<INT> pia(Amat->row + 1, 0), ria(2, 0);
    mgl[0].A = *Amat;
    mgl[0].num_levels = 2;
    mgl[0].P.row = Amat->row, mgl[0].P.col = 1, mgl[0].P.nnz = 0, mgl[0].P.IA = pia.data();
    mgl[0].R.row = 1, mgl[0].R.col = Amat->row, mgl[0].R.nnz = 0, mgl[0].R.IA = ria.data();
    mgl[1].A.row = mgl[1].A.col = 1, mgl[1].A.nnz = 1;
    mgl[1].A.IA = cia, mgl[1].A.JA = cja, mgl[1].A.val = cva;
    par.AMG_type = CLASSIC_AMG, par.smoother = SMOOTHER_POLY, par.cycle_type = V_CYCLE;
    par.polynomial_degree = (SHORT)ndeg, par.presmooth_iter = 1, par.postsmooth_iter = 0;
    par.relaxation = 1.0, par.maxit = 1, par.tol = 1e-6;
    Amg* h = amg_upload(mgl, &par);
    try {
        DVec db(brhs->val, n), du(usol->val, n);
        amg_smooth_only(*h, db.p, du.p, L);
        du.to_host(usol->val);
    } catch (...) {
        amg_free(h);
        throw;
    }
    amg_free(h);
    return FASP_SUCCESS;
    API_CATCH(code__)
}

// ------------------------------------------------------------------------------------
// setup-phase pieces on the device (setup.cu)
// ------------------------------------------------------------------------------------
INT fasp_cuda_dcsr_trans(const dCSRmat* A, dCSRmat* AT)
{
    API_TRY
    check_csr(A);
    if (!AT) fail(ERROR_INPUT_PAR, "fasp_cuda_dcsr_trans: null output");
    setup_transpose(A, AT);
    return FASP_SUCCESS;
    API_CATCH(code__)
}
INT fasp_cuda_blas_dcsr_rap(const dCSRmat* R, const dCSRmat* A, const dCSRmat* P, dCSRmat* RAP)
{
    API_TRY
    check_csr(R);
    check_csr(A);
    check_csr(P);
    if (!RAP) fail(ERROR_INPUT_PAR, "fasp_cuda_blas_dcsr_rap: null output");
    setup_rap(R, A, P, RAP);
    return FASP_SUCCESS;
    API_CATCH(code__)
}

// ------------------------------------------------------------------------------------
// resident objects
// ------------------------------------------------------------------------------------
fasp_cuda_csr* fasp_cuda_dcsr_upload(const dCSRmat* A)
{
    API_TRY
    ensure_init();
    check_csr(A);
    fasp_cuda_csr* d = new fasp_cuda_csr_s();
    try {
        csr_upload(d->m, A->row, A->col, A->nnz, A->IA, A->JA, A->val);
    } catch (...) {
        delete d;
        throw;
    }
    return d;
    API_CATCH(nullptr)
}
void fasp_cuda_dcsr_free(fasp_cuda_csr* dA)
{
    if (!dA) return;
    csr_free(dA->m);
    delete dA;
}

REAL* fasp_cuda_dvec_alloc(size_t n)
{
    API_TRY
    ensure_init();
    return dalloc<double>(n);
    API_CATCH(nullptr)
}
void fasp_cuda_dvec_free(REAL* d) { dfree(d); }
INT  fasp_cuda_dvec_h2d(REAL* d, const REAL* h, size_t n)
{
    API_TRY
    ensure_init();
    FC_CUDA(cudaMemcpyAsync(d, h, sizeof(double) * n, cudaMemcpyHostToDevice, ctx().stream));
    FC_CUDA(cudaStreamSynchronize(ctx().stream));
    return FASP_SUCCESS;
    API_CATCH(code__)
}
INT fasp_cuda_dvec_d2h(REAL* h, const REAL* d, size_t n)
{
    API_TRY
    ensure_init();
    FC_CUDA(cudaMemcpyAsync(h, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx().stream));
    FC_CUDA(cudaStreamSynchronize(ctx().stream));
    return FASP_SUCCESS;
    API_CATCH(code__)
}
INT fasp_cuda_sync(void)
{
    API_TRY
    ensure_init();
    FC_CUDA(cudaStreamSynchronize(ctx().stream));
    return FASP_SUCCESS;
    API_CATCH(code__)
}

INT fasp_cuda_dcsr_spmv_dev(const fasp_cuda_csr* dA, int mode, REAL alpha, const REAL* x,
                            const REAL* b, REAL* y)
{
    API_TRY
    if (!dA) fail(ERROR_INPUT_PAR, "null matrix handle");
    if (mode < 0 || mode > 2) fail(ERROR_INPUT_PAR, "spmv mode must be 0, 1 or 2");
    CsrArgs a;
    a.mode  = mode;
    a.alpha = alpha;
    a.x     = x;
    a.b     = b;
    a.y     = y;
    csr_launch(dA->m, a);
    return FASP_SUCCESS;
    API_CATCH(code__)
}

INT fasp_cuda_dcsr_smooth_dev(const fasp_cuda_csr* dA, int kind, REAL w, const REAL* b,
                              const REAL* u_in, REAL* u_out)
{
    API_TRY
    if (!dA) fail(ERROR_INPUT_PAR, "null matrix handle");
    DevCSR& m = const_cast<DevCSR&>(dA->m);
    CsrArgs a;
    if (kind == SMOOTHER_JACOBI) {
        csr_ensure_diag(m);
        a.mode = CSR_JACOBI;
    } else if (kind == SMOOTHER_L1DIAG) {
        csr_ensure_l1(m);
        a.mode = CSR_L1;
    } else
        fail(ERROR_AMG_SMOOTH_TYPE, "smooth_dev: kind must be SMOOTHER_JACOBI or SMOOTHER_L1DIAG");
    a.alpha = w;
    a.x     = u_in;
    a.b     = b;
    a.y     = u_out;
    csr_launch(m, a);
    return FASP_SUCCESS;
    API_CATCH(code__)
}

double fasp_cuda_dcsr_time_kernel(const fasp_cuda_csr* dA, int what, int warm, int reps,
                                  int flush)
{
    API_TRY
    if (!dA) fail(ERROR_INPUT_PAR, "null matrix handle");
    Ctx&    c = ctx();
    DevCSR& m = const_cast<DevCSR&>(dA->m);
    const size_t nx = (size_t)(m.cols > m.rows ? m.cols : m.rows);
    DVec    x(nx), y(nx), b(nx);
    std::vector<double> hx(nx);
    unsigned long long  sd = 88172645463325252ULL;
    for (size_t i = 0; i < nx; ++i) {
        sd ^= sd << 13, sd ^= sd >> 7, sd ^= sd << 17;
        hx[i] = (double)(sd >> 11) / 9007199254740992.0 * 2.0 - 1.0;
    }
    FC_CUDA(cudaMemcpy(x.p, hx.data(), sizeof(double) * nx, cudaMemcpyHostToDevice));
    FC_CUDA(cudaMemcpy(b.p, hx.data(), sizeof(double) * nx, cudaMemcpyHostToDevice));
    FC_CUDA(cudaMemset(y.p, 0, sizeof(double) * nx));
    CsrArgs a;
    a.x = x.p, a.b = b.p, a.y = y.p;
    switch (what) {
        case 0: a.mode = CSR_MXV; break;
        case 1: a.mode = CSR_AXPY, a.alpha = -1.0; break;
        case 2: a.mode = CSR_RESID; break;
        case 10: csr_ensure_diag(m), a.mode = CSR_JACOBI, a.alpha = 0.67; break;
        case 11: csr_ensure_l1(m), a.mode = CSR_L1; break;
        default: fail(ERROR_INPUT_PAR, "time_kernel: unknown kernel id %d", what);
    }
    cudaEvent_t e0, e1;
    FC_CUDA(cudaEventCreate(&e0));
    FC_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < warm; ++i) csr_launch(m, a);
    double total = 0.0;
    if (flush) {
        for (int i = 0; i < reps; ++i) {
            flush_l2();
            FC_CUDA(cudaEventRecord(e0, c.stream));
            csr_launch(m, a);
            FC_CUDA(cudaEventRecord(e1, c.stream));
            FC_CUDA(cudaEventSynchronize(e1));
            float ms = 0.f;
            FC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            total += ms;
        }
    } else {
        FC_CUDA(cudaEventRecord(e0, c.stream));
        for (int i = 0; i < reps; ++i) csr_launch(m, a);
        FC_CUDA(cudaEventRecord(e1, c.stream));
        FC_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        FC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        total = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return total / (reps > 0 ? reps : 1);
    API_CATCH(-1.0)
}

// ------------------------------------------------------------------------------------
// AMG hierarchy
// ------------------------------------------------------------------------------------
fasp_cuda_amg* fasp_cuda_amg_upload(AMG_data* mgl, AMG_param* param)
{
    API_TRY
    Amg*           h = amg_upload(mgl, param);
    fasp_cuda_amg* r = new fasp_cuda_amg_s();
    r->h             = h;
    return r;
    API_CATCH(nullptr)
}
void fasp_cuda_amg_free(fasp_cuda_amg* h)
{
    if (!h) return;
    amg_free(h->h);
    delete h;
}
size_t fasp_cuda_amg_bytes(const fasp_cuda_amg* h) { return h ? h->h->bytes : 0; }
INT    fasp_cuda_amg_levels(const fasp_cuda_amg* h) { return h ? h->h->nl : 0; }

INT fasp_cuda_amg_cycle_dev(fasp_cuda_amg* h, const REAL* r_dev, REAL* z_dev)
{
    API_TRY
    if (!h) fail(ERROR_INPUT_PAR, "null hierarchy handle");
    amg_apply(*h->h, r_dev, z_dev, Reduce(), nullptr);
    return FASP_SUCCESS;
    API_CATCH(code__)
}
INT fasp_cuda_amg_cycle_host(fasp_cuda_amg* h, const REAL* r, REAL* z)
{
    API_TRY
    if (!h) fail(ERROR_INPUT_PAR, "null hierarchy handle");
    const size_t n = h->h->lv[0].n;
    DVec         dr(r, n), dz(n);
    amg_apply(*h->h, dr.p, dz.p, Reduce(), nullptr);
    dz.to_host(z);
    return FASP_SUCCESS;
    API_CATCH(code__)
}

INT fasp_cuda_solver_mgcycle(AMG_data* mgl, AMG_param* param)
{
    API_TRY
    Amg* h = amg_upload(mgl, param);
    try {
        const size_t n = h->lv[0].n;
        DVec         db(mgl[0].b.val, n), dx(mgl[0].x.val, n);
        amg_cycle_inplace(*h, db.p, dx.p, false, Reduce(), nullptr);
        dx.to_host(mgl[0].x.val);
    } catch (...) {
        amg_free(h);
        throw;
    }
    amg_free(h);
    return FASP_SUCCESS;
    API_CATCH(code__)
}

// ------------------------------------------------------------------------------------
// precond plug-in
// ------------------------------------------------------------------------------------
} // extern "C"

namespace {
// Object behind precond.data for device-backed preconditioners. The first member is a
// magic tag so the Krylov entry points can recognise it and stay on the device.
constexpr unsigned long long kPrecMagic = 0xFA5BC0DA00A36B20ULL;
struct DevPrecData {
    unsigned long long magic = kPrecMagic;
    Amg*               h     = nullptr;
    AMG_data*          mgl   = nullptr;   // host hierarchy we own (from precond_setup), or null
    AMG_param          param;
    void*              hostlib = nullptr;
};
DevPrecData* as_dev_prec(precond* pc)
{
    if (!pc || !pc->data) return nullptr;
    DevPrecData* d = static_cast<DevPrecData*>(pc->data);
    return (pc->fct == fasp_cuda_precond_amg && d->magic == kPrecMagic) ? d : nullptr;
}
} // namespace

// ---- host FASP setup routines, resolved at run time -----------------------------------
namespace fc {
HostFasp& host_fasp()
{
    static HostFasp hf;
    if (hf.tried) return hf;
    hf.tried   = true;
    void* self = dlopen(nullptr, RTLD_NOW | RTLD_GLOBAL);
    auto  look = [&](void* lib) {
        hf.amg_data_create = (HostFasp::create_t)dlsym(lib, "fasp_amg_data_create");
        hf.amg_data_free   = (HostFasp::free_t)dlsym(lib, "fasp_amg_data_free");
        hf.setup_rs        = (HostFasp::setup_t)dlsym(lib, "fasp_amg_setup_rs");
        hf.setup_sa        = (HostFasp::setup_t)dlsym(lib, "fasp_amg_setup_sa");
        hf.setup_ua        = (HostFasp::setup_t)dlsym(lib, "fasp_amg_setup_ua");
        hf.dcsr_create     = (HostFasp::csrcreate_t)dlsym(lib, "fasp_dcsr_create");
        hf.dcsr_cp         = (HostFasp::csrcp_t)dlsym(lib, "fasp_dcsr_cp");
        hf.dvec_create     = (HostFasp::dveccreate_t)dlsym(lib, "fasp_dvec_create");
        hf.amg_data_bsr_create = (HostFasp::bcreate_t)dlsym(lib, "fasp_amg_data_bsr_create");
        hf.amg_data_bsr_free   = (HostFasp::bfree_t)dlsym(lib, "fasp_amg_data_bsr_free");
        hf.setup_sa_bsr        = (HostFasp::bsetup_t)dlsym(lib, "fasp_amg_setup_sa_bsr");
        hf.setup_ua_bsr        = (HostFasp::bsetup_t)dlsym(lib, "fasp_amg_setup_ua_bsr");
        hf.dbsr_create         = (HostFasp::bsrcreate_t)dlsym(lib, "fasp_dbsr_create");
        hf.dbsr_cp             = (HostFasp::bsrcp_t)dlsym(lib, "fasp_dbsr_cp");
        return hf.amg_data_create && hf.amg_data_free && hf.setup_rs && hf.dcsr_create &&
               hf.dcsr_cp && hf.dvec_create;
    };
    hf.ok = self && look(self);
    if (!hf.ok) {
        if (const char* path = getenv("FASP_CUDA_HOST_LIBFASP")) {
            void* lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
            hf.ok     = lib && look(lib);
        }
    }
    return hf;
}
void require_host_fasp()
{
    if (!host_fasp().ok)
        fail(ERROR_AMG_SETUP,
             "FASP's host setup routines (fasp_amg_setup_rs, ...) are not visible in this process: "
             "link the application with libfasp or set FASP_CUDA_HOST_LIBFASP=/path/to/libfasp.so");
}
} // namespace fc

extern "C" {

void fasp_cuda_precond_amg(REAL* r, REAL* z, void* data)
{
    DevPrecData* d = static_cast<DevPrecData*>(data);
    if (!d || d->magic != kPrecMagic || !d->h) {
        fprintf(stderr, "### ERROR: fasp_cuda_precond_amg called with foreign data\n");
        return;
    }
    try {
        const size_t n = d->h->lv[0].n;
        DVec         dr(r, n), dz(n);
        amg_apply(*d->h, dr.p, dz.p, Reduce(), nullptr);
        dz.to_host(z);
    } catch (const fc::Error& e) {
        fc::set_last_error(e.msg);
        fprintf(stderr, "### ERROR: fasp_cuda_precond_amg: %s\n", e.msg.c_str());
    }
}

precond* fasp_cuda_precond_from_mgl(AMG_data* mgl, AMG_param* amgparam)
{
    API_TRY
    DevPrecData* d = new DevPrecData();
    try {
        d->param = *amgparam;
        // fasp_precond_amg re-initialises the parameters and copies only some fields
        // (PreCSR.c:425-426, AuxParam.c:816-834): tol falls back to the default 1e-6
        d->param.tol = 1e-6;
        d->h         = amg_upload(mgl, &d->param);
    } catch (...) {
        delete d;
        throw;
    }
    precond* pc = (precond*)malloc(sizeof(precond));
    pc->data    = d;
    pc->fct     = fasp_cuda_precond_amg;
    return pc;
    API_CATCH(nullptr)
}

precond* fasp_cuda_precond_setup(const SHORT precond_type, AMG_param* amgparam,
                                 ILU_param* iluparam, dCSRmat* A)
{
    (void)iluparam;
    API_TRY
    if (precond_type != PREC_AMG)
        fail(ERROR_SOLVER_PRECTYPE, "fasp_cuda_precond_setup: only PREC_AMG (2) is on the device path");
    require_host_fasp();
    HostFasp& hf = host_fasp();
    check_csr(A);
    // PreCSR.c:78-101
    AMG_data* mgl = hf.amg_data_create(amgparam->max_levels);
    mgl[0].A      = hf.dcsr_create(A->row, A->col, A->nnz);
    hf.dcsr_cp(A, &mgl[0].A);
    mgl[0].b = hf.dvec_create(A->col);
    mgl[0].x = hf.dvec_create(A->col);
    INT st   = 0;
    switch (amgparam->AMG_type) {
        case SA_AMG: st = hf.setup_sa ? hf.setup_sa(mgl, amgparam) : ERROR_AMG_SETUP; break;
        case UA_AMG: st = hf.setup_ua ? hf.setup_ua(mgl, amgparam) : ERROR_AMG_SETUP; break;
        default: st = hf.setup_rs(mgl, amgparam);
    }
    if (st < 0) {
        hf.amg_data_free(mgl, amgparam);
        fail(ERROR_AMG_SETUP, "host AMG setup failed with status %d", st);
    }
    precond* pc = fasp_cuda_precond_from_mgl(mgl, amgparam);
    if (!pc) {
        hf.amg_data_free(mgl, amgparam);
        fail(ERROR_AMG_SETUP, "%s", fc::last_error());
    }
    static_cast<DevPrecData*>(pc->data)->mgl = mgl;
    return pc;
    API_CATCH(nullptr)
}

void fasp_cuda_precond_free(precond* pc)
{
    DevPrecData* d = as_dev_prec(pc);
    if (d) {
        amg_free(d->h);
        if (d->mgl && host_fasp().ok) host_fasp().amg_data_free(d->mgl, &d->param);
        delete d;
    }
    free(pc);
}

// ------------------------------------------------------------------------------------
// Krylov, host-pointer drop-ins
// ------------------------------------------------------------------------------------
} // extern "C"

namespace {
// choose the device preconditioner for a `precond*` given by the caller
struct PrecChoice {
    Prec* p = nullptr;
    ~PrecChoice() { delete p; }
};
void choose_prec(PrecChoice& out, precond* pc, size_t n)
{
    if (pc == nullptr) out.p = new IdentityPrec(n);
    else if (DevPrecData* d = as_dev_prec(pc)) out.p = new AmgPrec(d->h);
    else out.p = new HostPrec(pc, n);
}
} // namespace

extern "C" {

INT fasp_cuda_solver_dcsr_pcg(dCSRmat* A, dvector* b, dvector* u, precond* pc, const REAL tol,
                              const REAL abstol, const INT MaxIt, const SHORT StopType,
                              const SHORT PrtLvl)
{
    API_TRY
    ensure_init();
    check_csr(A);
    const size_t n = b->row;
    TmpCSR       dA(A);
    CsrOp        op(&dA.m);
    DVec         db(b->val, n), du(u->val, n);
    PrecChoice   pch;
    choose_prec(pch, pc, n);
    const int ret = pcg_solve(op, db.p, du.p, *pch.p, tol, abstol, MaxIt, StopType, PrtLvl, nullptr);
    du.to_host(u->val);
    return ret;
    API_CATCH(code__)
}

INT fasp_cuda_solver_dcsr_pvgmres(dCSRmat* A, dvector* b, dvector* x, precond* pc,
                                  const REAL tol, const REAL abstol, const INT MaxIt,
                                  const SHORT restart, const SHORT StopType, const SHORT PrtLvl)
{
    API_TRY
    ensure_init();
    check_csr(A);
    const size_t n = b->row;
    TmpCSR       dA(A);
    CsrOp        op(&dA.m);
    DVec         db(b->val, n), dx(x->val, n);
    PrecChoice   pch;
    choose_prec(pch, pc, n);
    const int ret = gmres_solve(op, db.p, dx.p, *pch.p, tol, abstol, MaxIt, restart, StopType,
                                PrtLvl, GM_VARIABLE, nullptr);
    dx.to_host(x->val);
    return ret;
    API_CATCH(code__)
}

INT fasp_cuda_solver_dcsr_pvfgmres(dCSRmat* A, dvector* b, dvector* x, precond* pc,
                                   const REAL tol, const REAL abstol, const INT MaxIt,
                                   const SHORT restart, const SHORT StopType, const SHORT PrtLvl)
{
    API_TRY
    ensure_init();
    check_csr(A);
    const size_t n = b->row;
    TmpCSR       dA(A);
    CsrOp        op(&dA.m);
    DVec         db(b->val, n), dx(x->val, n);
    PrecChoice   pch;
    choose_prec(pch, pc, n);
    const int ret = gmres_solve(op, db.p, dx.p, *pch.p, tol, abstol, MaxIt, restart, StopType,
                                PrtLvl, GM_FLEXIBLE, nullptr);
    dx.to_host(x->val);
    return ret;
    API_CATCH(code__)
}

INT fasp_cuda_solver_dcsr_pgmres(dCSRmat* A, dvector* b, dvector* x, precond* pc, const REAL tol,
                                 const REAL abstol, const INT MaxIt, const SHORT restart,
                                 const SHORT StopType, const SHORT PrtLvl)
{
    API_TRY
    ensure_init();
    check_csr(A);
    const size_t n = b->row;
    TmpCSR       dA(A);
    CsrOp        op(&dA.m);
    DVec         db(b->val, n), dx(x->val, n);
    PrecChoice   pch;
    choose_prec(pch, pc, n);
    const int ret = gmres_solve(op, db.p, dx.p, *pch.p, tol, abstol, MaxIt, restart, StopType,
                                PrtLvl, GM_FIXED, nullptr);
    dx.to_host(x->val);
    return ret;
    API_CATCH(code__)
}

INT fasp_cuda_solver_dcsr_itsolver(dCSRmat* A, dvector* b, dvector* x, precond* pc,
                                   ITS_param* itparam)
{
    // SolCSR.c:56-140
    const SHORT prt = itparam->print_level, stop = itparam->stop_type;
    const INT   maxit = itparam->maxit, restart = itparam->restart;
    const REAL  tol = itparam->tol, abstol = itparam->abstol;
    switch (itparam->itsolver_type) {
        case SOLVER_CG: return fasp_cuda_solver_dcsr_pcg(A, b, x, pc, tol, abstol, maxit, stop, prt);
        case SOLVER_GMRES:
            return fasp_cuda_solver_dcsr_pgmres(A, b, x, pc, tol, abstol, maxit, (SHORT)restart, stop, prt);
        case SOLVER_VGMRES:
            return fasp_cuda_solver_dcsr_pvgmres(A, b, x, pc, tol, abstol, maxit, (SHORT)restart, stop, prt);
        case SOLVER_VFGMRES:
            return fasp_cuda_solver_dcsr_pvfgmres(A, b, x, pc, tol, abstol, maxit, (SHORT)restart, stop, prt);
        default:
            set_last_error("itsolver_type not on the device path (supported: CG 1, GMRES 4, VGMRES 5, VFGMRES 6)");
            return ERROR_SOLVER_TYPE;
    }
}

// ------------------------------------------------------------------------------------
// drivers
// ------------------------------------------------------------------------------------
INT fasp_cuda_host_pin(void* p, size_t bytes)
{
    API_TRY
    ensure_init();
    if (!p || bytes == 0) fail(ERROR_INPUT_PAR, "fasp_cuda_host_pin: null buffer");
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return FASP_SUCCESS;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(ERROR_ALLOC_MEM, "cudaHostRegister(%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
    return FASP_SUCCESS;
    API_CATCH(code__)
}
INT fasp_cuda_host_unpin(void* p)
{
    API_TRY
    if (!p) return FASP_SUCCESS;
    FC_CUDA(cudaStreamSynchronize(ctx().stream));
    if (cudaHostUnregister(p) != cudaSuccess) {
        cudaGetLastError();
        fail(ERROR_INPUT_PAR, "fasp_cuda_host_unpin: buffer was not pinned");
    }
    return FASP_SUCCESS;
    API_CATCH(code__)
}

fasp_cuda_solver* fasp_cuda_krylov_amg_create(AMG_data* mgl, AMG_param* amgparam)
{
    API_TRY
    return solver_create_csr(mgl, amgparam);
    API_CATCH(nullptr)
}
void fasp_cuda_krylov_amg_destroy(fasp_cuda_solver* s) { solver_destroy(s); }

INT fasp_cuda_krylov_amg_solve_dev(fasp_cuda_solver* s, const REAL* b_dev, REAL* x_dev,
                                   ITS_param* itparam)
{
    API_TRY
    return solver_solve_dev(s, b_dev, x_dev, itparam);
    API_CATCH(code__)
}
INT fasp_cuda_krylov_amg_solve(fasp_cuda_solver* s, dvector* b, dvector* x, ITS_param* itparam)
{
    API_TRY
    return solver_solve_host(s, b->val, x->val, itparam);
    API_CATCH(code__)
}
double fasp_cuda_solver_stat(const fasp_cuda_solver* s, int what)
{
    return solver_stat(s, what);
}
INT fasp_cuda_solver_history(const fasp_cuda_solver* s, REAL* relres, INT max_entries)
{
    return solver_history(s, relres, max_entries);
}

INT fasp_cuda_solver_dcsr_krylov_amg(dCSRmat* A, dvector* b, dvector* x, ITS_param* itparam,
                                     AMG_param* amgparam)
{
    API_TRY
    ensure_init();
    check_csr(A);
    require_host_fasp();
    HostFasp& hf = host_fasp();
    // SolCSR.c:500-521: hierarchy by the host application's own FASP
    AMG_data* mgl = hf.amg_data_create(amgparam->max_levels);
    mgl[0].A      = hf.dcsr_create(A->row, A->col, A->nnz);
    hf.dcsr_cp(A, &mgl[0].A);
    mgl[0].b = hf.dvec_create(A->col);
    mgl[0].x = hf.dvec_create(A->col);
    INT st   = 0;
    switch (amgparam->AMG_type) {
        case SA_AMG: st = hf.setup_sa ? hf.setup_sa(mgl, amgparam) : ERROR_AMG_SETUP; break;
        case UA_AMG: st = hf.setup_ua ? hf.setup_ua(mgl, amgparam) : ERROR_AMG_SETUP; break;
        default: st = hf.setup_rs(mgl, amgparam);
    }
    if (st < 0) {
        hf.amg_data_free(mgl, amgparam);
        fail(ERROR_AMG_SETUP, "host AMG setup failed with status %d", st);
    }
    fasp_cuda_solver* s = nullptr;
    INT               ret;
    try {
        s   = solver_create_csr(mgl, amgparam);
        ret = solver_solve_host(s, b->val, x->val, itparam);
    } catch (...) {
        solver_destroy(s);
        hf.amg_data_free(mgl, amgparam);
        throw;
    }
    solver_destroy(s);
    hf.amg_data_free(mgl, amgparam);
    return ret;
    API_CATCH(code__)
}

INT fasp_cuda_amg_solve(AMG_data* mgl, AMG_param* param)
{
    API_TRY
    return solver_amg_solve(mgl, param);
    API_CATCH(code__)
}

// AMG as a solver (SolAMG.c:49-150): host setup, then cycles until ||b - A x|| / ||b|| < param->tol
INT fasp_cuda_solver_amg(dCSRmat* A, dvector* b, dvector* x, AMG_param* param)
{
    API_TRY
    ensure_init();
    check_csr(A);
    if (!b || !x || !param) fail(ERROR_INPUT_PAR, "fasp_cuda_solver_amg: null argument");
    if (param->cycle_type == AMLI_CYCLE || param->cycle_type == NL_AMLI_CYCLE)
        fail(ERROR_INPUT_PAR, "AMLI cycles are not on the device path");
    require_host_fasp();
    HostFasp& hf  = host_fasp();
    AMG_data* mgl = hf.amg_data_create(param->max_levels);
    mgl[0].A      = hf.dcsr_create(A->row, A->col, A->nnz);
    hf.dcsr_cp(A, &mgl[0].A);
    mgl[0].b = hf.dvec_create(A->col);
    mgl[0].x = hf.dvec_create(A->col);
    memcpy(mgl[0].b.val, b->val, sizeof(REAL) * (size_t)A->col);
    memcpy(mgl[0].x.val, x->val, sizeof(REAL) * (size_t)A->col);
    INT st = 0;
    switch (param->AMG_type) {
        case SA_AMG: st = hf.setup_sa ? hf.setup_sa(mgl, param) : ERROR_AMG_SETUP; break;
        case UA_AMG: st = hf.setup_ua ? hf.setup_ua(mgl, param) : ERROR_AMG_SETUP; break;
        default: st = hf.setup_rs(mgl, param);
    }
    if (st < 0) {
        hf.amg_data_free(mgl, param);
        fail(ERROR_AMG_SETUP, "host AMG setup failed with status %d", st);
    }
    INT iter;
    try {
        iter = solver_amg_solve(mgl, param);
        memcpy(x->val, mgl[0].x.val, sizeof(REAL) * (size_t)A->col);
    } catch (...) {
        hf.amg_data_free(mgl, param);
        throw;
    }
    hf.amg_data_free(mgl, param);
    return iter;
    API_CATCH(code__)
}

} // extern "C"

// ------------------------------------------------------------------------------------
// BSR twins
// ------------------------------------------------------------------------------------
struct fasp_cuda_bsr_s {
    DevBSR m;
};
struct fasp_cuda_bamg_s {
    BAmg* h;
};

namespace {
void check_bsr(const dBSRmat* A)
{
    if (!A || A->ROW < 0 || A->COL < 0 || A->NNZ < 0 || !A->IA || (A->NNZ > 0 && (!A->JA || !A->val)))
        fail(ERROR_DATA_STRUCTURE, "invalid dBSRmat");
    if (A->storage_manner != 0) fail(ERROR_INPUT_PAR, "dBSRmat: only row-major blocks (storage_manner 0)");
}
struct TmpBSR {
    DevBSR m;
    explicit TmpBSR(const dBSRmat* A)
    {
        check_bsr(A);
        bsr_upload(m, A->ROW, A->COL, A->NNZ, A->nb, A->IA, A->JA, A->val, true);
    }
    ~TmpBSR() { bsr_free(m); }
};
int host_bsr_spmv(const dBSRmat* A, int mode, double alpha, const double* x, double* y)
{
    ensure_init();
    TmpBSR dA(A);
    const size_t nx = (size_t)A->COL * A->nb, ny = (size_t)A->ROW * A->nb;
    DVec   dx(x, nx), dy(ny);
    if (mode == BSR_AXPY)
        FC_CUDA(cudaMemcpyAsync(dy.p, y, sizeof(double) * ny, cudaMemcpyHostToDevice, ctx().stream));
    BsrArgs a;
    a.mode  = mode;
    a.alpha = alpha;
    a.x     = dx.p;
    a.y     = dy.p;
    bsr_launch(dA.m, a);
    dy.to_host(y);
    return FASP_SUCCESS;
}
int bsr_krylov_host(dBSRmat* A, dvector* b, dvector* x, precond* pc, double tol, double abstol, int MaxIt,
                    int restart, int StopType, int PrtLvl, int which)
{
    ensure_init();
    TmpBSR       dA(A);
    BsrOp        op(&dA.m);
    const size_t n = b->row;
    DVec         db(b->val, n), dx(x->val, n);
    PrecChoice   pch;
    if (pc == nullptr) pch.p = new IdentityPrec(n);
    else pch.p = new HostPrec(pc, n);
    int ret;
    if (which == 0) ret = pcg_solve(op, db.p, dx.p, *pch.p, tol, abstol, MaxIt, StopType, PrtLvl, nullptr);
    else ret = gmres_solve(op, db.p, dx.p, *pch.p, tol, abstol, MaxIt, restart, StopType, PrtLvl,
                           which == 3 ? GM_FLEXIBLE : (which == 2 ? GM_VARIABLE : GM_FIXED), nullptr);
    dx.to_host(x->val);
    return ret;
}
} // namespace

extern "C" {

INT fasp_cuda_blas_dbsr_mxv(const dBSRmat* A, const REAL* x, REAL* y)
{
    API_TRY
    return host_bsr_spmv(A, BSR_MXV, 1.0, x, y);
    API_CATCH(code__)
}
INT fasp_cuda_blas_dbsr_aAxpy(const REAL alpha, const dBSRmat* A, const REAL* x, REAL* y)
{
    API_TRY
    return host_bsr_spmv(A, BSR_AXPY, alpha, x, y);
    API_CATCH(code__)
}
void fasp_cuda_blas_mxv_bsr(const void* A, const REAL* x, REAL* y)
{
    fasp_cuda_blas_dbsr_mxv(static_cast<const dBSRmat*>(A), x, y);
}

INT fasp_cuda_smoother_dbsr_jacobi1(dBSRmat* A, dvector* b, dvector* u, REAL* diaginv)
{
    API_TRY
    ensure_init();
    TmpBSR       dA(A);
    const size_t n = (size_t)A->ROW * A->nb, nd = (size_t)A->ROW * A->nb * A->nb;
    DVec         db(b->val, n), du(u->val, n), dv(n), dd(diaginv, nd);
    BsrArgs      a;
    a.mode = BSR_JACOBI, a.x = du.p, a.b = db.p, a.y = dv.p, a.diaginv = dd.p;
    bsr_launch(dA.m, a);
    dv.to_host(u->val);
    return FASP_SUCCESS;
    API_CATCH(code__)
}

fasp_cuda_bsr* fasp_cuda_dbsr_upload(const dBSRmat* A)
{
    API_TRY
    ensure_init();
    check_bsr(A);
    fasp_cuda_bsr* d = new fasp_cuda_bsr_s();
    try {
        bsr_upload(d->m, A->ROW, A->COL, A->NNZ, A->nb, A->IA, A->JA, A->val);
    } catch (...) {
        delete d;
        throw;
    }
    return d;
    API_CATCH(nullptr)
}
void fasp_cuda_dbsr_free(fasp_cuda_bsr* dA)
{
    if (!dA) return;
    bsr_free(dA->m);
    delete dA;
}
INT fasp_cuda_dbsr_spmv_dev(const fasp_cuda_bsr* dA, int mode, REAL alpha, const REAL* x, const REAL* b, REAL* y)
{
    API_TRY
    if (!dA) fail(ERROR_INPUT_PAR, "null matrix handle");
    if (mode < 0 || mode > 2) fail(ERROR_INPUT_PAR, "spmv mode must be 0, 1 or 2");
    BsrArgs a;
    a.mode = mode, a.alpha = alpha, a.x = x, a.b = b, a.y = y;
    bsr_launch(dA->m, a);
    return FASP_SUCCESS;
    API_CATCH(code__)
}
double fasp_cuda_dbsr_time_kernel(const fasp_cuda_bsr* dA, int what, int warm, int reps, int flush)
{
    API_TRY
    if (!dA) fail(ERROR_INPUT_PAR, "null matrix handle");
    Ctx&          c  = ctx();
    const DevBSR& m  = dA->m;
    const size_t  nx = (size_t)(m.COL > m.ROW ? m.COL : m.ROW) * m.nb;
    DVec          x(nx), y(nx), b(nx), dinv((size_t)m.ROW * m.nb * m.nb);
    std::vector<double> hx(nx);
    unsigned long long  sd = 88172645463325252ULL;
    for (size_t i = 0; i < nx; ++i) {
        sd ^= sd << 13, sd ^= sd >> 7, sd ^= sd << 17;
        hx[i] = (double)(sd >> 11) / 9007199254740992.0 * 2.0 - 1.0;
    }
    FC_CUDA(cudaMemcpy(x.p, hx.data(), sizeof(double) * nx, cudaMemcpyHostToDevice));
    FC_CUDA(cudaMemcpy(b.p, hx.data(), sizeof(double) * nx, cudaMemcpyHostToDevice));
    FC_CUDA(cudaMemset(y.p, 0, sizeof(double) * nx));
    FC_CUDA(cudaMemset(dinv.p, 0, sizeof(double) * dinv.n));
    BsrArgs a;
    a.x = x.p, a.b = b.p, a.y = y.p, a.diaginv = dinv.p;
    switch (what) {
        case 0: a.mode = BSR_MXV; break;
        case 1: a.mode = BSR_AXPY, a.alpha = -1.0; break;
        case 2: a.mode = BSR_RESID; break;
        case 10: a.mode = BSR_JACOBI; break;
        default: fail(ERROR_INPUT_PAR, "time_kernel: unknown kernel id %d", what);
    }
    cudaEvent_t e0, e1;
    FC_CUDA(cudaEventCreate(&e0));
    FC_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < warm; ++i) bsr_launch(m, a);
    double total = 0.0;
    for (int i = 0; i < reps; ++i) {
        if (flush) flush_l2();
        FC_CUDA(cudaEventRecord(e0, c.stream));
        bsr_launch(m, a);
        FC_CUDA(cudaEventRecord(e1, c.stream));
        FC_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        FC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return total / (reps > 0 ? reps : 1);
    API_CATCH(-1.0)
}

fasp_cuda_bamg* fasp_cuda_bamg_upload(AMG_data_bsr* mgl, AMG_param* param)
{
    API_TRY
    BAmg*           h = bamg_upload(mgl, param);
    fasp_cuda_bamg* r = new fasp_cuda_bamg_s();
    r->h              = h;
    return r;
    API_CATCH(nullptr)
}
void fasp_cuda_bamg_free(fasp_cuda_bamg* h)
{
    if (!h) return;
    bamg_free(h->h);
    delete h;
}

INT fasp_cuda_solver_mgcycle_bsr(AMG_data_bsr* mgl, AMG_param* param)
{
    API_TRY
    BAmg* h = bamg_upload(mgl, param);
    try {
        const size_t n = h->lv[0].n;
        DVec         db(mgl[0].b.val, n), dx(mgl[0].x.val, n);
        bamg_cycle_inplace(*h, db.p, dx.p, false, Reduce(), nullptr);
        dx.to_host(mgl[0].x.val);
    } catch (...) {
        bamg_free(h);
        throw;
    }
    bamg_free(h);
    return FASP_SUCCESS;
    API_CATCH(code__)
}

INT fasp_cuda_solver_dbsr_pcg(dBSRmat* A, dvector* b, dvector* u, precond* pc, const REAL tol, const REAL abstol,
                              const INT MaxIt, const SHORT StopType, const SHORT PrtLvl)
{
    API_TRY
    return bsr_krylov_host(A, b, u, pc, tol, abstol, MaxIt, 0, StopType, PrtLvl, 0);
    API_CATCH(code__)
}
INT fasp_cuda_solver_dbsr_pgmres(dBSRmat* A, dvector* b, dvector* x, precond* pc, const REAL tol, const REAL abstol,
                                 const INT MaxIt, const SHORT restart, const SHORT StopType, const SHORT PrtLvl)
{
    API_TRY
    return bsr_krylov_host(A, b, x, pc, tol, abstol, MaxIt, restart, StopType, PrtLvl, 1);
    API_CATCH(code__)
}
INT fasp_cuda_solver_dbsr_pvgmres(dBSRmat* A, dvector* b, dvector* x, precond* pc, const REAL tol, const REAL abstol,
                                  const INT MaxIt, const SHORT restart, const SHORT StopType, const SHORT PrtLvl)
{
    API_TRY
    return bsr_krylov_host(A, b, x, pc, tol, abstol, MaxIt, restart, StopType, PrtLvl, 2);
    API_CATCH(code__)
}
INT fasp_cuda_solver_dbsr_pvfgmres(dBSRmat* A, dvector* b, dvector* x, precond* pc, const REAL tol, const REAL abstol,
                                   const INT MaxIt, const SHORT restart, const SHORT StopType, const SHORT PrtLvl)
{
    API_TRY
    return bsr_krylov_host(A, b, x, pc, tol, abstol, MaxIt, restart, StopType, PrtLvl, 3);
    API_CATCH(code__)
}
INT fasp_cuda_solver_dbsr_itsolver(dBSRmat* A, dvector* b, dvector* x, precond* pc, ITS_param* itparam)
{
    // SolBSR.c:55-140
    const SHORT prt = itparam->print_level, stop = itparam->stop_type;
    const INT   maxit = itparam->maxit, restart = itparam->restart;
    const REAL  tol = itparam->tol, abstol = itparam->abstol;
    switch (itparam->itsolver_type) {
        case SOLVER_CG: return fasp_cuda_solver_dbsr_pcg(A, b, x, pc, tol, abstol, maxit, stop, prt);
        case SOLVER_GMRES: return fasp_cuda_solver_dbsr_pgmres(A, b, x, pc, tol, abstol, maxit, (SHORT)restart, stop, prt);
        case SOLVER_VGMRES: return fasp_cuda_solver_dbsr_pvgmres(A, b, x, pc, tol, abstol, maxit, (SHORT)restart, stop, prt);
        case SOLVER_VFGMRES: return fasp_cuda_solver_dbsr_pvfgmres(A, b, x, pc, tol, abstol, maxit, (SHORT)restart, stop, prt);
        default:
            set_last_error("itsolver_type not on the device path (supported: CG 1, GMRES 4, VGMRES 5, VFGMRES 6)");
            return ERROR_SOLVER_TYPE;
    }
}

fasp_cuda_solver* fasp_cuda_krylov_bamg_create(AMG_data_bsr* mgl, AMG_param* amgparam)
{
    API_TRY
    return solver_create_bsr(mgl, amgparam);
    API_CATCH(nullptr)
}

INT fasp_cuda_solver_dbsr_krylov_amg(dBSRmat* A, dvector* b, dvector* x, ITS_param* itparam, AMG_param* amgparam)
{
    API_TRY
    ensure_init();
    check_bsr(A);
    require_host_fasp();
    HostFasp& hf = host_fasp();
    if (!hf.amg_data_bsr_create || !hf.amg_data_bsr_free || !hf.setup_ua_bsr || !hf.dbsr_create || !hf.dbsr_cp)
        fail(ERROR_AMG_SETUP, "the host FASP library lacks the BSR AMG setup routines");
    // SolBSR.c:371-390
    AMG_data_bsr* mgl = hf.amg_data_bsr_create(amgparam->max_levels);
    mgl[0].A          = hf.dbsr_create(A->ROW, A->COL, A->NNZ, A->nb, A->storage_manner);
    mgl[0].b          = hf.dvec_create(mgl[0].A.ROW * mgl[0].A.nb);
    mgl[0].x          = hf.dvec_create(mgl[0].A.COL * mgl[0].A.nb);
    hf.dbsr_cp(A, &mgl[0].A);
    INT st = (amgparam->AMG_type == SA_AMG && hf.setup_sa_bsr) ? hf.setup_sa_bsr(mgl, amgparam)
                                                               : hf.setup_ua_bsr(mgl, amgparam);
    if (st < 0) {
        hf.amg_data_bsr_free(mgl, amgparam);
        fail(ERROR_AMG_SETUP, "host BSR AMG setup failed with status %d", st);
    }
    fasp_cuda_solver* s = nullptr;
    INT               ret;
    try {
        s   = solver_create_bsr(mgl, amgparam);
        ret = solver_solve_host(s, b->val, x->val, itparam);
    } catch (...) {
        solver_destroy(s);
        hf.amg_data_bsr_free(mgl, amgparam);
        throw;
    }
    solver_destroy(s);
    hf.amg_data_bsr_free(mgl, amgparam);
    return ret;
    API_CATCH(code__)
}

} // extern "C"

extern "C" {

INT fasp_cuda_smoother_dcsr_gs_multicolor(dvector* u, dCSRmat* A, dvector* b, INT L, INT order)
{
    API_TRY
    ensure_init();
    check_csr(A);
    std::vector<int> ic, icmap;
    gs_multicolor_host(A->row, A->IA, A->JA, ic, icmap);
    TmpCSR dA(A);
    DVec   db(b->val, A->row), du(u->val, A->row);
    int*   rows = dalloc<int>(icmap.size() ? icmap.size() : 1);
    try {
        FC_CUDA(cudaMemcpyAsync(rows, icmap.data(), sizeof(int) * icmap.size(), cudaMemcpyHostToDevice,
                                ctx().stream));
        gs_multicolor_sweeps(dA.m, rows, ic, db.p, du.p, L, order == -1 ? -1 : 1, nullptr);
        du.to_host(u->val);
    } catch (...) {
        dfree(rows);
        throw;
    }
    dfree(rows);
    return FASP_SUCCESS;
    API_CATCH(code__)
}

} // extern "C"

#include "comm.cuh"
#include "p2p.cuh"
extern "C" {
INT fasp_cuda_comm_unique_id(void* id128)
{
    API_TRY
    comm_unique_id(id128);
    return FASP_SUCCESS;
    API_CATCH(code__)
}
INT fasp_cuda_comm_init(const void* id128, int rank, int nranks)
{
    API_TRY
    comm_init(id128, rank, nranks);
    return FASP_SUCCESS;
    API_CATCH(code__)
}
INT fasp_cuda_comm_finalize(void)
{
    API_TRY
    comm_finalize();
    return FASP_SUCCESS;
    API_CATCH(code__)
}
int fasp_cuda_comm_rank(void) { return comm_rank(); }
int fasp_cuda_comm_size(void) { return comm_size(); }
int fasp_cuda_comm_peer_memory(void) { return p2p_active() ? 1 : 0; }
} // extern "C"

extern "C" INT fasp_cuda_multicolor_host(INT n, const INT* IA, const INT* JA, INT* IC, INT* ICMAP)
{
    API_TRY
    std::vector<int> ic, icmap;
    gs_multicolor_host(n, IA, JA, ic, icmap);
    for (size_t i = 0; i < ic.size(); ++i) IC[i] = ic[i];
    for (size_t i = 0; i < icmap.size(); ++i) ICMAP[i] = icmap[i];
    return (INT)ic.size() - 1;
    API_CATCH(code__)
}

// ------------------------------------------------------------------------------------
// multi-GPU: row-partitioned hierarchy
// ------------------------------------------------------------------------------------
#include "dist.cuh"
extern "C" {

fasp_cuda_solver* fasp_cuda_dist_krylov_amg_create(AMG_data* mgl, AMG_param* amgparam, INT agg_rows)
{
    API_TRY
    return solver_create_dist(mgl, amgparam, agg_rows);
    API_CATCH(nullptr)
}

fasp_cuda_solver* fasp_cuda_dist_krylov_amg_create_slabs(INT nlev, const fasp_cuda_slab_level* levels,
                                                         const INT* tail_row_off, AMG_data* tail,
                                                         AMG_param* amgparam)
{
    API_TRY
    return solver_create_dist_slabs(nlev, levels, tail_row_off, tail, amgparam);
    API_CATCH(nullptr)
}

INT fasp_cuda_dist_row_range(const fasp_cuda_solver* s, INT* row_begin, INT* row_end)
{
    if (!s || !s->amg) return ERROR_INPUT_PAR;
    const Amg& h = *s->amg;
    if (h.dist && !h.off0.empty()) {
        *row_begin = h.off0[comm_rank()];
        *row_end   = h.off0[comm_rank() + 1];
    } else {
        *row_begin = 0;
        *row_end   = h.lv[0].n;
    }
    return FASP_SUCCESS;
}

INT fasp_cuda_dist_extract_host(const dCSRmat* A, INT nranks, INT rank, INT* ia, INT* ja, INT* ghosts,
                                INT ghost_cap, INT* nghost, INT* send_idx, INT send_cap, INT* send_counts)
{
    API_TRY
    check_csr(A);
    if (A->row != A->col) fail(ERROR_MAT_SIZE, "dist_extract_host: square operators only");
    if (rank < 0 || rank >= nranks) fail(ERROR_INPUT_PAR, "rank out of range");
    const std::vector<int> off = dist_partition(A->row, nranks);
    LocalCSR               loc;
    dist_extract(*A, off[rank], off[rank + 1], off, rank, false, loc);
    for (size_t i = 0; i < loc.ia.size(); ++i) ia[i] = loc.ia[i];
    for (size_t i = 0; i < loc.ja.size(); ++i) ja[i] = loc.ja[i];
    *nghost = (INT)loc.ghosts.size();
    if ((INT)loc.ghosts.size() > ghost_cap) fail(ERROR_INPUT_PAR, "ghost buffer too small");
    for (size_t i = 0; i < loc.ghosts.size(); ++i) ghosts[i] = loc.ghosts[i];
    std::vector<std::vector<int>> send;
    dist_send_lists(*A, off, off, rank, send);
    INT pos = 0;
    for (int q = 0; q < nranks; ++q) {
        send_counts[q] = (INT)send[q].size();
        for (int c : send[q]) {
            if (pos >= send_cap) fail(ERROR_INPUT_PAR, "send buffer too small");
            send_idx[pos++] = c;
        }
    }
    return (INT)loc.ja.size();
    API_CATCH(code__)
}

} // extern "C"
