// common.cuh — shared declarations of libfasp_cuda (B200 / sm_100a).
//
// Conventions
//  * everything lives in namespace fc; only capi.cu defines extern "C" symbols
//  * internal code throws fc::Error; the C-ABI wrappers translate to FASP status codes
//  * all kernels are launched on ctx().stream through FC_LAUNCH so that the library can
//    report how many of its own kernels ran (bench.py "gpu_launches")
//  * every kernel that is part of a Krylov iteration takes `const int* done` and returns
//    immediately when *done != 0: the host enqueues iterations ahead of the convergence
//    test, so there is no host round-trip per iteration
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>

#include "fasp_cuda.h"

namespace fc {

// ------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------
struct Error {
    int         code;
    std::string msg;
};
void        set_last_error(const std::string& s);
const char* last_error();

[[noreturn]] void fail(int code, const char* fmt, ...);

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line)
{
    if (e == cudaSuccess) return;
    int code = (e == cudaErrorMemoryAllocation) ? ERROR_ALLOC_MEM : ERROR_SOLVER_MISC;
    cudaGetLastError(); // clear sticky-less error state
    fail(code, "CUDA error %s (%s) at %s:%d: %s", cudaGetErrorName(e), what, file, line,
         cudaGetErrorString(e));
}
#define FC_CUDA(call) ::fc::cuda_check((call), #call, __FILE__, __LINE__)

// ------------------------------------------------------------------------------------
// process-wide context (one process per GPU)
// ------------------------------------------------------------------------------------
struct Options {
    int    strict           = 0;
    int    coarse_dense     = 1;
    int    coarse_dense_max = 8192;
    int    graph            = 1;
    int    zero_guess       = 1;
    int    lookahead        = 2;   // Krylov iterations enqueued ahead of the status read
    int    vec_min_avg      = 48;  // rows averaging >= this many nonzeros use the vector kernel
    int    pipe             = 1;   // stream kernel: persistent TMA-pipelined variant
    int    pipe_ctas        = 8;   // its CTAs per SM (upper bound)
    int    pipe_stages      = 2;   // its shared-memory stages per CTA
    int    host_register    = 1;   // 2: page-lock unknown caller buffers in place and remember them (see solver.cu)
    int    sort_rows        = 1;   // sort the entries of vector-kernel rows by column at upload
    int    pipe_cap_mult    = 16;   // row-block capacity <= this many nonzeros per thread
    int    pipe_tpb         = 128; // its threads per CTA = max rows per row block (64/128/256)
    int    wide_threads     = 400000; // long rows: double the lanes per row (up to 256 = csr_wide_kernel) while
                                      // rows x lanes stays below this and a lane keeps >= 8 entries
    int    two_phase_mask   = 0x6;  // pipelined kernel: bit MODE set = gather rounds in two enforced phases (spmv.cu).
                                   // Default: y += a A x (transfer operators) and r = b - A x, where it measured faster
                                   // (profiles/r02_two_phase_sweep.txt); the L1 / Jacobi sweeps and y = A x lose with it
    int    gather16_min_avg = 6;   // pipelined kernel: rows averaging >= this use the 16-deep gather variant (0 = never)
    int    rowwise_max      = 64;  // blocks averaging <= this many nonzeros per row: one thread per row
    int    vec_lpr          = 0;   // > 0: force this many lanes per row in the vector kernel
    int    fuse_restrict    = 1;   // R's epilogue also writes the next level's zero-guess sweep x = s b / d
                                   // (1: multi-GPU hierarchies only, 2: always, 0: never)
    int    vec_u            = 0;   // vector kernel: loads in flight per lane (4 / 8; 0 = 8 when a lane owns >= 12 entries of an
                                   // average row: levels 4-7 of the 256^3 hierarchy gain 5-13 %, level 3 (8 per lane) loses 4 %,
                                   // profiles/r02_vector_kernel_depth.txt)
    int    gs_multicolor    = 0;   // accept SMOOTHER_GS in the cycle as multicolour GS (OpenMP-FASP semantics)
    int    profile          = 0;   // record CUDA events around every matrix kernel (no graphs)
    int    ghost_redundant  = 1;   // multi-GPU: P and R also compute the ghost rows of the level they write to, so the
                                   // post-smoother and the residual start without their own ghost exchange (2 per level)
    int    overlap          = 1;   // multi-GPU: rows without ghost columns run on a second stream while the ghosts travel
    int    overlap_min_rows = 8192; // ... when the operator's interior has at least this many rows: on smaller slabs the two extra
                                    // graph nodes cost more than the overlap hides (4 GPUs, 7-pt 256^3: 33.2 ms with 256, 27.1 with
                                    // 8192, 31.2 with 100000, 32.2 without overlap; profiles/r02_dist_sweep_N4.txt)
    int    bsr_rb           = 64;  // BSR pipelined kernel: block rows per CTA (32 / 64), nb <= 4. Measured on the 3x3-block
                                   // 7-point matrix (160^3): 32/4 0.69 of the copy peak, 32/8 0.76, 64/4 0.78, 64/8 0.81-0.85
    int    bsr_u            = 8;   // its blocks in flight per thread (4 / 8), nb <= 4
    int    bsr_stages       = 2;   // its shared-memory stages per CTA (2..4)
};

struct Ctx {
    bool         inited   = false;
    int          device   = 0;
    int          sm_count = 148;
    size_t       l2_bytes = 126u << 20;
    cudaStream_t stream   = nullptr;
    cudaStream_t side     = nullptr;    // second stream: interior rows of a partitioned operator (overlap with the exchange)
    cudaStream_t launch_stream = nullptr;   // where FC_LAUNCH puts kernels: `stream`, or `side` inside a fork
    cudaEvent_t  ev_fork = nullptr, ev_join = nullptr;
    long long    launches = 0;
    bool         capturing = false;     // inside a stream capture: launches counted per replay
    long long    captured  = 0;
    Options      opt;
    long long    opt_epoch = 0;     // bumped by every set_option: invalidates captured graphs
    // scratch for grid reductions (partials + ticket), grown on demand
    double*       red_partials = nullptr;
    size_t        red_cap      = 0;
    unsigned int* red_ticket   = nullptr;
    // the same for kernels on the side stream (they may run concurrently with a reducing kernel on `stream`)
    double*       red_partials_side = nullptr;
    size_t        red_cap_side      = 0;
    unsigned int* red_ticket_side   = nullptr;
    double*       side_tot          = nullptr;   // totals of a side-stream reduction, added by the main-stream kernel
    // per-launch profile records (opt.profile): tag, events, algorithmic bytes
    struct ProfRec {
        cudaEvent_t e0, e1;
        int         kind, rows;
        long long   nnz;
        double      bytes;
    };
    std::vector<ProfRec> prof;
    // L2 flush buffer for timing helpers
    char*  flush_buf   = nullptr;
    size_t flush_bytes = 0;
};
Ctx& ctx();
void ensure_init();

#define FC_LAUNCH(kernel, grid, block, smem, ...)                                          \
    do {                                                                                   \
        ::fc::Ctx& c__ = ::fc::ctx();                                                      \
        kernel<<<(grid), (block), (smem), c__.launch_stream>>>(__VA_ARGS__);               \
        if (c__.capturing) c__.captured++; else c__.launches++;                            \
        FC_CUDA(cudaPeekAtLastError());                                                    \
    } while (0)

// profiling scope: records events around the enclosed launches when opt.profile is on
struct ProfScope {
    bool   on;
    size_t idx = 0;   // record index (scopes nest: a halo exchange inside a matrix kernel)
    ProfScope(int kind, int rows, long long nnz, double bytes);
    ~ProfScope();
};

// device allocation helpers (throw ERROR_ALLOC_MEM)
void* dmalloc(size_t bytes);
void  dfree(void* p);
template <class T> T* dalloc(size_t n) { return static_cast<T*>(dmalloc(n * sizeof(T))); }

// ------------------------------------------------------------------------------------
// device CSR matrix
// ------------------------------------------------------------------------------------
// HBM layout:
//   ia     int32[rows+1]          row offsets (matrices with nnz >= 2^31 use ia64)
//   ja     int32[nnz_pad]         column ids, padded by 8 entries so vector loads may overrun
//   val    f64  [nnz_pad]         entries (nullptr for pattern-only operators: all ones)
//   rowblk int32[nblk+1]          first row of every row block; a block is a run of rows
//                                 with <= blk_cap nonzeros and <= 256 rows, or one long row
//   diag   f64[rows], dpos int32[rows]   diagonal entry and its offset in the row (Jacobi)
//   l1     f64[rows]              sum_k |a_ik| in storage order (L1 smoother)
struct HaloPlan;   // dist.cu: ghost exchange plan of a row-partitioned operator

struct DevCSR {
    int       rows = 0, cols = 0;
    long long nnz  = 0;
    int*      ia   = nullptr;
    int*      ja   = nullptr;
    double*   val  = nullptr;
    int*      rowblk  = nullptr;
    int2*     blkdesc = nullptr;  // {first row, ia[first row]} per row-block boundary
    int       nblk    = 0;
    int       blk_cap = 0;        // shared-memory products per CTA (entries)
    int       blk_tpb = 256;      // rows per block bound = threads of the pipelined kernel
    int       vec_lpr = 0;        // > 0: vector kernel with this many lanes per row
    double*   diag = nullptr;
    int*      dpos = nullptr;
    double*   l1   = nullptr;
    double*   dinv = nullptr;     // 1/diag (first stored diagonal entry), poly smoother
    size_t    bytes = 0;
    bool      dup_diag = false;   // some row stores more than one (i,i) entry
    // multi-GPU: the longest run of rows (and of whole row blocks) without ghost columns; empty = no split
    int       int_row0 = 0, int_row1 = 0, int_blk0 = 0, int_blk1 = 0;
    int       vec_cap = 0;        // multi-GPU: entries a gathered vector holds (uniform over the ranks)
    int       nghost = 0;         // multi-GPU: ghost entries behind the owned part of a gathered vector
    HaloPlan* halo = nullptr;     // multi-GPU: ghosts of the gathered vector are exchanged before the kernel
};

// Upload a host CSR (FASP layout). `pattern_only` drops the values (UA-AMG P/R: all ones).
// ghost_col0 >= 0 (multi-GPU slabs): columns >= ghost_col0 are ghosts; the interior row / block range is recorded.
void csr_upload(DevCSR& d, int rows, int cols, long long nnz, const int* ia, const int* ja,
                const double* val, bool pattern_only = false, int ghost_col0 = -1);
void csr_free(DevCSR& d);
// lazily built smoother side data
void csr_ensure_diag(DevCSR& d);   // diag + dpos (+ dup_diag)
void csr_ensure_l1(DevCSR& d);
void csr_ensure_dinv(DevCSR& d);
double csr_dinv_a_norminf(DevCSR& d);   // || D^-1 A ||_inf (ItrSmootherCSRpoly.c:428)

// algorithmic bytes of one pass over the matrix (SURVEY.md §8d bytes model)
inline double csr_spmv_bytes(const DevCSR& d, bool read_y)
{
    double per_nnz = d.val ? 12.0 : 4.0;
    return per_nnz * (double)d.nnz + 4.0 * (d.rows + 1) + 8.0 * d.cols + 8.0 * d.rows +
           (read_y ? 8.0 * d.rows : 0.0);
}

// ------------------------------------------------------------------------------------
// fused reduction targets
// ------------------------------------------------------------------------------------
// A kernel epilogue can accumulate up to two sums over the rows it produces; the last CTA
// to finish adds the per-CTA partials in CTA order (deterministic) and stores the totals.
struct Reduce {
    const double* dot_with = nullptr;  // sum out_i * dot_with[i]
    double*       dot_out  = nullptr;  // device scalar
    double*       nrm2_out = nullptr;  // device scalar: sum out_i^2  (not square-rooted)
    bool          global   = false;    // multi-GPU: the sums are all-reduced over the ranks
};
// all-reduce the outputs of a fused reduction when it is global and a communicator is active
// (`gate`: the device flag that gated the producing kernel; the peer-memory collectives are skipped with it)
void reduce_finish(const Reduce& red, const int* gate = nullptr);
// ghost exchange (dist.cu); x must have room for the plan's ghosts behind its owned entries
void halo_exchange(const HaloPlan& h, double* x, const int* gate = nullptr, bool branch = false);

// ------------------------------------------------------------------------------------
// CSR row kernels (spmv.cu)
// ------------------------------------------------------------------------------------
enum CsrMode {
    CSR_MXV = 0,     // y = A x
    CSR_AXPY = 1,    // y = y + alpha A x
    CSR_RESID = 2,   // y = b - A x
    CSR_JACOBI = 3,  // y = (1-w) x + w (b - sum_{j!=i} a_ij x_j)/a_ii
    CSR_L1 = 4,      // y = x + (b - A x)/l1
    CSR_POLY1 = 5,   // poly smoother first step  (see spmv.cu)
    CSR_POLYJ = 6,   // poly smoother recurrence step
    CSR_RESID_DINV = 7, // y = b - A x ; aux_out = dinv .* y   (poly smoother residual)
    CSR_MXV_DIV = 8     // y = A x ; div_out = div_s * y ./ div_d  (restriction + the next level's zero-guess sweep).
                        // A mode of its own so that the code of plain y = A x stays exactly what was measured.
};

struct CsrArgs {
    int           mode  = CSR_MXV;
    double        alpha = 1.0;      // AXPY alpha / Jacobi weight
    const double* alpha_dev = nullptr;  // AXPY: alpha read from the device when set
    const double* x     = nullptr;  // gathered vector
    const double* b     = nullptr;  // rhs (RESID/JACOBI/L1/POLY*: r)
    double*       y     = nullptr;  // output
    // polynomial smoother extras
    const double* v0 = nullptr;     // POLYJ: v_{j-1}
    double*       v0_out = nullptr; // POLY1: v0 out ; POLYJ: new v0 (= old v1)
    double*       u_acc  = nullptr; // POLYJ (last step): u += vnew
    double        k1 = 0, k2 = 0, k3 = 0, k4 = 0, k5 = 0;
    // CSR_MXV_DIV extra output (restriction fused with the next level's zero-guess sweep): div_out_i = div_s * y_i / div_d_i
    double*       div_out = nullptr;
    const double* div_d   = nullptr;
    double        div_s   = 1.0;
    Reduce        red;
    const double* red_add = nullptr;     // totals of the other part of a split launch, added in the finalize step
    bool          skip_halo = false;     // multi-GPU: the ghosts of x are already up to date (computed redundantly)
    const int*    done = nullptr;
    bool          conditional = false;   // launch gated by a rarely-taken branch flag (profiling tag)
};
void csr_launch(const DevCSR& A, const CsrArgs& a);

// ------------------------------------------------------------------------------------
// BLAS-1 (blas1.cu) — all on device scalars, all early-exit on *done
// ------------------------------------------------------------------------------------
void vec_set(double* x, double v, size_t n, const int* done = nullptr);
void vec_copy(double* y, const double* x, size_t n, const int* done = nullptr);
// y = a*x + b*y with host scalars (FASP fasp_blas_darray_axpby, BlaArray.c:620)
void vec_axpby(double a, const double* x, double b, double* y, size_t n,
               const int* done = nullptr);
// y += a*x (fasp_blas_darray_axpy, BlaArray.c:90) ; x *= a (fasp_blas_darray_ax, BlaArray.c:43)
void vec_axpy(double a, const double* x, double* y, size_t n, const int* done = nullptr);
void vec_ax(double a, double* x, size_t n, const int* done = nullptr);
void vec_axpy_dev(const double* a_dev, const double* x, double* y, size_t n, const int* done = nullptr);
void vec_norm1_inf_host(const double* x, size_t n, double* norm1, double* norminf);
// y = s .* x / d  (zero-guess first sweep: L1: s=1, d=l1 ; Jacobi: s=w, d=diag)
void vec_scale_div(double* y, double s, const double* x, const double* d, size_t n,
                   const Reduce& red = Reduce(), const int* done = nullptr);
void vec_dot(const double* x, const double* y, size_t n, double* out_dev,
             const int* done = nullptr);
void vec_reduce(const double* x, size_t n, const Reduce& red, const int* done = nullptr);
void vec_mul(double* y, const double* d, const double* x, size_t n, const int* done = nullptr);
void scaling_alpha(double* s3, const int* done = nullptr);   // s[0] = min(s[1]/s[2], 1)
// host-returning reductions (synchronise): used by drop-in level-1 calls and tests
double vec_dot_host(const double* x, const double* y, size_t n);
double vec_norm2_host(const double* x, size_t n);

// grid-reduction scratch: returns partial buffer large enough for `nblocks` CTAs x 2 sums
double*       red_partials(size_t nblocks);
unsigned int* red_ticket();

// ------------------------------------------------------------------------------------
// setup-phase pieces on the device (setup.cu): A^T and R A P in the reference's output layout
// ------------------------------------------------------------------------------------
void setup_transpose(const dCSRmat* A, dCSRmat* AT);
void setup_rap(const dCSRmat* R, const dCSRmat* A, const dCSRmat* P, dCSRmat* RAP);

// ------------------------------------------------------------------------------------
// timing helpers
// ------------------------------------------------------------------------------------
void flush_l2();

} // namespace fc

// ------------------------------------------------------------------------------------
// CUDA-graph capture of a fixed kernel sequence issued on ctx().stream
// ------------------------------------------------------------------------------------
namespace fc {
void p2p_reset_order_hook();   // p2p.cu: a captured graph must not assume which vector was exchanged last
struct CapturedGraph {
    cudaGraph_t     graph   = nullptr;
    cudaGraphExec_t exec    = nullptr;
    long long       kernels = 0;
    ~CapturedGraph() { reset(); }
    void reset()
    {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        exec  = nullptr;
        graph = nullptr;
    }
    // Capture `f` (stream capture, nothing executes) and instantiate it, without launching: lets a
    // solver build its graphs before the timed region starts.
    template <class F> void prepare(bool enable, F&& f)
    {
        if (!enable || exec) return;
        capture(f);
    }
    // First call captures `f` and instantiates it (unless prepare() did); every call then launches the
    // instantiated graph. With enable == false, f runs directly.
    template <class F> void run(bool enable, F&& f)
    {
        Ctx& c = ctx();
        if (!enable) {
            f();
            return;
        }
        if (!exec) capture(f);
        FC_CUDA(cudaGraphLaunch(exec, c.stream));
        c.launches += kernels;
    }
    template <class F> void capture(F&& f)
    {
        Ctx& c = ctx();
        {
            p2p_reset_order_hook();
            c.capturing = true;
            c.captured  = 0;
            cudaError_t e = cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal);
            if (e != cudaSuccess) {
                c.capturing = false;
                FC_CUDA(e);
            }
            try {
                f();
            } catch (...) {
                cudaGraph_t tmp = nullptr;
                cudaStreamEndCapture(c.stream, &tmp);
                if (tmp) cudaGraphDestroy(tmp);
                c.capturing = false;
                throw;
            }
            c.capturing = false;
            FC_CUDA(cudaStreamEndCapture(c.stream, &graph));
            kernels = c.captured;
            FC_CUDA(cudaGraphInstantiate(&exec, graph, 0));
            p2p_reset_order_hook();   // what the capture "exchanged last" never ran
        }
    }
};
} // namespace fc
