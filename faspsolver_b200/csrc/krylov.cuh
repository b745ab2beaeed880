// krylov.cuh — device-resident Krylov loops (PCG, GMRES(m), vGMRES) over abstract operators.
#pragma once
#include "common.cuh"
#include "amg.cuh"

namespace fc {

// y = A x style operator on device vectors (CSR or BSR behind it)
struct LinOp {
    int n = 0;
    virtual ~LinOp() {}
    virtual const void* key() const { return this; }   // identity of the resident operator
    virtual size_t      vec_capacity() const { return (size_t)n; }   // entries a gathered vector must hold
    virtual const char* format() const { return "CSR"; }             // name in the "Calling ... solver" line
    // rows partitioned over the ranks: dot products / norms of the Krylov loop are all-reduced. A plain
    // single-GPU operator stays local even while a communicator is active (rank 0 of a multi-GPU job can
    // run the one-GPU solve next to the partitioned one: bench.py's parity line).
    virtual bool        distributed() const { return false; }
    // mode: CSR_MXV / CSR_AXPY / CSR_RESID semantics
    virtual void apply(int mode, double alpha, const double* x, const double* b, double* y,
                       const Reduce& red, const int* done, bool conditional = false) = 0;
};
struct CsrOp : LinOp {
    const DevCSR* A;
    explicit CsrOp(const DevCSR* a) : A(a) { n = a->rows; }
    const void* key() const override { return A; }
    bool        distributed() const override { return A->halo != nullptr; }
    size_t      vec_capacity() const override
    {
        return A->vec_cap > 0 ? (size_t)A->vec_cap : (size_t)A->rows + (size_t)A->nghost;
    }
    void apply(int mode, double alpha, const double* x, const double* b, double* y,
               const Reduce& red, const int* done, bool conditional = false) override
    {
        CsrArgs a;
        a.conditional = conditional;
        a.mode  = mode;
        a.alpha = alpha;
        a.x     = x;
        a.b     = b;
        a.y     = y;
        a.red   = red;
        a.done  = done;
        csr_launch(*A, a);
    }
};

struct BsrOp : LinOp {
    const DevBSR* A;
    explicit BsrOp(const DevBSR* a) : A(a) { n = a->ROW * a->nb; }
    const void* key() const override { return A; }
    const char* format() const override { return "BSR"; }
    void apply(int mode, double alpha, const double* x, const double* b, double* y,
               const Reduce& red, const int* done, bool conditional = false) override
    {
        BsrArgs a;
        a.conditional = conditional;
        a.mode  = (mode == CSR_MXV) ? BSR_MXV : (mode == CSR_AXPY ? BSR_AXPY : BSR_RESID);
        a.alpha = alpha;
        a.x     = x;
        a.b     = b;
        a.y     = y;
        a.red   = red;
        a.done  = done;
        bsr_launch(*A, a);
    }
};

// z = B r on device vectors
struct Prec {
    virtual ~Prec() {}
    virtual const void* key() const { return this; }
    virtual void apply(const double* r, double* z, const Reduce& red, const int* done) = 0;
    virtual bool capturable() const { return true; }
};
struct IdentityPrec : Prec {
    size_t n;
    explicit IdentityPrec(size_t n_) : n(n_) {}
    void apply(const double* r, double* z, const Reduce& red, const int* done) override
    {
        vec_copy(z, r, n, done);
        vec_reduce(z, n, red, done);
    }
};
struct AmgPrec : Prec {
    Amg* h;
    explicit AmgPrec(Amg* h_) : h(h_) {}
    const void* key() const override { return h; }
    void apply(const double* r, double* z, const Reduce& red, const int* done) override
    {
        amg_apply(*h, r, z, red, done);
    }
};
struct BAmgPrec : Prec {
    BAmg* h;
    explicit BAmgPrec(BAmg* h_) : h(h_) {}
    const void* key() const override { return h; }
    void apply(const double* r, double* z, const Reduce& red, const int* done) override
    {
        bamg_apply(*h, r, z, red, done);
    }
};
// arbitrary host callback pc->fct(r, z, data) with host pointers: D2H r, call, H2D z
struct HostPrec : Prec {
    precond* pc;
    size_t   n;
    double * hr = nullptr, *hz = nullptr;
    HostPrec(precond* pc_, size_t n_);
    ~HostPrec() override;
    void apply(const double* r, double* z, const Reduce& red, const int* done) override;
    bool capturable() const override { return false; }
};

struct SolveStats {
    int       iters    = 0;
    double    relres   = 0.0;
    double    ms       = 0.0;     // device time of the Krylov loop (CUDA events)
    long long launches = 0;
    std::vector<double> hist_relres, hist_absres, hist_factor;
};

// Workspace + captured graphs of a PCG solve, reusable across solves on the same operator,
// preconditioner and vectors (a solver object keeps one: no allocation, capture or
// instantiation on the repeated-solve path).
struct PcgCache {
    double* work = nullptr;
    void*   st   = nullptr;
    int*    pin  = nullptr;
    size_t  n    = 0, ncap = 0;
    bool    registered = false;   // work is peer-mapped (multi-GPU)
    int     hcap = 0, look = 0;
    std::vector<cudaEvent_t> ev;
    cudaEvent_t   t0 = nullptr, t1 = nullptr;
    CapturedGraph g_init, g_iter;
    const void *  kA = nullptr, *kb = nullptr, *ku = nullptr, *kpc = nullptr;
    int           kstop = 0;
    long long     kepoch = -1;
    ~PcgCache();
    void release();
};

// The same for GMRES: basis / Hessenberg workspace, state, pinned status word and one captured graph per
// inner-step index (the basis pointers are baked into the nodes) survive across solves, so a repeated
// solve allocates, captures and instantiates nothing.
struct GmresCache {
    double* work = nullptr;
    void*   st   = nullptr;
    void *  pin_d = nullptr, *pin_h = nullptr;
    int*    pin_flags = nullptr;          // pinned ring of skip_inner flags read `look` steps behind the launches
    std::vector<cudaEvent_t> ev;
    int     look = 0;
    size_t  n = 0, ldp = 0;
    int     R = 0, hcap = 0, kind = -1;
    bool    registered = false;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    std::vector<CapturedGraph> step_graph;
    CapturedGraph start_graph, end_graph, init_graph;
    const void *  kA = nullptr, *kb = nullptr, *kx = nullptr, *kpc = nullptr;
    int           kstop = 0;
    long long     kepoch = -1;
    ~GmresCache();
    void release();
};

// All vectors are device pointers. Returns FASP status (>=0 iterations, <0 ERROR_*).
int pcg_solve(LinOp& A, const double* b, double* u, Prec& pc, double tol, double abstol,
              int MaxIt, int StopType, int PrtLvl, SolveStats* stats, PcgCache* cache = nullptr);
// restart > 0; kind: fixed restart (KryPgmres.c), Baker/Jessup/Kolev restart adaptation
// (KryPvgmres.c), or the flexible variant that stores the preconditioned basis (KryPvfgmres.c)
enum { GM_FIXED = 0, GM_VARIABLE = 1, GM_FLEXIBLE = 2 };
int gmres_solve(LinOp& A, const double* b, double* x, Prec& pc, double tol, double abstol,
                int MaxIt, int restart, int StopType, int PrtLvl, int kind,
                SolveStats* stats, GmresCache* cache = nullptr);

// FASP's iteration table / final line (AuxMessage.c:41-76, KryUtil.inl:93-103)
void print_itinfo(int prtlvl, int stop_type, int iter, double relres, double absres,
                  double factor);
void print_final(int iter, int maxit, double relres);

} // namespace fc
