// amg.cuh — device-resident AMG hierarchy and multigrid cycle (CSR and BSR).
#pragma once
#include "common.cuh"
#include "bsr.cuh"

namespace fc {

// dense coarsest-level operator: x = Ainv * b
struct DenseInv {
    int     n    = 0;
    double* ainv = nullptr;   // n x n row-major
};
// Blocked Gauss-Jordan with partial pivoting (dense.cu). Return false, with D left empty, when the matrix is
// numerically singular (a pivot below 1e-14 of the largest entry): the caller falls back to the iterative
// coarse solve or reports ERROR_AMG_SETUP instead of applying an inf/NaN inverse.
bool dense_invert_csr(DenseInv& D, const DevCSR& A);
bool dense_invert_host(DenseInv& D, int n, const std::vector<double>& a_rowmajor);
bool dense_invert_bsr(DenseInv& D, const DevBSR& A);          // expanded to (ROW nb)^2
void dense_apply(const DenseInv& D, const double* b, double* x, const int* done);
void dense_free(DenseInv& D);

// iterative coarsest-level solve (coarse.cu): CG to tol in one persistent cooperative kernel
struct CoarseCG {
    int     n = 0, grid = 0, maxit = 0;
    double  tol  = 1e-10;
    double* work = nullptr;   // p, r, t (n each) + per-CTA partials
    int*    iters = nullptr;  // device: iterations of the last solve
};
void coarse_cg_setup(CoarseCG& C, const DevCSR& A, double tol);
void coarse_cg_apply(const CoarseCG& C, const DevCSR& A, const double* b, double* x, const int* done);
void coarse_cg_free(CoarseCG& C);
int  coarse_cg_last_iters(const CoarseCG& C);

struct Level {
    DevCSR  A, P, R;          // P: level l <- l+1 ; R: level l+1 <- l   (R = P^T stored explicitly)
    int     n  = 0;          // local rows (= global rows on one GPU)
    int     cap = 0;          // entries allocated per level vector: n + room for ghosts
    bool    dist = false;     // multi-GPU: rows of this level are partitioned over the ranks
    int     nglobal = 0, row0 = 0;
    std::vector<size_t> gcounts, gdispls;   // next level replicated: every rank's slice of its vectors
    HaloPlan *hA = nullptr, *hP = nullptr, *hR = nullptr;
    bool    p2p_registered = false;
    // redundant ghost rows (dist.cu): P_l also updates the ghost entries of x_l (those A_l gathers), R_l also
    // produces the ghost entries of b_{l+1}; dscale_ext = the smoother's divisor (l1 / diagonal) incl. ghost rows
    bool    p_ext = false, r_ext = false;
    double* dscale_ext = nullptr;
    double* b  = nullptr;     // right-hand side on this level (level 0: caller's r)
    double* xa = nullptr;     // iterate ping-pong buffers
    double* xb = nullptr;
    double* w  = nullptr;     // residual / scratch
    // polynomial smoother (ItrSmootherCSRpoly.c:94-107)
    double  pk[6] = {0, 0, 0, 0, 0, 0};
    double* pv[3] = {nullptr, nullptr, nullptr};   // rbar/v0, v1, vnew rotation
    // multicolour Gauss-Seidel (BlaSparseCSR.c:1687,2123)
    int  ncolors = 0;
    int* color_rows = nullptr;       // rows grouped by colour
    std::vector<int> color_ptr;      // host offsets into color_rows
};

struct Amg {
    std::vector<Level> lv;
    int    nl = 0;
    bool   dist = false;            // multi-GPU hierarchy (dist.cu)
    std::vector<int> off0;          // level-0 row partition (size nranks+1)
    // parameters (AMG_param / precond_data members used by the cycle, PreMGCycle.c:50-59)
    short  amg_type = CLASSIC_AMG, smoother = SMOOTHER_GS, cycle_type = V_CYCLE;
    short  presmooth = 1, postsmooth = 1, ndeg = 3, coarse_scaling = OFF, coarse_solver = 0;
    short  smooth_order = CF_ORDER;
    double relax = 1.0, tol = 1e-6;
    int    maxit = 1;             // cycles per preconditioner application (PreCSR.c:432)
    DenseInv coarse;
    CoarseCG coarse_cg;           // used instead of `coarse` when the dense inverse is unavailable
    bool     coarse_iterative = false;
    double*  scal = nullptr;      // device scalars: [0] alpha (coarse scaling), [1],[2] dots
    size_t   bytes = 0;
    long long kernels_per_cycle = 0;
};

// multicolour Gauss-Seidel (gs.cu)
void gs_multicolor_host(int n, const int* IA, const int* JA, std::vector<int>& IC, std::vector<int>& ICMAP);
void gs_multicolor_sweeps(const DevCSR& A, const int* color_rows, const std::vector<int>& color_ptr,
                          const double* b, double* u, int L, int order, const int* done);

// Build from a host hierarchy produced by FASP's setup (fasp.h:804-888).
Amg* amg_upload(AMG_data* mgl, AMG_param* param);
void amg_setup_coarse(Amg& h);   // dense inverse of the coarsest level, or the iterative fallback
void amg_free(Amg* h);
void amg_set_params(Amg& h, const AMG_param* param);
void amg_level_smoother_data(Amg& h, Level& L, const dCSRmat* hostA);
void amg_level_poly_vectors(Amg& h, Level& L);
void amg_level_vectors(Amg& h, Level& L);

// One preconditioner application z = B r: x0 = 0, h.maxit cycles (PreCSR.c:416-435 +
// PreMGCycle.c:48-274). r is used in place as the level-0 right-hand side; the result is
// written to z. `red` (optional) is fused into the last kernel that writes z.
void amg_apply(Amg& h, const double* r, double* z, const Reduce& red, const int* done);

// One cycle on explicit level-0 vectors with a non-zero initial guess in x (in/out):
// used by fasp_cuda_solver_mgcycle and the AMG-as-solver loop (PreMGSolve.c:49).
void amg_cycle_inplace(Amg& h, const double* b, double* x, bool x_is_zero, const Reduce& red,
                       const int* done);

// ---- BSR twin (PreMGCycle.c:287-566, PreBSR.c:1149) ----
struct BLevel {
    DevBSR  A, P, R;
    int     n = 0;            // ROW * nb
    double* b = nullptr;
    double* xa = nullptr;
    double* xb = nullptr;
    double* w = nullptr;
    double* diaginv = nullptr;   // ROW * nb*nb inverted diagonal blocks
    double *peh = nullptr, *aeh = nullptr;   // coarse_scaling: P e and A P e (PreMGCycle.c:467-471)
};
struct BAmg {
    std::vector<BLevel> lv;
    int    nl = 0, nb = 1;
    short  smoother = SMOOTHER_JACOBI, cycle_type = V_CYCLE, presmooth = 1, postsmooth = 1;
    short  coarse_scaling = OFF;
    double relax = 1.0, tol = 1e-6;
    int    maxit = 1;
    DenseInv coarse;
    double*  scal = nullptr;   // device scalars of the coarse-grid scaling
    size_t bytes = 0;
};
BAmg* bamg_upload(AMG_data_bsr* mgl, AMG_param* param);
void  bamg_free(BAmg* h);
void  bamg_apply(BAmg& h, const double* r, double* z, const Reduce& red, const int* done);
void  bamg_cycle_inplace(BAmg& h, const double* b, double* x, bool x_is_zero, const Reduce& red,
                         const int* done);

} // namespace fc
