/*
 * fasp_cuda.h — C-ABI of libfasp_cuda: the B200 (sm_100a) solve-phase hot path of FASP.
 *
 * Every entry point below is `extern "C"`, takes plain pointers / FASP's own structs by
 * pointer and returns FASP's INT status convention (>= 0 iteration count, < 0 ERROR_*).
 * Each declaration cites the reference interface it replaces as `file:line` relative to
 * the FASP 2.8.7 source tree (base/...).
 *
 * Usage from an existing FASP application (C99):
 *
 *     #include "fasp.h"          // the application's own FASP headers (optional)
 *     #include "fasp_functs.h"
 *     #include "fasp_cuda.h"     // sees __FASP_HEADER__ and re-uses FASP's types
 *     ...
 *     fasp_cuda_abi_check(sizeof(dCSRmat), sizeof(AMG_data), sizeof(AMG_param));
 *     status = fasp_cuda_solver_dcsr_krylov_amg(&A, &b, &x, &itparam, &amgparam);
 *
 * When fasp.h is NOT included first, this header defines layout-identical mirror types
 * (same names, same member names, same order) so a caller can be built without the FASP
 * tree. The mirror is for the sequential (non-OpenMP) FASP ABI unless
 * FASP_CUDA_OPENMP_ABI is defined (the OpenMP build of FASP adds members to dCSRmat and
 * AMG_data: base/include/fasp.h:171-178, 883-886).
 */
#ifndef FASP_CUDA_H
#define FASP_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------ */
/* Types: taken from fasp.h / fasp_block.h when present, mirrored otherwise              */
/* ------------------------------------------------------------------------------------ */
#if defined(__FASP_HEADER__)
#  ifndef __FASPBLOCK_HEADER__
#    include "fasp_block.h"
#  endif
#else
#  include "fasp_cuda_types.h"
#endif

/* opaque device-side objects (owned by the library, freed by the matching *_free) */
typedef struct fasp_cuda_csr_s  fasp_cuda_csr;   /* a CSR matrix resident in HBM        */
typedef struct fasp_cuda_bsr_s  fasp_cuda_bsr;   /* a BSR matrix resident in HBM        */
typedef struct fasp_cuda_amg_s  fasp_cuda_amg;   /* an uploaded AMG hierarchy + scratch */
typedef struct fasp_cuda_bamg_s fasp_cuda_bamg;  /* an uploaded BSR AMG hierarchy       */

/* ------------------------------------------------------------------------------------ */
/* Library / ABI                                                                          */
/* ------------------------------------------------------------------------------------ */

/* Returns 0 if the caller's struct sizes equal the ones this library was built with
 * (guards against mixing an OpenMP-ABI FASP with a sequential-ABI libfasp_cuda:
 * base/include/fasp.h:42-50,171-178,883-886). Nonzero = ERROR_DATA_STRUCTURE (-21). */
INT fasp_cuda_abi_check(size_t sizeof_dCSRmat, size_t sizeof_AMG_data, size_t sizeof_AMG_param);

/* Human readable description of the last error raised on the calling thread. */
const char* fasp_cuda_last_error(void);

/* Selects the CUDA device for this process (one process per GPU). Returns 0 or ERROR_*. */
INT fasp_cuda_init(int device);

/* Number of kernels launched by this library since load / since the last reset. */
long long fasp_cuda_launch_count(void);
void      fasp_cuda_launch_count_reset(void);

/* With option "profile" = 1 every matrix kernel is bracketed by CUDA events (graphs off);
 * this returns the records gathered since the last call as text lines
 * "<kind> <rows> <nnz> <ms> <algorithmic bytes>" and clears them. */
long long fasp_cuda_profile_dump(char* buf, long long cap);

/* Options that have no slot in FASP's parameter structs (ABI stays unchanged):
 *   "strict"       0/1  every CSR row summed left-to-right without FMA (bit-identical to
 *                       the sequential CPU loops, slower on long rows)           default 0
 *   "coarse_dense" 0/1  coarsest level solved by a precomputed dense inverse instead of the
 *                       reference's iterative SPCG (PreMGUtil.inl:37-58)         default 1
 *   "coarse_dense_max"  largest coarsest-level size solved densely               default 8192
 *   "graph"        0/1  capture V-cycle / Krylov iteration into CUDA graphs      default 1
 *   "zero_guess"   0/1  skip the A-pass of the first pre-smoothing sweep when x==0
 *                       (algebraically identical)                 default 1
 * Returns 0, or ERROR_INPUT_PAR for an unknown key. */
INT fasp_cuda_set_option(const char* key, double value);
double fasp_cuda_get_option(const char* key);

/* ------------------------------------------------------------------------------------ */
/* Level-1 drop-ins with HOST pointers (H2D, kernel, D2H inside the call)                 */
/* ------------------------------------------------------------------------------------ */

/* y = A*x.                         replaces fasp_blas_dcsr_mxv      BlaSpmvCSR.c:242 */
INT fasp_cuda_blas_dcsr_mxv(const dCSRmat* A, const REAL* x, REAL* y);
/* y = y + alpha*A*x.               replaces fasp_blas_dcsr_aAxpy    BlaSpmvCSR.c:494 */
INT fasp_cuda_blas_dcsr_aAxpy(const REAL alpha, const dCSRmat* A, const REAL* x, REAL* y);
/* y = A*x, entries of A taken as 1 replaces fasp_blas_dcsr_mxv_agg  BlaSpmvCSR.c:438 */
INT fasp_cuda_blas_dcsr_mxv_agg(const dCSRmat* A, const REAL* x, REAL* y);
/* y = y + alpha*A*x, A entries = 1 replaces fasp_blas_dcsr_aAxpy_agg BlaSpmvCSR.c:727 */
INT fasp_cuda_blas_dcsr_aAxpy_agg(const REAL alpha, const dCSRmat* A, const REAL* x, REAL* y);
/* returns y'*A*x (NaN on failure). replaces fasp_blas_dcsr_vmv      BlaSpmvCSR.c:839 */
REAL fasp_cuda_blas_dcsr_vmv(const dCSRmat* A, const REAL* x, const REAL* y);
/* y = A*x (block CSR).             replaces fasp_blas_dbsr_mxv      BlaSpmvBSR.c:1055 */
INT fasp_cuda_blas_dbsr_mxv(const dBSRmat* A, const REAL* x, REAL* y);
/* y = y + alpha*A*x (block CSR).   replaces fasp_blas_dbsr_aAxpy    BlaSpmvBSR.c:514 */
INT fasp_cuda_blas_dbsr_aAxpy(const REAL alpha, const dBSRmat* A, const REAL* x, REAL* y);

/* BLAS-1 with host pointers (BlaArray.c). Element-wise results are bit-identical to the CPU
 * loops; the reductions return NaN on failure and agree to <= 1e-14 relative (tree sums).
 * x = a*x.                         replaces fasp_blas_darray_ax       BlaArray.c:43  */
INT fasp_cuda_blas_darray_ax(const INT n, const REAL a, REAL* x);
/* y = a*x + y.                     replaces fasp_blas_darray_axpy     BlaArray.c:90  */
INT fasp_cuda_blas_darray_axpy(const INT n, const REAL a, const REAL* x, REAL* y);
/* y = a*x + b*y.                   replaces fasp_blas_darray_axpby    BlaArray.c:620 */
INT fasp_cuda_blas_darray_axpby(const INT n, const REAL a, const REAL* x, const REAL b, REAL* y);
/* (x, y).                          replaces fasp_blas_darray_dotprod  BlaArray.c:771 */
REAL fasp_cuda_blas_darray_dotprod(const INT n, const REAL* x, const REAL* y);
/* ||x||_2, ||x||_1, ||x||_inf.     replace fasp_blas_darray_norm2/_norm1/_norminf BlaArray.c:691,663,719 */
REAL fasp_cuda_blas_darray_norm2(const INT n, const REAL* x);
REAL fasp_cuda_blas_darray_norm1(const INT n, const REAL* x);
REAL fasp_cuda_blas_darray_norminf(const INT n, const REAL* x);

/* matrix-free operator plug-in (fasp.h:1109-1117 `mxv_matfree.fct`, shims in
 * BlaSpmvMatFree.inl:31-107): `A` is a const dCSRmat* / const dBSRmat*.                */
void fasp_cuda_blas_mxv_csr(const void* A, const REAL* x, REAL* y);
void fasp_cuda_blas_mxv_bsr(const void* A, const REAL* x, REAL* y);
/* replaces fasp_solver_matfree_init  SolMatFree.c:201 for MAT_CSR (1) / MAT_BSR (2): sets mf->fct to the
 * functions above and mf->data = A, so that FASP's own matrix-free Krylov loops (fasp_solver_pcg KryPcg.c:1260,
 * fasp_solver_pvgmres KryPvgmres.c:1468) run their SpMV on the device.                                      */
INT fasp_cuda_solver_matfree_init(INT matrix_format, mxv_matfree* mf, void* A);

/* Dense inverse of an n x n row-major matrix (host pointers) by the blocked Gauss-Jordan kernel that factors
 * the coarsest AMG level in place of fasp_coarse_itsolver (PreMGUtil.inl:37). ERROR_AMG_SETUP if singular. */
INT fasp_cuda_dense_inverse(INT n, const REAL* a, REAL* ainv);

/* Smoothers, same argument lists as the reference (host pointers).
 * replaces fasp_smoother_dcsr_jacobi   ItrSmootherCSR.c:98   */
INT fasp_cuda_smoother_dcsr_jacobi(dvector* u, const INT i_1, const INT i_n, const INT s,
                                   dCSRmat* A, dvector* b, INT L, const REAL w);
/* replaces fasp_smoother_dcsr_L1diag   ItrSmootherCSR.c:1509 */
INT fasp_cuda_smoother_dcsr_L1diag(dvector* u, const INT i_1, const INT i_n, const INT s,
                                   dCSRmat* A, dvector* b, INT L);
/* replaces fasp_smoother_dcsr_poly     ItrSmootherCSRpoly.c:67 */
INT fasp_cuda_smoother_dcsr_poly(dCSRmat* Amat, dvector* brhs, dvector* usol, INT n, INT ndeg,
                                 INT L);
/* replaces fasp_smoother_dcsr_gs_multicolor BlaSparseCSR.c:2123 (colouring recomputed by the
 * library with the reference's greedy rule, BlaSparseCSR.c:1687)                         */
INT fasp_cuda_smoother_dcsr_gs_multicolor(dvector* u, dCSRmat* A, dvector* b, INT L, INT order);
/* the colouring itself (host, no GPU): IC[ncolors+1] offsets into ICMAP[row] as dCSRmat_Multicoloring
 * (BlaSparseCSR.c:1687) would leave them in A->IC / A->ICMAP; returns the number of colours        */
INT fasp_cuda_multicolor_host(INT n, const INT* IA, const INT* JA, INT* IC, INT* ICMAP);
/* replaces fasp_smoother_dbsr_jacobi1  ItrSmootherBSR.c:263 (diaginv = inverted diagonal
 * blocks, nb*nb per block row, as produced by fasp_smoother_dbsr_jacobi_setup :163)      */
INT fasp_cuda_smoother_dbsr_jacobi1(dBSRmat* A, dvector* b, dvector* u, REAL* diaginv);

/* ------------------------------------------------------------------------------------ */
/* Setup-phase pieces on the device (the two matrix-matrix steps of every AMG level)      */
/* ------------------------------------------------------------------------------------ */

/* AT = A^T, host pointers; AT->IA/JA/val are calloc'ed (free them as FASP does, fasp_dcsr_free). The result is
 * identical to the reference's, entry for entry.        replaces fasp_dcsr_trans     BlaSparseCSR.c:952 */
INT fasp_cuda_dcsr_trans(const dCSRmat* A, dCSRmat* AT);
/* RAP = R*A*P (Galerkin coarse operator), host pointers, output calloc'ed. Same entry ORDER inside every row
 * (diagonal first, then first-met order of the R->A->P walk) and the same bits in every value as the
 * reference's sequential code.                          replaces fasp_blas_dcsr_rap  BlaSpmvCSR.c:999   */
INT fasp_cuda_blas_dcsr_rap(const dCSRmat* R, const dCSRmat* A, const dCSRmat* P, dCSRmat* RAP);
/* libfasp_cuda_setup.so (same directory) defines fasp_dcsr_trans / fasp_blas_dcsr_rap themselves and forwards
 * them to the two functions above: LD_PRELOAD it (or link it before libfasp) and FASP's own fasp_amg_setup_rs /
 * _sa run their transposes and triple products on the device, unmodified. See INTEGRATION.md.             */

/* ------------------------------------------------------------------------------------ */
/* Device-resident objects                                                                */
/* ------------------------------------------------------------------------------------ */

/* Upload a CSR / BSR matrix once; kernels then run on the resident copy. */
fasp_cuda_csr* fasp_cuda_dcsr_upload(const dCSRmat* A);
void           fasp_cuda_dcsr_free(fasp_cuda_csr* dA);
fasp_cuda_bsr* fasp_cuda_dbsr_upload(const dBSRmat* A);
void           fasp_cuda_dbsr_free(fasp_cuda_bsr* dA);

/* Device vectors are plain `REAL*` in HBM. */
REAL* fasp_cuda_dvec_alloc(size_t n);
void  fasp_cuda_dvec_free(REAL* d);
INT   fasp_cuda_dvec_h2d(REAL* d, const REAL* h, size_t n);
INT   fasp_cuda_dvec_d2h(REAL* h, const REAL* d, size_t n);
INT   fasp_cuda_sync(void);

/* Resident-matrix kernels on DEVICE vectors (asynchronous on the library stream).
 * mode: 0 y=A*x ; 1 y+=alpha*A*x ; 2 y=b-A*x (b passed in `b`, may alias y)              */
INT fasp_cuda_dcsr_spmv_dev(const fasp_cuda_csr* dA, int mode, REAL alpha, const REAL* x,
                            const REAL* b, REAL* y);
INT fasp_cuda_dbsr_spmv_dev(const fasp_cuda_bsr* dA, int mode, REAL alpha, const REAL* x,
                            const REAL* b, REAL* y);
/* One smoother sweep on device vectors: kind = SMOOTHER_JACOBI (1) | SMOOTHER_L1DIAG (10). */
INT fasp_cuda_dcsr_smooth_dev(const fasp_cuda_csr* dA, int kind, REAL w, const REAL* b,
                              const REAL* u_in, REAL* u_out);

/* Timing helper for benchmarks: runs `reps` launches of the given resident-kernel after
 * `warm` warm-ups and returns the mean milliseconds per launch measured with CUDA events on
 * the library stream (0 y=Ax, 1 y+=aAx, 2 r=b-Ax, 10 Jacobi sweep, 11 L1 sweep).
 * If flush_l2 != 0 a >L2-sized buffer is rewritten between launches (outside the events). */
double fasp_cuda_dcsr_time_kernel(const fasp_cuda_csr* dA, int what, int warm, int reps,
                                  int flush_l2);
double fasp_cuda_dbsr_time_kernel(const fasp_cuda_bsr* dA, int what, int warm, int reps,
                                  int flush_l2);

/* ------------------------------------------------------------------------------------ */
/* AMG hierarchy (built by FASP's own host setup) -> HBM                                  */
/* ------------------------------------------------------------------------------------ */

/* Walk mgl[0..num_levels) (fasp.h:804-888) and upload A/P/R of every level, precompute the
 * per-level smoother data (diagonal, l1 row sums, poly constants, colour sets) and the
 * coarsest-level factor. `param` supplies smoother/cycle parameters (fasp.h:455-595).     */
fasp_cuda_amg* fasp_cuda_amg_upload(AMG_data* mgl, AMG_param* param);
void           fasp_cuda_amg_free(fasp_cuda_amg* h);
/* bytes of HBM held by the hierarchy, number of levels */
size_t fasp_cuda_amg_bytes(const fasp_cuda_amg* h);
INT    fasp_cuda_amg_levels(const fasp_cuda_amg* h);

/* BSR twin: mgl of AMG_data_bsr (fasp_block.h:146-247). */
fasp_cuda_bamg* fasp_cuda_bamg_upload(AMG_data_bsr* mgl, AMG_param* param);
void            fasp_cuda_bamg_free(fasp_cuda_bamg* h);

/* ------------------------------------------------------------------------------------ */
/* Level-4: multigrid cycle and preconditioner callbacks                                  */
/* ------------------------------------------------------------------------------------ */

/* One V/W/VW/WV cycle on mgl[0].b -> mgl[0].x with host data (uploads the hierarchy, runs
 * the device cycle, writes mgl[0].x.val). replaces fasp_solver_mgcycle  PreMGCycle.c:48   */
INT fasp_cuda_solver_mgcycle(AMG_data* mgl, AMG_param* param);
/* BSR twin.                               replaces fasp_solver_mgcycle_bsr PreMGCycle.c:287 */
INT fasp_cuda_solver_mgcycle_bsr(AMG_data_bsr* mgl, AMG_param* param);

/* Cycle on a resident hierarchy, device vectors: z = B r (x0 = 0, `ncycles` cycles). */
INT fasp_cuda_amg_cycle_dev(fasp_cuda_amg* h, const REAL* r_dev, REAL* z_dev);
/* Same with host vectors (H2D r, cycle, D2H z). */
INT fasp_cuda_amg_cycle_host(fasp_cuda_amg* h, const REAL* r, REAL* z);

/* precond plug-in (fasp.h:1095-1103): pc->fct(r, z, pc->data) with HOST r,z.
 * replaces fasp_precond_amg    PreCSR.c:416 ; `data` must come from fasp_cuda_precond_setup */
void fasp_cuda_precond_amg(REAL* r, REAL* z, void* data);
/* replaces fasp_precond_setup  PreCSR.c:46 (type must be PREC_AMG; the hierarchy is built by
 * the host application's FASP setup routines, then uploaded).  Free with .._precond_free.  */
precond* fasp_cuda_precond_setup(const SHORT precond_type, AMG_param* amgparam,
                                 ILU_param* iluparam, dCSRmat* A);
void     fasp_cuda_precond_free(precond* pc);
/* Wrap an already built host hierarchy into a device-backed precond. */
precond* fasp_cuda_precond_from_mgl(AMG_data* mgl, AMG_param* amgparam);

/* ------------------------------------------------------------------------------------ */
/* Level-3: Krylov loops, same argument lists as the reference                            */
/* ------------------------------------------------------------------------------------ */

/* replaces fasp_solver_dcsr_pcg      KryPcg.c:96.   `pc` may be NULL (identity), a precond
 * made by fasp_cuda_precond_* (the loop then never leaves the device) or any host callback
 * (each apply costs D2H r / H2D z).                                                       */
INT fasp_cuda_solver_dcsr_pcg(dCSRmat* A, dvector* b, dvector* u, precond* pc, const REAL tol,
                              const REAL abstol, const INT MaxIt, const SHORT StopType,
                              const SHORT PrtLvl);
/* replaces fasp_solver_dcsr_pvgmres  KryPvgmres.c:66 */
INT fasp_cuda_solver_dcsr_pvgmres(dCSRmat* A, dvector* b, dvector* x, precond* pc,
                                  const REAL tol, const REAL abstol, const INT MaxIt,
                                  const SHORT restart, const SHORT StopType,
                                  const SHORT PrtLvl);
/* replaces fasp_solver_dcsr_pgmres   KryPgmres.c:66 */
INT fasp_cuda_solver_dcsr_pgmres(dCSRmat* A, dvector* b, dvector* x, precond* pc,
                                 const REAL tol, const REAL abstol, const INT MaxIt,
                                 const SHORT restart, const SHORT StopType,
                                 const SHORT PrtLvl);
/* replaces fasp_solver_amg            SolAMG.c:49  (AMG as a solver: host setup, then device cycles until
 * the relative residual drops below param->tol or param->maxit cycles; AMLI cycles are rejected)           */
INT fasp_cuda_solver_amg(dCSRmat* A, dvector* b, dvector* x, AMG_param* param);
/* replaces fasp_solver_dcsr_pvfgmres KryPvfgmres.c:67 (flexible: stores z_j = B p_j, so the
 * preconditioner may change between iterations; stops on ||r|| <= tol * ||b||)            */
INT fasp_cuda_solver_dcsr_pvfgmres(dCSRmat* A, dvector* b, dvector* x, precond* pc,
                                   const REAL tol, const REAL abstol, const INT MaxIt,
                                   const SHORT restart, const SHORT StopType,
                                   const SHORT PrtLvl);
/* replaces fasp_solver_dbsr_pcg      KryPcg.c:386 */
INT fasp_cuda_solver_dbsr_pcg(dBSRmat* A, dvector* b, dvector* u, precond* pc, const REAL tol,
                              const REAL abstol, const INT MaxIt, const SHORT StopType,
                              const SHORT PrtLvl);
/* replaces fasp_solver_dbsr_pvgmres  KryPvgmres.c:416 */
INT fasp_cuda_solver_dbsr_pvgmres(dBSRmat* A, dvector* b, dvector* x, precond* pc,
                                  const REAL tol, const REAL abstol, const INT MaxIt,
                                  const SHORT restart, const SHORT StopType,
                                  const SHORT PrtLvl);
/* replaces fasp_solver_dbsr_pgmres   KryPgmres.c:376 */
INT fasp_cuda_solver_dbsr_pgmres(dBSRmat* A, dvector* b, dvector* x, precond* pc,
                                 const REAL tol, const REAL abstol, const INT MaxIt,
                                 const SHORT restart, const SHORT StopType,
                                 const SHORT PrtLvl);
/* replaces fasp_solver_dbsr_pvfgmres KryPvfgmres.c:386 */
INT fasp_cuda_solver_dbsr_pvfgmres(dBSRmat* A, dvector* b, dvector* x, precond* pc,
                                   const REAL tol, const REAL abstol, const INT MaxIt,
                                   const SHORT restart, const SHORT StopType,
                                   const SHORT PrtLvl);

 /* Page-lock / release an application array (b, x of repeated solves) so that the host-pointer
 * entry points copy it by DMA instead of staging it. Thin wrappers over cudaHostRegister /
 * cudaHostUnregister for callers without the CUDA headers; unpin before freeing the array.  */
INT fasp_cuda_host_pin(void* p, size_t bytes);
INT fasp_cuda_host_unpin(void* p);

/* ------------------------------------------------------------------------------------ */
/* Level-5: drivers                                                                       */
/* ------------------------------------------------------------------------------------ */

/* replaces fasp_solver_dcsr_itsolver  SolCSR.c:56 (itsolver_type CG / GMRES / VGMRES) */
INT fasp_cuda_solver_dcsr_itsolver(dCSRmat* A, dvector* b, dvector* x, precond* pc,
                                   ITS_param* itparam);
/* replaces fasp_solver_dbsr_itsolver  SolBSR.c:55 */
INT fasp_cuda_solver_dbsr_itsolver(dBSRmat* A, dvector* b, dvector* x, precond* pc,
                                   ITS_param* itparam);

/* replaces fasp_solver_dcsr_krylov_amg SolCSR.c:476 — AMG setup by the host application's
 * FASP (fasp_amg_setup_rs/sa/ua, resolved at run time from the process or from the library
 * named by $FASP_CUDA_HOST_LIBFASP), upload once, Krylov + V-cycle on device, download x.  */
INT fasp_cuda_solver_dcsr_krylov_amg(dCSRmat* A, dvector* b, dvector* x, ITS_param* itparam,
                                     AMG_param* amgparam);
/* replaces fasp_solver_dbsr_krylov_amg SolBSR.c:349 */
INT fasp_cuda_solver_dbsr_krylov_amg(dBSRmat* A, dvector* b, dvector* x, ITS_param* itparam,
                                     AMG_param* amgparam);

/* Split form for callers who solve many right-hand sides with one hierarchy (and for
 * timing the solve phase alone): setup+upload once, solve repeatedly.
 * `fasp_cuda_krylov_amg_solve` copies b (and x0) H2D and x D2H around the device solve.   */
typedef struct fasp_cuda_solver_s fasp_cuda_solver;
fasp_cuda_solver* fasp_cuda_krylov_amg_create(AMG_data* mgl, AMG_param* amgparam);
fasp_cuda_solver* fasp_cuda_krylov_bamg_create(AMG_data_bsr* mgl, AMG_param* amgparam);
INT  fasp_cuda_krylov_amg_solve(fasp_cuda_solver* s, dvector* b, dvector* x, ITS_param* itparam);
/* device-resident form: b_dev / x_dev are device vectors; returns the FASP status       */
INT  fasp_cuda_krylov_amg_solve_dev(fasp_cuda_solver* s, const REAL* b_dev, REAL* x_dev,
                                    ITS_param* itparam);
void fasp_cuda_krylov_amg_destroy(fasp_cuda_solver* s);

/* Statistics of the last solve on a solver object:
 *   what = 0 iterations, 1 final relres, 2 device ms of the Krylov loop (CUDA events),
 *          3 kernels launched, 4 ms including H2D/D2H of b and x                          */
double fasp_cuda_solver_stat(const fasp_cuda_solver* s, int what);
/* relres history of the last solve (entry 0 = initial); returns entries written */
INT fasp_cuda_solver_history(const fasp_cuda_solver* s, REAL* relres, INT max_entries);

/* AMG as a stand-alone iterative solver. replaces fasp_amg_solve PreMGSolve.c:49
 * (param->maxit cycles, stop at ||r||/||b|| < param->tol)                               */
INT fasp_cuda_amg_solve(AMG_data* mgl, AMG_param* param);

/* ------------------------------------------------------------------------------------ */
/* Multi-GPU (one process per GPU; rows partitioned)            */
/* ------------------------------------------------------------------------------------ */

/* NCCL bootstrap: rank 0 calls get_unique_id (128 bytes), the host broadcasts the bytes by
 * any means, every rank calls comm_init. Communicator is process-global.                 */
INT fasp_cuda_comm_unique_id(void* id128);
INT fasp_cuda_comm_init(const void* id128, int rank, int nranks);
INT fasp_cuda_comm_finalize(void);
int fasp_cuda_comm_rank(void);
int fasp_cuda_comm_size(void);
/* 1 when ghost exchanges and reductions go through peer-mapped memory (CUDA IPC over NVLink),
 * 0 when they fall back to NCCL send/recv + all-reduce                                       */
int fasp_cuda_comm_peer_memory(void);

/* Row-partitioned solver: every rank passes the SAME host hierarchy (FASP's deterministic setup run
 * redundantly); levels with >= agg_rows global rows are split into contiguous row slabs (rank r owns
 * rows [begin, end) of level 0, see _row_range), smaller levels are replicated. The object is used
 * with fasp_cuda_krylov_amg_solve / _solve_dev / _destroy; b and x are the rank's LOCAL slices.    */
fasp_cuda_solver* fasp_cuda_dist_krylov_amg_create(AMG_data* mgl, AMG_param* amgparam, INT agg_rows);
INT fasp_cuda_dist_row_range(const fasp_cuda_solver* s, INT* row_begin, INT* row_end);
/* The same solver from per-rank SLABS: no rank holds a global matrix, so the fine levels may exceed what one
 * dCSRmat can address (INT is 32-bit, fasp.h:72: the 27-point 512^3 system has 3.6 G nonzeros). Level l < nlev:
 *   A  this rank's rows [row_off[rank], row_off[rank+1]) of A_l, GLOBAL column numbers, col = global size;
 *   P  the same rows of P_l (global coarse columns), optionally followed by n_pext extra rows = the rows of P_l
 *      for the ghost columns of the A slab, ascending (the cycle then updates the ghosts of x_l redundantly
 *      instead of exchanging them; n_pext = 0: not supplied);
 *   R  this rank's rows of R_l = P_l^T (rows [row_off_{l+1}[rank], ...), global fine columns), optionally
 *      followed by n_rext extra rows = the rows for the ghost columns of the A_{l+1} slab.
 * tail_row_off partitions the first replicated level; `tail` is an ordinary FASP hierarchy (fasp_amg_setup_*)
 * whose level-0 matrix is that level, identical on every rank. Built by FASP's own per-level routines on each
 * slab + a distributed Galerkin product: faspsolver_b200/slabsetup.py, DESIGN.md §6.                        */
typedef struct fasp_cuda_slab_level {
    dCSRmat    A, P, R;
    const INT* row_off; /* nranks + 1 */
    INT        n_pext, n_rext;
} fasp_cuda_slab_level;
fasp_cuda_solver* fasp_cuda_dist_krylov_amg_create_slabs(INT nlev, const fasp_cuda_slab_level* levels,
                                                         const INT* tail_row_off, AMG_data* tail,
                                                         AMG_param* amgparam);
/* Host-only (no GPU, no communicator): the local slab of a square operator for `rank` of `nranks`:
 * local ia/ja (columns renumbered [owned | ghosts]), the ghosts' global columns, and per peer the
 * owned local indices that peer needs (send_idx concatenated by peer, send_counts[nranks]).
 * Returns the local nnz or ERROR_*. Used by the CPU (gloo) tests of the partition logic.       */
INT fasp_cuda_dist_extract_host(const dCSRmat* A, INT nranks, INT rank, INT* ia, INT* ja, INT* ghosts,
                                INT ghost_cap, INT* nghost, INT* send_idx, INT send_cap, INT* send_counts);

#ifdef __cplusplus
}
#endif
#endif /* FASP_CUDA_H */
