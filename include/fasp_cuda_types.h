/*
 * fasp_cuda_types.h — layout-identical mirror of the FASP 2.8.7 structs that cross the
 * libfasp_cuda boundary, for translation units that do not include FASP's own fasp.h.
 *
 * This is an ABI description, not FASP source: member names/order/types follow
 * base/include/fasp.h and base/include/fasp_block.h (citations per struct) because the
 * library receives pointers to objects allocated by the host application's FASP.
 * tests/test_abi.py compiles this header and the real fasp.h side by side and compares
 * sizeof/offsetof of every member used by the library.
 *
 * Default = sequential FASP ABI. Define FASP_CUDA_OPENMP_ABI for the ABI of a FASP built
 * with -fopenmp (MULTI_COLOR_ORDER ON: fasp.h:42-50).
 */
#ifndef FASP_CUDA_TYPES_H
#define FASP_CUDA_TYPES_H

#ifdef __FASP_HEADER__
#error "include fasp_cuda.h (not fasp_cuda_types.h) after fasp.h"
#endif

/* scalar types (fasp.h:69-77) */
#ifndef SHORT
#define SHORT short
#endif
#ifndef INT
#define INT int
#endif
#ifndef REAL
#define REAL double
#endif

#ifdef FASP_CUDA_OPENMP_ABI
#define FASP_CUDA_MULTI_COLOR_ORDER 1
#else
#define FASP_CUDA_MULTI_COLOR_ORDER 0
#endif

/* ---- status codes used by the library (fasp_const.h:19-56) ---- */
#ifndef __FASP_CONST__
#define FASP_SUCCESS 0
#define ERROR_INPUT_PAR (-13)
#define ERROR_MAT_SIZE (-15)
#define ERROR_MISC (-19)
#define ERROR_ALLOC_MEM (-20)
#define ERROR_DATA_STRUCTURE (-21)
#define ERROR_DATA_ZERODIAG (-22)
#define ERROR_AMG_SMOOTH_TYPE (-31)
#define ERROR_AMG_SETUP (-39)
#define ERROR_SOLVER_TYPE (-40)
#define ERROR_SOLVER_PRECTYPE (-41)
#define ERROR_SOLVER_STAG (-42)
#define ERROR_SOLVER_SOLSTAG (-43)
#define ERROR_SOLVER_TOLSMALL (-44)
#define ERROR_SOLVER_MISC (-46)
#define ERROR_SOLVER_MAXIT (-48)
#define ERROR_SOLVER_EXIT (-49)
#define ERROR_UNKNOWN (-99)
/* print levels (fasp_const.h:73-78) */
#define PRINT_NONE 0
#define PRINT_MIN 1
#define PRINT_SOME 2
#define PRINT_MORE 4
#define PRINT_MOST 8
#define PRINT_ALL 10
/* iterative solvers (fasp_const.h:101-110) */
#define SOLVER_DEFAULT 0
#define SOLVER_CG 1
#define SOLVER_GMRES 4
#define SOLVER_VGMRES 5
#define SOLVER_VFGMRES 6
/* stopping criteria (fasp_const.h:132-134) */
#define STOP_REL_RES 1
#define STOP_REL_PRECRES 2
#define STOP_MOD_REL_RES 3
/* preconditioners (fasp_const.h:139-144) */
#define PREC_NULL 0
#define PREC_DIAG 1
#define PREC_AMG 2
/* AMG flavours, cycles, smoothers (fasp_const.h:163-200) */
#define CLASSIC_AMG 1
#define SA_AMG 2
#define UA_AMG 3
#define V_CYCLE 1
#define W_CYCLE 2
#define AMLI_CYCLE 3
#define NL_AMLI_CYCLE 4
#define VW_CYCLE 12
#define WV_CYCLE 21
#define SMOOTHER_JACOBI 1
#define SMOOTHER_GS 2
#define SMOOTHER_SGS 3
#define SMOOTHER_POLY 9
#define SMOOTHER_L1DIAG 10
#define NO_ORDER 0
#define CF_ORDER 1
/* numerical guards (fasp_const.h:255-265) */
#define BIGREAL 1e+20
#define SMALLREAL 1e-20
#define SMALLREAL2 1e-40
#define MAX_AMG_LVL 20
#define MAX_RESTART 20
#define MAX_STAG 20
#define STAG_RATIO 1e-4
#define ON 1
#define OFF 0
#endif /* __FASP_CONST__ */

/* ---- matrices and vectors ---- */

/* CSR matrix, 0-based (fasp.h:151-180) */
typedef struct dCSRmat {
    INT   row, col, nnz;
    INT*  IA;  /* row+1 offsets   */
    INT*  JA;  /* nnz column ids  */
    REAL* val; /* nnz entries     */
#if FASP_CUDA_MULTI_COLOR_ORDER
    INT  color;
    INT* IC;
    INT* ICMAP;
#endif
} dCSRmat;

/* REAL / INT vectors (fasp.h:354-362, 368-376) */
typedef struct dvector {
    INT   row;
    REAL* val;
} dvector;
typedef struct ivector {
    INT  row;
    INT* val;
} ivector;

/* block CSR matrix, nb x nb row-major blocks (fasp_block.h:34-66) */
typedef struct dBSRmat {
    INT   ROW, COL, NNZ;
    INT   nb;
    INT   storage_manner;
    REAL* val;
    INT*  IA;
    INT*  JA;
} dBSRmat;

/* ---- parameters ---- */

/* iterative solver parameters (fasp.h:386-398) */
typedef struct {
    SHORT print_level, itsolver_type, decoup_type, precond_type, stop_type;
    INT   restart, maxit;
    REAL  tol, abstol;
} ITS_param;

/* ILU / Schwarz parameters: only passed through (fasp.h:404-424, 430-447) */
typedef struct {
    SHORT print_level, ILU_type;
    INT   ILU_lfil;
    REAL  ILU_droptol, ILU_relax, ILU_permtol;
} ILU_param;
typedef struct {
    SHORT print_level, SWZ_type;
    INT   SWZ_maxlvl, SWZ_mmsize, SWZ_blksolver;
} SWZ_param;

/* AMG parameters (fasp.h:455-595) */
typedef struct {
    SHORT AMG_type, print_level;
    INT   maxit;
    REAL  tol;
    SHORT max_levels;
    INT   coarse_dof;
    SHORT cycle_type;
    REAL  quality_bound;
    SHORT smoother, smooth_order, presmooth_iter, postsmooth_iter;
    REAL  relaxation;
    SHORT polynomial_degree, coarse_solver, coarse_scaling, amli_degree;
    REAL* amli_coef;
    SHORT nl_amli_krylov_type, coarsening_type, aggregation_type, aggregation_norm_type,
        interpolation_type;
    REAL strong_threshold, max_row_sum, truncation_threshold;
    INT  aggressive_level, aggressive_path, pair_number;
    REAL strong_coupled;
    INT  max_aggregation;
    REAL tentative_smooth;
    SHORT smooth_filter, smooth_restriction, ILU_levels, ILU_type;
    INT   ILU_lfil;
    REAL  ILU_droptol, ILU_relax, ILU_permtol;
    INT   SWZ_levels, SWZ_mmsize, SWZ_maxlvl, SWZ_type, SWZ_blksolver;
    REAL  theta;
} AMG_param;

/* ---- per-level hierarchy data (host side, produced by FASP's setup) ---- */

/* third-party solver handles embedded in AMG_data; all WITH_* switches are 0 in the builds
 * this library pairs with (fasp.h:609-636) */
typedef struct {
    INT job;
} Mumps_data;
typedef struct {
    void* pt[64];
} Pardiso_data;

/* ILU factors (fasp.h:642-706): never touched, needed for layout only */
typedef struct {
    dCSRmat* A;
    INT      type, row, col, nzlu;
    INT*     ijlu;
    REAL*    luval;
    INT      nb, nwork;
    REAL*    work;
    INT*     iperm;
    INT      ncolors;
    INT *    ic, *icmap, *uptr;
    INT      nlevL, nlevU;
    INT *    ilevL, *ilevU, *jlevL, *jlevU;
} ILU_data;

/* Schwarz data (fasp.h:714-798): never touched, needed for layout only */
typedef struct {
    dCSRmat     A;
    INT         nblk;
    INT *       iblock, *jblock;
    REAL*       rhsloc;
    dvector     rhsloc1, xloc1;
    REAL *      au, *al;
    INT         SWZ_type, blk_solver, memt;
    INT*        mask;
    INT         maxbs;
    INT*        maxa;
    dCSRmat*    blk_data;
    Mumps_data* mumps;
    SWZ_param*  swzparam;
} SWZ_data;

/* one AMG level in CSR format (fasp.h:804-888); mgl[0].num_levels = levels in use */
typedef struct {
    SHORT        max_levels, num_levels;
    dCSRmat      A, R, P;
    dvector      b, x;
    void*        Numeric;
    Pardiso_data pdata;
    ivector      cfmark;
    INT          ILU_levels;
    ILU_data     LU;
    INT          near_kernel_dim;
    REAL**       near_kernel_basis;
    INT          SWZ_levels;
    SWZ_data     Schwarz;
    dvector      w;
    Mumps_data   mumps;
    INT          cycle_type;
    INT *        ic, *icmap;
    INT          colors;
    REAL         weight;
#if FASP_CUDA_MULTI_COLOR_ORDER
    REAL GS_Theta;
#endif
} AMG_data;

/* one AMG level in BSR format (fasp_block.h:146-247) */
typedef struct {
    INT          max_levels, num_levels;
    dBSRmat      A, R, P;
    dvector      b, x, diaginv;
    dCSRmat      Ac;
    void*        Numeric;
    Pardiso_data pdata;
    dCSRmat      PP;
    AMG_data*    mglP;
    dCSRmat      TT;
    AMG_data*    mglT;
    dBSRmat      PT;
    REAL*        pw;
    dBSRmat      SS;
    REAL*        sw;
    dvector      diaginv_SS;
    ILU_data     PP_LU;
    ivector      cfmark;
    INT          ILU_levels;
    ILU_data     LU;
    INT          near_kernel_dim;
    REAL**       near_kernel_basis;
    dCSRmat *    A_nk, *P_nk, *R_nk;
    dvector      w;
    Mumps_data   mumps;
} AMG_data_bsr;

/* data behind precond.data for fasp_precond_amg (fasp.h:894-981) */
typedef struct {
    SHORT AMG_type, print_level;
    INT   maxit;
    SHORT max_levels;
    REAL  tol;
    SHORT cycle_type, smoother, smooth_order, presmooth_iter, postsmooth_iter;
    REAL  relaxation;
    SHORT polynomial_degree, coarsening_type, coarse_solver, coarse_scaling, amli_degree,
        nl_amli_krylov_type;
    REAL      tentative_smooth;
    REAL*     amli_coef;
    AMG_data* mgl_data;
    ILU_data* LU;
    dCSRmat * A, *A_nk, *P_nk, *R_nk;
    dvector   r;
    REAL*     w;
} precond_data;

/* BSR twin (fasp_block.h:271-356) */
typedef struct {
    SHORT AMG_type, print_level;
    INT   maxit, max_levels;
    REAL  tol;
    SHORT cycle_type, smoother, smooth_order, presmooth_iter, postsmooth_iter,
        coarsening_type;
    REAL  relaxation;
    SHORT coarse_solver, coarse_scaling, amli_degree;
    REAL* amli_coef;
    REAL  tentative_smooth;
    SHORT nl_amli_krylov_type;
    AMG_data_bsr* mgl_data;
    AMG_data*     pres_mgl_data;
    ILU_data*     LU;
    dBSRmat*      A;
    dCSRmat *     A_nk, *P_nk, *R_nk;
    dvector       r;
    REAL*         w;
} precond_data_bsr;

/* preconditioner plug-in: z = fct(r) (fasp.h:1095-1103) */
typedef struct {
    void* data;
    void (*fct)(REAL*, REAL*, void*);
} precond;

/* matrix-free operator plug-in (fasp.h:1109-1117) */
typedef struct {
    void* data;
    void (*fct)(const void*, const REAL*, REAL*);
} mxv_matfree;

#endif /* FASP_CUDA_TYPES_H */
