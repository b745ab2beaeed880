/* fasp_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, sequential CPU restatement of the solve-phase hot path of FASP 2.8.7 (the
 * reference, /root/reference), written for this repository. It is the parity checker of
 * libfasp_cuda when the compiled reference (oracle/_ref/libfasp_seq.so) is not at hand, and it
 * is itself pinned against that reference and against the committed golden vectors by
 * tests/test_oracle.py (FE problem of the reference's regression suite: test/out/reg.gcc).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file. The
 * product (libfasp_cuda.so) never links, loads or calls it.
 *
 * Every function cites the reference source it restates (path:line under base/src). The code
 * keeps the reference's order of floating-point operations (row sums left to right, multiply
 * and add rounded separately) because bit-level agreement with it is part of what is tested.
 *
 * Types are plain arrays: CSR = (n, ia, ja, val), 0-based.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SMALLREAL 1e-20
#define SMALLREAL2 1e-40
#define BIGREAL 1e+20
#define MAX_STAG 20
#define MAX_RESTART 20
#define STAG_RATIO 1e-4
#define ERROR_SOLVER_STAG (-42)
#define ERROR_SOLVER_SOLSTAG (-43)
#define ERROR_SOLVER_TOLSMALL (-44)
#define ERROR_SOLVER_MAXIT (-48)
#define MAXLVL 20

typedef struct {
    int           n, m; /* rows, cols */
    const int*    ia;
    const int*    ja;
    const double* val; /* NULL = all ones (UA-AMG transfer operators) */
} ocsr;

/* ---- BLAS-1: BlaArray.c:43-795 (sequential left-to-right sums) ---- */
static double o_dot(int n, const double* x, const double* y)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += x[i] * y[i]; /* BlaArray.c:771-795 */
    return s;
}
static double o_norm2(int n, const double* x) { return sqrt(o_dot(n, x, x)); } /* :691 */
static double o_norminf(int n, const double* x)
{
    double m = 0.0;
    for (int i = 0; i < n; ++i) m = fmax(m, fabs(x[i])); /* :719 */
    return m;
}
static void o_axpy(int n, double a, const double* x, double* y)
{
    for (int i = 0; i < n; ++i) y[i] += a * x[i]; /* :90 (a = +-1 branches give the same bits) */
}
static void o_axpby(int n, double a, const double* x, double b, double* y)
{
    for (int i = 0; i < n; ++i) y[i] = a * x[i] + b * y[i]; /* :620 */
}

/* ---- SpMV: BlaSpmvCSR.c:242 (mxv), :494 (aAxpy), :438/:727 (agg) ---- */
void oracle_dcsr_mxv(int n, const int* ia, const int* ja, const double* val, const double* x, double* y)
{
    for (int i = 0; i < n; ++i) {
        double t = 0.0;
        for (int k = ia[i]; k < ia[i + 1]; ++k) t += (val ? val[k] : 1.0) * x[ja[k]];
        y[i] = t;
    }
}
void oracle_dcsr_aAxpy(double alpha, int n, const int* ia, const int* ja, const double* val, const double* x,
                       double* y)
{
    for (int i = 0; i < n; ++i) {
        double t = 0.0;
        for (int k = ia[i]; k < ia[i + 1]; ++k) t += (val ? val[k] : 1.0) * x[ja[k]];
        if (alpha == 1.0) y[i] += t;
        else if (alpha == -1.0) y[i] -= t;
        else y[i] += t * alpha; /* BlaSpmvCSR.c:575-588 */
    }
}

/* ---- smoothers ---- */
/* ItrSmootherCSR.c:98-230: weighted Jacobi, simultaneous update, rows with |d| <= 1e-20 skipped */
void oracle_smoother_jacobi(int n, const int* ia, const int* ja, const double* val, const double* b, double* u,
                            int L, double w)
{
    double* t = (double*)calloc(n, sizeof(double));
    double* d = (double*)calloc(n, sizeof(double));
    while (L--) {
        for (int i = 0; i < n; ++i) {
            t[i] = b[i];
            for (int k = ia[i]; k < ia[i + 1]; ++k) {
                if (ja[k] != i) t[i] -= val[k] * u[ja[k]];
                else d[i] = val[k];
            }
        }
        for (int i = 0; i < n; ++i)
            if (fabs(d[i]) > SMALLREAL) u[i] = (1 - w) * u[i] + w * t[i] / d[i];
    }
    free(t);
    free(d);
}
/* ItrSmootherCSR.c:1509-1630: u += (b - A u) / sum_k |a_ik| */
void oracle_smoother_l1diag(int n, const int* ia, const int* ja, const double* val, const double* b, double* u,
                            int L)
{
    double* t = (double*)calloc(n, sizeof(double));
    double* d = (double*)calloc(n, sizeof(double));
    while (L--) {
        for (int i = 0; i < n; ++i) {
            t[i] = b[i];
            d[i] = 0.0;
            for (int k = ia[i]; k < ia[i + 1]; ++k) {
                t[i] -= val[k] * u[ja[k]];
                d[i] += fabs(val[k]);
            }
        }
        for (int i = 0; i < n; ++i)
            if (fabs(d[i]) > SMALLREAL) u[i] += t[i] / d[i];
    }
    free(t);
    free(d);
}
/* ItrSmootherCSRpoly.c:67-145 (driver), :392 Diaginv, :428 DinvAnorminf, :551 Rr */
void oracle_smoother_poly(int n, const int* ia, const int* ja, const double* val, const double* b, double* u,
                          int ndeg, int L)
{
    double *Dinv = calloc(n, 8), *r = calloc(n, 8), *rbar = calloc(n, 8), *v0 = calloc(n, 8),
           *v1 = calloc(n, 8), *vnew = calloc(n, 8);
    double k[6], norm = 0.0;
    for (int i = 0; i < n; ++i) {
        int j = ia[i];
        for (; j < ia[i + 1]; ++j)
            if (ja[j] == i) break;
        Dinv[i] = 1.0 / val[j];
    }
    for (int i = 0; i < n; ++i) {
        double t = 0.0;
        for (int j = ia[i]; j < ia[i + 1]; ++j) t += fabs(val[j]);
        t *= Dinv[i];
        norm = fmax(norm, t);
    }
    double mu0 = 1.0 / norm, mu1 = 4.0 * mu0, smu0 = sqrt(mu0), smu1 = sqrt(mu1);
    k[1] = (mu0 + mu1) / 2.0;
    k[2] = (smu0 + smu1) * (smu0 + smu1) / 2.0;
    k[3] = mu0 * mu1;
    k[4] = 2.0 * k[3] / k[2];
    k[5] = (mu1 - 2.0 * smu0 * smu1 + mu0) / (mu1 + 2.0 * smu0 * smu1 + mu0);
    for (int s = 0; s < L; ++s) {
        oracle_dcsr_mxv(n, ia, ja, val, u, r);
        for (int i = 0; i < n; ++i) r[i] = -1 * r[i] + b[i]; /* axpyz(n,-1,r,b,r) */
        for (int i = 0; i < n; ++i) rbar[i] = Dinv[i] * r[i];
        oracle_dcsr_mxv(n, ia, ja, val, rbar, v1);
        for (int i = 0; i < n; ++i) v1[i] = Dinv[i] * v1[i];
        for (int i = 0; i < n; ++i) {
            v0[i] = k[1] * rbar[i];
            v1[i] = k[2] * rbar[i] - k[3] * v1[i];
        }
        for (int j = 1; j < ndeg; ++j) {
            oracle_dcsr_mxv(n, ia, ja, val, v1, rbar);
            for (int i = 0; i < n; ++i) {
                rbar[i] = (r[i] - rbar[i]) * Dinv[i];
                vnew[i] = v1[i] + k[5] * (v1[i] - v0[i]) + k[4] * rbar[i];
                v0[i]   = v1[i];
                v1[i]   = vnew[i];
            }
        }
        for (int i = 0; i < n; ++i) u[i] += vnew[i]; /* axpy(n, 1, error, u) */
    }
    free(Dinv), free(r), free(rbar), free(v0), free(v1), free(vnew);
}

/* ---- BSR: BlaSpmvBSR.c:1055 (mxv), :514 (aAxpy); BlaSmallMat.c:673/:779 (block kernels);
 *      ItrSmootherBSR.c:263 (jacobi1) ---- */
static double blk_row(const double* A, const double* x, int nb)
{ /* (((A0 x0 + A1 x1) + A2 x2) + ...) of fasp_blas_smat_ypAx_nc2..7 */
    double e = A[0] * x[0];
    for (int j = 1; j < nb; ++j) e = e + A[j] * x[j];
    return e;
}
void oracle_dbsr_mxv(int ROW, int nb, const int* IA, const int* JA, const double* val, const double* x, double* y)
{
    const int nb2 = nb * nb;
    for (int i = 0; i < ROW * nb; ++i) y[i] = 0.0;
    for (int I = 0; I < ROW; ++I)
        for (int k = IA[I]; k < IA[I + 1]; ++k)
            for (int i = 0; i < nb; ++i) {
                if (nb <= 7) y[I * nb + i] += blk_row(val + (size_t)k * nb2 + i * nb, x + JA[k] * nb, nb);
                else
                    for (int j = 0; j < nb; ++j) y[I * nb + i] += val[(size_t)k * nb2 + i * nb + j] * x[JA[k] * nb + j];
            }
}
void oracle_dbsr_aAxpy(double alpha, int ROW, int nb, const int* IA, const int* JA, const double* val,
                       const double* x, double* y)
{
    const int nb2 = nb * nb, size = ROW * nb;
    if (alpha == 0.0) return;
    if (alpha != 1.0) {
        const double t = 1.0 / alpha;
        for (int i = 0; i < size; ++i) y[i] *= t;
    }
    for (int I = 0; I < ROW; ++I)
        for (int k = IA[I]; k < IA[I + 1]; ++k)
            for (int i = 0; i < nb; ++i) {
                if (nb <= 7) y[I * nb + i] += blk_row(val + (size_t)k * nb2 + i * nb, x + JA[k] * nb, nb);
                else
                    for (int j = 0; j < nb; ++j) y[I * nb + i] += val[(size_t)k * nb2 + i * nb + j] * x[JA[k] * nb + j];
            }
    if (alpha != 1.0)
        for (int i = 0; i < size; ++i) y[i] *= alpha;
}
void oracle_dbsr_jacobi1(int ROW, int nb, const int* IA, const int* JA, const double* val, const double* b,
                         double* u, const double* diaginv)
{
    const int nb2 = nb * nb, size = ROW * nb;
    double*   t = (double*)malloc(size * sizeof(double));
    memcpy(t, b, size * sizeof(double));
    for (int I = 0; I < ROW; ++I)
        for (int k = IA[I]; k < IA[I + 1]; ++k)
            if (JA[k] != I)
                for (int i = 0; i < nb; ++i) {
                    if (nb <= 7) t[I * nb + i] -= blk_row(val + (size_t)k * nb2 + i * nb, u + JA[k] * nb, nb);
                    else
                        for (int j = 0; j < nb; ++j) t[I * nb + i] -= val[(size_t)k * nb2 + i * nb + j] * u[JA[k] * nb + j];
                }
    for (int I = 0; I < ROW; ++I)
        for (int i = 0; i < nb; ++i) {
            if (nb <= 7) u[I * nb + i] = blk_row(diaginv + (size_t)I * nb2 + i * nb, t + I * nb, nb);
            else {
                double e = 0.0;
                for (int j = 0; j < nb; ++j) e += diaginv[(size_t)I * nb2 + i * nb + j] * t[I * nb + j];
                u[I * nb + i] = e;
            }
        }
    free(t);
}

/* ---- multigrid cycle: PreMGCycle.c:48-274; smoother dispatch PreMGSmoother.inl:49,155 ---- */
typedef struct {
    int     nl;
    ocsr    A[MAXLVL], P[MAXLVL], R[MAXLVL];
    double *b[MAXLVL], *x[MAXLVL], *w[MAXLVL];
    int     smoother, cycle_type, presmooth, postsmooth, ndeg, coarse_scaling;
    double  relax, tol;
} omg;

/* coarsest level: unpreconditioned CG to a relative residual of ctol (fasp_coarse_itsolver,
 * PreMGUtil.inl:37-58 calls the safeguarded fasp_solver_dcsr_spcg, KrySPcg.c:60; the plain
 * recurrence below takes the same iterates until the tolerance is met) */
static void o_coarse_cg(const ocsr* A, const double* b, double* x, double ctol)
{
    const int n = A->n;
    int maxit = n * n < 1000 ? n * n : 1000;
    if (maxit < 250) maxit = 250;
    double *r = malloc(n * 8), *p = malloc(n * 8), *t = malloc(n * 8);
    memcpy(r, b, n * 8);
    oracle_dcsr_aAxpy(-1.0, n, A->ia, A->ja, A->val, x, r);
    const double normr0 = fmax(SMALLREAL, o_norm2(n, r));
    memcpy(p, r, n * 8);
    double rr = o_dot(n, r, r);
    for (int it = 0; it < maxit && sqrt(rr) / normr0 >= ctol; ++it) {
        oracle_dcsr_mxv(n, A->ia, A->ja, A->val, p, t);
        const double tp = o_dot(n, t, p);
        if (fabs(tp) <= SMALLREAL2) break;
        const double alpha = rr / tp;
        o_axpy(n, alpha, p, x);
        o_axpy(n, -alpha, t, r);
        const double rr1 = o_dot(n, r, r);
        o_axpby(n, 1.0, r, rr1 / rr, p);
        rr = rr1;
    }
    free(r), free(p), free(t);
}

void oracle_mgcycle(omg* g)
{
    const int nl = g->nl;
    int num_lvl[MAXLVL] = {0}, ncycles[MAXLVL], l = 0;
    for (int i = 0; i < MAXLVL; ++i) ncycles[i] = 1;
    if (g->cycle_type == 12) { for (int i = MAXLVL - 2; i > 0; i -= 2) ncycles[i] = 2; }
    else if (g->cycle_type == 21) { for (int i = MAXLVL - 1; i > 0; i -= 2) ncycles[i] = 2; }
    else for (int i = 0; i < MAXLVL; ++i) ncycles[i] = g->cycle_type;
#define SMOOTH(l, L)                                                                              \
    do {                                                                                          \
        const ocsr* A_ = &g->A[l];                                                                \
        if (g->smoother == 1) oracle_smoother_jacobi(A_->n, A_->ia, A_->ja, A_->val, g->b[l], g->x[l], L, g->relax); \
        else if (g->smoother == 10) oracle_smoother_l1diag(A_->n, A_->ia, A_->ja, A_->val, g->b[l], g->x[l], L);    \
        else oracle_smoother_poly(A_->n, A_->ia, A_->ja, A_->val, g->b[l], g->x[l], g->ndeg, L);  \
    } while (0)
    for (;;) {
        while (l < nl - 1) { /* ForwardSweep, :96-152 */
            num_lvl[l]++;
            SMOOTH(l, g->presmooth);
            memcpy(g->w[l], g->b[l], g->A[l].n * 8);
            oracle_dcsr_aAxpy(-1.0, g->A[l].n, g->A[l].ia, g->A[l].ja, g->A[l].val, g->x[l], g->w[l]);
            oracle_dcsr_mxv(g->R[l].n, g->R[l].ia, g->R[l].ja, g->R[l].val, g->w[l], g->b[l + 1]);
            ++l;
            memset(g->x[l], 0, g->A[l].n * 8);
        }
        o_coarse_cg(&g->A[nl - 1], g->b[nl - 1], g->x[nl - 1], g->tol * 1e-4); /* :58, :199-201 */
        while (l > 0) { /* BackwardSweep, :205-266 */
            --l;
            double alpha = 1.0;
            if (g->coarse_scaling) {
                const ocsr* Ac = &g->A[l + 1];
                double*     t  = malloc(Ac->n * 8);
                oracle_dcsr_mxv(Ac->n, Ac->ia, Ac->ja, Ac->val, g->x[l + 1], t);
                alpha = o_dot(Ac->n, g->x[l + 1], g->b[l + 1]) / o_dot(Ac->n, g->x[l + 1], t); /* vmv :839 */
                if (alpha > 1.0) alpha = 1.0;
                free(t);
            }
            oracle_dcsr_aAxpy(alpha, g->P[l].n, g->P[l].ia, g->P[l].ja, g->P[l].val, g->x[l + 1], g->x[l]);
            SMOOTH(l, g->postsmooth);
            if (num_lvl[l] < ncycles[l]) break;
            num_lvl[l] = 0;
        }
        if (l == 0) break;
    }
#undef SMOOTH
}

/* handle-based wrappers so that ctypes can drive the cycle */
omg* oracle_mg_new(int nl, int smoother, int cycle_type, int presmooth, int postsmooth, int ndeg,
                   int coarse_scaling, double relax, double tol)
{
    omg* g = (omg*)calloc(1, sizeof(omg));
    g->nl = nl, g->smoother = smoother, g->cycle_type = cycle_type, g->presmooth = presmooth;
    g->postsmooth = postsmooth, g->ndeg = ndeg, g->coarse_scaling = coarse_scaling, g->relax = relax, g->tol = tol;
    return g;
}
void oracle_mg_set_level(omg* g, int l, int n, const int* ia, const int* ja, const double* val, int pn, int pm,
                         const int* pia, const int* pja, const double* pval, int rn, int rm, const int* ria,
                         const int* rja, const double* rval)
{
    g->A[l] = (ocsr){n, n, ia, ja, val};
    g->P[l] = (ocsr){pn, pm, pia, pja, pval};
    g->R[l] = (ocsr){rn, rm, ria, rja, rval};
    g->b[l] = calloc(n, 8), g->x[l] = calloc(n, 8), g->w[l] = calloc(n, 8);
}
void oracle_mg_free(omg* g)
{
    for (int l = 0; l < g->nl; ++l) free(g->b[l]), free(g->x[l]), free(g->w[l]);
    free(g);
}
/* z = B r: x0 = 0, `maxit` cycles (fasp_precond_amg, PreCSR.c:416-435) */
void oracle_precond_amg(omg* g, const double* r, double* z, int maxit)
{
    const int n = g->A[0].n;
    memcpy(g->b[0], r, n * 8);
    memset(g->x[0], 0, n * 8);
    for (int i = 0; i < maxit; ++i) oracle_mgcycle(g);
    memcpy(z, g->x[0], n * 8);
}
/* AMG as a solver: cycles until ||b - A x|| / ||b|| < tol (fasp_amg_solve, PreMGSolve.c:49-122) */
int oracle_amg_solve(omg* g, const double* b, double* x, double tol, int MaxIt, double* relres_out)
{
    const int    n    = g->A[0].n;
    const double sumb = o_norm2(n, b);
    double       relres1 = 1.0;
    int          iter = 0;
    memcpy(g->b[0], b, n * 8);
    memcpy(g->x[0], x, n * 8);
    if (sumb <= SMALLREAL) memset(g->x[0], 0, n * 8);
    while ((iter++ < MaxIt) & (sumb > SMALLREAL)) {
        oracle_mgcycle(g);
        memcpy(g->w[0], g->b[0], n * 8);
        oracle_dcsr_aAxpy(-1.0, n, g->A[0].ia, g->A[0].ja, g->A[0].val, g->x[0], g->w[0]);
        relres1 = o_norm2(n, g->w[0]) / fmax(SMALLREAL, sumb);
        if (relres1 < tol) break;
    }
    memcpy(x, g->x[0], n * 8);
    if (relres_out) *relres_out = relres1;
    return iter > MaxIt ? ERROR_SOLVER_MAXIT : iter;
}
void oracle_mg_cycle_on(omg* g, const double* b, double* x)
{
    const int n = g->A[0].n;
    memcpy(g->b[0], b, n * 8);
    memcpy(g->x[0], x, n * 8);
    oracle_mgcycle(g);
    memcpy(x, g->x[0], n * 8);
}

/* ---- PCG: KryPcg.c:96-362 (STOP_REL_RES), preconditioner = AMG handle or identity ---- */
int oracle_pcg(int n, const int* ia, const int* ja, const double* val, const double* b, double* u, omg* pc,
               double tol, double abstol, int MaxIt, double* relres_out)
{
    const double maxdiff = tol * STAG_RATIO;
    int    iter = 0, stag = 1, more_step = 1;
    double absres0, absres = BIGREAL, relres, normr0, factor, alpha, beta, temp1, temp2;
    double *p = calloc(n, 8), *z = calloc(n, 8), *r = calloc(n, 8), *t = calloc(n, 8);
#define PREC(r_, z_) do { if (pc) oracle_precond_amg(pc, r_, z_, 1); else memcpy(z_, r_, n * 8); } while (0)
#define RESID() do { memcpy(r, b, n * 8); oracle_dcsr_aAxpy(-1.0, n, ia, ja, val, u, r); } while (0)
    RESID();
    PREC(r, z);
    absres0 = o_norm2(n, r);
    normr0  = fmax(SMALLREAL, absres0);
    relres  = absres0 / normr0;
    if (relres < tol || absres0 < abstol) goto FINISHED;
    memcpy(p, z, n * 8);
    temp1 = o_dot(n, z, r);
    while (iter++ < MaxIt) {
        oracle_dcsr_mxv(n, ia, ja, val, p, t);
        temp2 = o_dot(n, t, p);
        if (fabs(temp2) > SMALLREAL2) alpha = temp1 / temp2;
        else goto FINISHED;
        o_axpy(n, alpha, p, u);
        o_axpy(n, -alpha, t, r);
        absres = o_norm2(n, r);
        relres = absres / normr0;
        factor = absres / absres0;
        if (factor > 0.9) { /* :212-274 */
            if (o_norminf(n, u) <= SMALLREAL) { iter = ERROR_SOLVER_SOLSTAG; break; }
            const double normu = o_norm2(n, u), reldiff = fabs(alpha) * o_norm2(n, p) / normu;
            if ((stag <= MAX_STAG) & (reldiff < maxdiff)) {
                RESID();
                absres = o_norm2(n, r);
                relres = absres / normr0;
                if (relres < tol) break;
                if (stag >= MAX_STAG) { iter = ERROR_SOLVER_STAG; break; }
                memset(p, 0, n * 8);
                ++stag;
            }
        }
        if (relres < tol) { /* :277-324 */
            RESID();
            absres = o_norm2(n, r);
            relres = absres / normr0;
            if (relres < tol) break;
            if (more_step >= MAX_RESTART) { iter = ERROR_SOLVER_TOLSMALL; break; }
            memset(p, 0, n * 8);
            ++more_step;
        }
        absres0 = absres;
        PREC(r, z);
        temp2 = o_dot(n, z, r);
        beta  = temp2 / temp1;
        temp1 = temp2;
        o_axpby(n, 1.0, z, beta, p);
    }
FINISHED:
#undef PREC
#undef RESID
    if (relres_out) *relres_out = relres;
    free(p), free(z), free(r), free(t);
    return iter > MaxIt ? ERROR_SOLVER_MAXIT : iter;
}

/* ---- restarted GMRES, right preconditioning: KryPvgmres.c:66-387 (variable = 1) and
 *      KryPgmres.c:66 (variable = 0); STOP_REL_RES ---- */
int oracle_gmres(int n, const int* ia, const int* ja, const double* val, const double* b, double* x, omg* pc,
                 double tol, double abstol, int MaxIt, int restart, int variable, double* relres_out)
{
    const double cr_max = 0.99, cr_min = 0.174;
    int    iter = 0, i, j, k, d = 3, restart_max = restart, restart_min = 3;
    int    Restart = variable ? restart : (restart < MaxIt ? restart : MaxIt);
    const int R1 = Restart + 1;
    double r_norm, gamma, t, absres0, relres = BIGREAL, cr = 1.0, r_norm_old = 0.0;
    double *r = calloc(n, 8), *w = calloc(n, 8), *rs = calloc(R1 + 1, 8), *c = calloc(R1, 8), *s = calloc(R1, 8);
    double** p  = malloc(R1 * sizeof(double*));
    double** hh = malloc(R1 * sizeof(double*));
    for (i = 0; i < R1; ++i) p[i] = calloc(n, 8), hh[i] = calloc(R1, 8);
#define PREC(r_, z_) do { if (pc) oracle_precond_amg(pc, r_, z_, 1); else memcpy(z_, r_, n * 8); } while (0)
    memcpy(p[0], b, n * 8);
    oracle_dcsr_aAxpy(-1.0, n, ia, ja, val, x, p[0]);
    r_norm  = o_norm2(n, p[0]);
    absres0 = fmax(SMALLREAL, r_norm);
    relres  = r_norm / absres0;
    if (relres < tol || absres0 < abstol) goto FINISHED;
    while (iter < MaxIt && (variable || relres > tol)) {
        rs[0] = r_norm_old = r_norm;
        t = 1.0 / r_norm;
        for (k = 0; k < n; ++k) p[0][k] *= t;
        if (variable) {
            if (cr > cr_max || iter == 0) Restart = restart_max;
            else if (cr < cr_min) { }
            else if (Restart - d > restart_min) Restart -= d;
            else Restart = restart_max;
        }
        i = 0;
        while (i < Restart && iter < MaxIt) {
            i++, iter++;
            PREC(p[i - 1], r);
            oracle_dcsr_mxv(n, ia, ja, val, r, p[i]);
            for (j = 0; j < i; j++) {
                hh[j][i - 1] = o_dot(n, p[j], p[i]);
                o_axpy(n, -hh[j][i - 1], p[j], p[i]);
            }
            t = o_norm2(n, p[i]);
            hh[i][i - 1] = t;
            if (variable ? (t != 0.0) : (fabs(t) > SMALLREAL)) {
                t = 1.0 / t;
                for (k = 0; k < n; ++k) p[i][k] *= t;
            }
            for (j = 1; j < i; ++j) {
                t = hh[j - 1][i - 1];
                hh[j - 1][i - 1] = s[j - 1] * hh[j][i - 1] + c[j - 1] * t;
                hh[j][i - 1]     = -s[j - 1] * t + c[j - 1] * hh[j][i - 1];
            }
            t = hh[i][i - 1] * hh[i][i - 1];
            t += hh[i - 1][i - 1] * hh[i - 1][i - 1];
            gamma = sqrt(t);
            if (variable) { if (gamma == 0.0) gamma = SMALLREAL; }
            else gamma = fmax(gamma, SMALLREAL);
            c[i - 1] = hh[i - 1][i - 1] / gamma;
            s[i - 1] = hh[i][i - 1] / gamma;
            rs[i]     = -s[i - 1] * rs[i - 1];
            rs[i - 1] = c[i - 1] * rs[i - 1];
            hh[i - 1][i - 1] = s[i - 1] * hh[i][i - 1] + c[i - 1] * hh[i - 1][i - 1];
            r_norm = fabs(rs[i]);
            relres = r_norm / absres0;
            if (relres < tol) break;
        }
        rs[i - 1] = rs[i - 1] / hh[i - 1][i - 1];
        for (k = i - 2; k >= 0; k--) {
            t = 0.0;
            for (j = k + 1; j < i; j++) t -= hh[k][j] * rs[j];
            t += rs[k];
            rs[k] = t / hh[k][k];
        }
        memcpy(w, p[i - 1], n * 8);
        for (k = 0; k < n; ++k) w[k] *= rs[i - 1];
        for (j = i - 2; j >= 0; j--) o_axpy(n, rs[j], p[j], w);
        PREC(w, r);
        o_axpy(n, 1.0, r, x);
        if (relres < tol) {
            memcpy(r, b, n * 8);
            oracle_dcsr_aAxpy(-1.0, n, ia, ja, val, x, r);
            r_norm = o_norm2(n, r);
            relres = r_norm / absres0;
            if (relres < tol) break;
            memcpy(p[0], r, n * 8);
            i = 0;
        }
        for (j = i; j > 0; j--) {
            rs[j - 1] = -s[j - 1] * rs[j];
            rs[j]     = c[j - 1] * rs[j];
        }
        if (i) o_axpy(n, rs[i] - 1.0, p[i], p[i]);
        for (j = i - 1; j > 0; j--) o_axpy(n, rs[j], p[j], p[i]);
        if (i) {
            o_axpy(n, rs[0] - 1.0, p[0], p[0]);
            o_axpy(n, 1.0, p[i], p[0]);
        }
        cr = r_norm / r_norm_old;
    }
FINISHED:
#undef PREC
    if (relres_out) *relres_out = relres;
    for (i = 0; i < R1; ++i) free(p[i]), free(hh[i]);
    free(p), free(hh), free(r), free(w), free(rs), free(c), free(s);
    return iter >= MaxIt ? ERROR_SOLVER_MAXIT : iter;
}

/* ---- flexible variable-restart GMRES: KryPvfgmres.c:67-358. Differences from oracle_gmres:
 *      z_j = B p_j is kept per step (:226-231) and the update is x += sum rs_j z_j (:286-291);
 *      the inner exit is r_norm <= tol * den_norm with den_norm = ||b|| (or ||r0|| if b = 0)
 *      (:153-158, :273); the true-residual recheck uses r_norm / den_norm <= tol (:294-317).
 *      relres_out receives r_norm / den_norm as printed by ITS_FINAL (:358). ---- */
int oracle_fgmres(int n, const int* ia, const int* ja, const double* val, const double* b, double* x, omg* pc,
                  double tol, double abstol, int MaxIt, int restart, double* relres_out)
{
    const double cr_max = 0.99, cr_min = 0.174;
    int    iter = 0, i, j, k, d = 3, restart_max = restart, restart_min = 3, Restart = restart;
    const int R1 = restart + 1;
    double r_norm, b_norm, den_norm, epsilon, gamma, t, cr = 1.0, r_norm_old = 0.0;
    double *r = calloc(n, 8), *rs = calloc(R1 + 1, 8), *c = calloc(R1, 8), *s = calloc(R1, 8);
    double** p  = malloc(R1 * sizeof(double*));
    double** z  = malloc(R1 * sizeof(double*));
    double** hh = malloc(R1 * sizeof(double*));
    for (i = 0; i < R1; ++i) p[i] = calloc(n, 8), z[i] = calloc(n, 8), hh[i] = calloc(R1, 8);
    memcpy(p[0], b, n * 8);
    oracle_dcsr_aAxpy(-1.0, n, ia, ja, val, x, p[0]);
    b_norm   = o_norm2(n, b);
    r_norm   = o_norm2(n, p[0]);
    den_norm = b_norm > 0.0 ? b_norm : r_norm;
    epsilon  = tol * den_norm;
    if (r_norm < epsilon || r_norm < abstol) goto FINISHED;
    while (iter < MaxIt) {
        rs[0] = r_norm_old = r_norm;
        if (r_norm == 0.0) break;
        if (cr > cr_max || iter == 0) Restart = restart_max;
        else if (cr < cr_min) { }
        else if (Restart - d > restart_min) Restart -= d;
        else Restart = restart_max;
        t = 1.0 / r_norm;
        for (k = 0; k < n; ++k) p[0][k] *= t;
        i = 0;
        while (i < Restart && iter < MaxIt) {
            i++, iter++;
            if (pc) oracle_precond_amg(pc, p[i - 1], z[i - 1], 1);
            else memcpy(z[i - 1], p[i - 1], n * 8);
            oracle_dcsr_mxv(n, ia, ja, val, z[i - 1], p[i]);
            for (j = 0; j < i; j++) {
                hh[j][i - 1] = o_dot(n, p[j], p[i]);
                o_axpy(n, -hh[j][i - 1], p[j], p[i]);
            }
            t = o_norm2(n, p[i]);
            hh[i][i - 1] = t;
            if (t != 0.0) {
                t = 1.0 / t;
                for (k = 0; k < n; ++k) p[i][k] *= t;
            }
            for (j = 1; j < i; ++j) {
                t = hh[j - 1][i - 1];
                hh[j - 1][i - 1] = s[j - 1] * hh[j][i - 1] + c[j - 1] * t;
                hh[j][i - 1]     = -s[j - 1] * t + c[j - 1] * hh[j][i - 1];
            }
            t = hh[i][i - 1] * hh[i][i - 1];
            t += hh[i - 1][i - 1] * hh[i - 1][i - 1];
            gamma = sqrt(t);
            if (gamma == 0.0) gamma = SMALLREAL;
            c[i - 1] = hh[i - 1][i - 1] / gamma;
            s[i - 1] = hh[i][i - 1] / gamma;
            rs[i]     = -s[i - 1] * rs[i - 1];
            rs[i - 1] = c[i - 1] * rs[i - 1];
            hh[i - 1][i - 1] = s[i - 1] * hh[i][i - 1] + c[i - 1] * hh[i - 1][i - 1];
            r_norm = fabs(rs[i]);
            if (r_norm <= epsilon) break;
        }
        rs[i - 1] = rs[i - 1] / hh[i - 1][i - 1];
        for (k = i - 2; k >= 0; k--) {
            t = 0.0;
            for (j = k + 1; j < i; j++) t -= hh[k][j] * rs[j];
            t += rs[k];
            rs[k] = t / hh[k][k];
        }
        memcpy(r, z[i - 1], n * 8);
        if (rs[i - 1] != 1.0) for (k = 0; k < n; ++k) r[k] *= rs[i - 1];
        for (j = i - 2; j >= 0; j--) o_axpy(n, rs[j], z[j], r);
        o_axpy(n, 1.0, r, x);
        if (r_norm <= epsilon) {
            memcpy(r, b, n * 8);
            oracle_dcsr_aAxpy(-1.0, n, ia, ja, val, x, r);
            r_norm = o_norm2(n, r);
            if (r_norm / den_norm <= tol) break;
            memcpy(p[0], r, n * 8);
            i = 0;
        }
        for (j = i; j > 0; j--) {
            rs[j - 1] = -s[j - 1] * rs[j];
            rs[j]     = c[j - 1] * rs[j];
        }
        if (i) o_axpy(n, rs[i] - 1.0, p[i], p[i]);
        for (j = i - 1; j > 0; j--) o_axpy(n, rs[j], p[j], p[i]);
        if (i) {
            o_axpy(n, rs[0] - 1.0, p[0], p[0]);
            o_axpy(n, 1.0, p[i], p[0]);
        }
        cr = r_norm / r_norm_old;
    }
FINISHED:
    if (relres_out) *relres_out = den_norm > 0.0 ? r_norm / den_norm : 0.0;
    for (i = 0; i < R1; ++i) free(p[i]), free(z[i]), free(hh[i]);
    free(p), free(z), free(hh), free(r), free(rs), free(c), free(s);
    return iter >= MaxIt ? ERROR_SOLVER_MAXIT : iter;
}

/* ---- multicolour Gauss-Seidel (the smoother of the reference's OpenMP build) ----
 * Colouring: dCSRmat_Multicoloring, BlaSparseCSR.c:1687-1770. Rows circulate in a queue; the row at
 * the front opens a new colour if the queue has wrapped (its index is not larger than the previous
 * one), is pushed back if a row of the current colour references it, and joins the current colour
 * otherwise. IC[c] .. IC[c+1] delimit colour c inside ICMAP. Returns the number of colours. */
int oracle_multicolor(int n, const int* ia, const int* ja, int* IC, int* ICMAP)
{
    if (n <= 0) { IC[0] = 0; return 0; }
    int* q    = malloc(sizeof(int) * (n + 1));
    int* mark = malloc(sizeof(int) * (n + 1));
    for (int k = 0; k < n; ++k) q[k] = k, mark[k] = -1;
    int head = n - 1, tail = n - 1, ncol = 0, filled = 0, last = 0;
    IC[0] = 0;
    do {
        head = (head + 1 == n) ? 0 : head + 1;
        const int i = q[head];
        if (i <= last || mark[i] != ncol) {
            if (i <= last) IC[ncol++] = filled;       /* wrapped: row i starts colour ncol */
            ICMAP[filled++] = i;
            for (int k = ia[i]; k < ia[i + 1]; ++k) mark[ja[k]] = ncol;
        } else {                                      /* coupled to the colour being built */
            tail = (tail + 1 == n) ? 0 : tail + 1;
            q[tail] = i;
        }
        last = i;
    } while (tail != head);
    IC[ncol] = filled;
    free(q), free(mark);
    return ncol;
}

/* L sweeps, colours ascending (order != -1) or descending (order == -1); inside a colour the rows
 * are independent: u_i = (b_i - sum_{j != i} a_ij u_j) / a_ii, d = the last stored diagonal entry
 * (fasp_smoother_dcsr_gs_multicolor, BlaSparseCSR.c:2123-2190) */
void oracle_gs_multicolor(int n, const int* ia, const int* ja, const double* val, const double* b, double* u,
                          int L, int order, int ncol, const int* IC, const int* ICMAP)
{
    (void)n;
    double d = 0.0; /* as in the reference, d carries over when a row stores no diagonal */
    while (L--)
        for (int cc = 0; cc < ncol; ++cc) {
            const int c = (order == -1) ? ncol - 1 - cc : cc;
            for (int I = IC[c]; I < IC[c + 1]; ++I) {
                const int i = ICMAP[I];
                double    t = b[i];
                for (int k = ia[i]; k < ia[i + 1]; ++k) {
                    if (ja[k] != i) t -= val[k] * u[ja[k]];
                    else d = val[k];
                }
                if (fabs(d) > SMALLREAL) u[i] = t / d;
            }
        }
}
