/* poisson_amg_cuda.c — a FASP application switched to libfasp_cuda (plain C99).
 *
 * What a user of the reference's tutorial programs (tutorial/main/poisson-amg.c, poisson-pcg.c)
 * changes: the three calls marked [cuda] below. Everything else — parameter structs, the AMG
 * setup, the matrix containers — stays FASP's own. The 7-point Poisson matrix is generated in
 * memory (interior nodes of an n^3 grid, Dirichlet eliminated, diag 6, off-diagonals -1).
 *
 *   usage:   poisson_amg_cuda [n=32] [mode]          mode: 0 drop-in solve (default)
 *                                                          1 preconditioner plug-in + PCG
 *                                                          2 one hierarchy, several right-hand sides
 *   build:   gcc -std=c99 -Iinclude examples/poisson_amg_cuda.c -Lfaspsolver_b200/lib -lfasp_cuda \
 *                -Loracle/_ref -l:libfasp_seq.so -lm        (tests/test_boundary.py does exactly this)
 *
 * Exit status: 0 on success, 1 if a solve failed or missed the tolerance, 2 without a CUDA device
 * (the library has no CPU fallback; it says so through fasp_cuda_last_error()).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fasp_cuda.h" /* brings layout-identical mirrors of FASP's structs when fasp.h is absent */

/* the handful of FASP host routines this program calls (base/include/fasp_functs.h) */
void      fasp_param_amg_init(AMG_param* amgparam);                 /* AuxParam.c:431 */
void      fasp_param_solver_init(ITS_param* itsparam);              /* AuxParam.c:572 */
dCSRmat   fasp_dcsr_create(const INT m, const INT n, const INT nnz); /* BlaSparseCSR.c:47 */
void      fasp_dcsr_cp(const dCSRmat* A, dCSRmat* B);               /* BlaSparseCSR.c:851 */
void      fasp_dcsr_free(dCSRmat* A);                               /* BlaSparseCSR.c:184 */
dvector   fasp_dvec_create(const INT m);                            /* AuxVector.c:62 */
void      fasp_dvec_free(dvector* u);                               /* AuxVector.c:145 */
AMG_data* fasp_amg_data_create(SHORT max_levels);                   /* PreDataInit.c:64 */
void      fasp_amg_data_free(AMG_data* mgl, AMG_param* param);      /* PreDataInit.c:101 */
SHORT     fasp_amg_setup_rs(AMG_data* mgl, AMG_param* param);       /* PreAMGSetupRS.c:52 */

static dCSRmat poisson7(int n)
{
    const int N = n * n * n;
    long long nnz = 7LL * N - 6LL * n * n;
    dCSRmat   A   = fasp_dcsr_create(N, N, (INT)nnz);
    int       k   = 0;
    for (int z = 0; z < n; ++z)
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) {
                const int i = (z * n + y) * n + x;
                A.IA[i]     = k;
                A.JA[k] = i, A.val[k++] = 6.0; /* diagonal first, as FASP's generator stores it */
                if (x > 0) A.JA[k] = i - 1, A.val[k++] = -1.0;
                if (x < n - 1) A.JA[k] = i + 1, A.val[k++] = -1.0;
                if (y > 0) A.JA[k] = i - n, A.val[k++] = -1.0;
                if (y < n - 1) A.JA[k] = i + n, A.val[k++] = -1.0;
                if (z > 0) A.JA[k] = i - n * n, A.val[k++] = -1.0;
                if (z < n - 1) A.JA[k] = i + n * n, A.val[k++] = -1.0;
            }
    A.IA[N] = k;
    return A;
}

static double relres(const dCSRmat* A, const double* b, const double* x)
{
    double rr = 0.0, bb = 0.0;
    for (int i = 0; i < A->row; ++i) {
        double t = b[i];
        for (int k = A->IA[i]; k < A->IA[i + 1]; ++k) t -= A->val[k] * x[A->JA[k]];
        rr += t * t, bb += b[i] * b[i];
    }
    return sqrt(rr / bb);
}

static int report(const char* what, INT status, const dCSRmat* A, const dvector* b, const dvector* x, double tol)
{
    if (status < 0) {
        printf("%s: FAILED, status %d (%s)\n", what, (int)status, fasp_cuda_last_error());
        return 1;
    }
    const double r = relres(A, b->val, x->val);
    printf("%s: %d iterations, true relative residual %.3e\n", what, (int)status, r);
    return r <= tol * 1.001 ? 0 : 1;
}

int main(int argc, char** argv)
{
    const int n    = argc > 1 ? atoi(argv[1]) : 32;
    const int mode = argc > 2 ? atoi(argv[2]) : 0;
    if (n < 4 || n > 400) {
        fprintf(stderr, "n must be in [4, 400]\n");
        return 1;
    }
    /* start-up handshake: same struct layout on both sides (sequential vs OpenMP FASP build) */
    if (fasp_cuda_abi_check(sizeof(dCSRmat), sizeof(AMG_data), sizeof(AMG_param)) < 0) {
        fprintf(stderr, "%s\n", fasp_cuda_last_error());
        return 1;
    }
    if (fasp_cuda_init(0) < 0) {
        fprintf(stderr, "libfasp_cuda: %s\n", fasp_cuda_last_error());
        return 2;
    }

    AMG_param amgparam;
    ITS_param itparam;
    fasp_param_amg_init(&amgparam);
    fasp_param_solver_init(&itparam);
    amgparam.smoother     = SMOOTHER_L1DIAG; /* a data-parallel smoother (the default GS is sequential) */
    amgparam.print_level  = 0;
    itparam.itsolver_type = SOLVER_CG;
    itparam.tol           = 1e-8;
    itparam.maxit         = 200;
    itparam.print_level   = 1;

    dCSRmat A = poisson7(n);
    dvector b = fasp_dvec_create(A.row), x = fasp_dvec_create(A.row);
    for (int i = 0; i < A.row; ++i) b.val[i] = 1.0, x.val[i] = 0.0;
    printf("7-point Poisson %d^3: %d rows, %d nonzeros\n", n, A.row, A.nnz);

    int bad = 0;
    if (mode == 0) {
        /* [cuda] was: fasp_solver_dcsr_krylov_amg(&A, &b, &x, &itparam, &amgparam);  (SolCSR.c:476) */
        INT st = fasp_cuda_solver_dcsr_krylov_amg(&A, &b, &x, &itparam, &amgparam);
        bad    = report("AMG-PCG (drop-in)", st, &A, &b, &x, itparam.tol);
    } else if (mode == 1) {
        /* [cuda] was: fasp_precond_setup(PREC_AMG, ...) + fasp_solver_dcsr_pcg(...)  (PreCSR.c:46, KryPcg.c:96) */
        precond* pc = fasp_cuda_precond_setup(PREC_AMG, &amgparam, NULL, &A);
        if (!pc) {
            printf("precond setup FAILED (%s)\n", fasp_cuda_last_error());
            bad = 1;
        } else {
            INT st = fasp_cuda_solver_dcsr_pcg(&A, &b, &x, pc, itparam.tol, itparam.abstol, itparam.maxit,
                                               itparam.stop_type, itparam.print_level);
            bad    = report("PCG + AMG plug-in", st, &A, &b, &x, itparam.tol);
            fasp_cuda_precond_free(pc);
        }
    } else {
        /* FASP's own setup once, hierarchy uploaded once, then only b and x move */
        AMG_data* mgl = fasp_amg_data_create(amgparam.max_levels);
        mgl[0].A      = fasp_dcsr_create(A.row, A.col, A.nnz);
        fasp_dcsr_cp(&A, &mgl[0].A);
        mgl[0].b = fasp_dvec_create(A.row);
        mgl[0].x = fasp_dvec_create(A.row);
        if (fasp_amg_setup_rs(mgl, &amgparam) < 0) {
            printf("fasp_amg_setup_rs FAILED\n");
            bad = 1;
        } else {
            fasp_cuda_solver* s = fasp_cuda_krylov_amg_create(mgl, &amgparam); /* [cuda] */
            if (!s) {
                printf("upload FAILED (%s)\n", fasp_cuda_last_error());
                bad = 1;
            } else {
                fasp_cuda_host_pin(b.val, sizeof(REAL) * (size_t)A.row); /* DMA instead of staging */
                fasp_cuda_host_pin(x.val, sizeof(REAL) * (size_t)A.row);
                for (int rhs = 0; rhs < 3 && !bad; ++rhs) {
                    for (int i = 0; i < A.row; ++i) b.val[i] = 1.0 + rhs * sin(0.01 * i), x.val[i] = 0.0;
                    INT  st = fasp_cuda_krylov_amg_solve(s, &b, &x, &itparam);
                    char what[64];
                    snprintf(what, sizeof(what), "right-hand side %d (%.2f ms on the device)", rhs,
                             fasp_cuda_solver_stat(s, 2));
                    bad = report(what, st, &A, &b, &x, itparam.tol);
                }
                fasp_cuda_host_unpin(b.val);
                fasp_cuda_host_unpin(x.val);
                fasp_cuda_krylov_amg_destroy(s);
            }
        }
        fasp_amg_data_free(mgl, &amgparam);
    }

    fasp_dcsr_free(&A);
    fasp_dvec_free(&b);
    fasp_dvec_free(&x);
    return bad;
}
