/* TEST/BENCH INFRASTRUCTURE ONLY — driver for `bench.py --impl reference`.
 *
 * Runs the UNMODIFIED reference (FASP 2.8.7, OpenMP build: oracle/_ref/libfasp_omp.so) through
 * its own public API on BASELINE configs[1]: 3-D 7-point Poisson n^3, rhs = 1, AMG-PCG tol 1e-8,
 * classical RS setup. This file is written for this repository (it contains no reference source);
 * it is compiled against the reference's headers where they lie (/root/reference/base/include)
 * by oracle/build_ref.sh and linked to libfasp_omp.so.
 *
 *   fasp_ref_bench n steps warmup [sample_iters]
 *
 * Setup runs once. Every warm-up / timed step is one FULL solve (fasp_solver_dcsr_pcg +
 * fasp_precond_amg to tol 1e-8, x0 = 0): nothing is extrapolated. Only when sample_iters > 0 is given
 * (a host too slow for full solves) is a step a bounded sample of `sample_iters` PCG iterations, scaled by
 * (iterations+1)/(sample_iters+1), and the JSON says so. Prints one JSON object on stdout.
 *
 * NOTE (SURVEY.md finding 2): the OpenMP build of FASP ignores the requested smoother and runs
 * multicolour Gauss-Seidel, so its iteration count differs from the sequential oracle.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "fasp.h"
#include "fasp_functs.h"

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* natural x-fastest ordering, ascending columns, diag 6 (n+1)^2, off -(n+1)^2 */
static dCSRmat poisson7(int n)
{
    const long N = (long)n * n * n;
    long nnz = 7 * N - 6L * n * n;
    dCSRmat A = fasp_dcsr_create((INT)N, (INT)N, (INT)nnz);
    const double s = (double)(n + 1) * (n + 1);
    long k = 0;
    for (int z = 0; z < n; ++z)
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) {
                const long i = x + (long)n * (y + (long)n * z);
                A.IA[i] = (INT)k;
                if (z > 0) { A.JA[k] = (INT)(i - (long)n * n); A.val[k++] = -s; }
                if (y > 0) { A.JA[k] = (INT)(i - n); A.val[k++] = -s; }
                if (x > 0) { A.JA[k] = (INT)(i - 1); A.val[k++] = -s; }
                A.JA[k] = (INT)i; A.val[k++] = 6 * s;
                if (x < n - 1) { A.JA[k] = (INT)(i + 1); A.val[k++] = -s; }
                if (y < n - 1) { A.JA[k] = (INT)(i + n); A.val[k++] = -s; }
                if (z < n - 1) { A.JA[k] = (INT)(i + (long)n * n); A.val[k++] = -s; }
            }
    A.IA[N] = (INT)k;
    return A;
}

int main(int argc, char** argv)
{
    const int n       = argc > 1 ? atoi(argv[1]) : 64;
    const int steps   = argc > 2 ? atoi(argv[2]) : 3;
    const int warmup  = argc > 3 ? atoi(argv[3]) : 1;
    int       sample  = argc > 4 ? atoi(argv[4]) : 0;   /* 0 = full solves */
    int threads = 1;
#ifdef _OPENMP
    threads = omp_get_max_threads();
#endif
    dCSRmat A = poisson7(n);
    const INT N = A.row;
    dvector b = fasp_dvec_create(N), x = fasp_dvec_create(N);
    fasp_dvec_set(N, &b, 1.0);

    ITS_param itparam;
    AMG_param amgparam;
    fasp_param_solver_init(&itparam);
    fasp_param_amg_init(&amgparam);
    itparam.itsolver_type = SOLVER_CG;
    itparam.tol           = 1e-8;
    itparam.maxit         = 500;
    itparam.print_level   = 0;
    amgparam.print_level  = 0;
    amgparam.smoother     = SMOOTHER_L1DIAG; /* ignored by the OpenMP build */

#if MULTI_COLOR_ORDER
    A.color = 0; A.IC = NULL; A.ICMAP = NULL;
#endif
    /* setup, as fasp_solver_dcsr_krylov_amg does (SolCSR.c:500-533) */
    double t0 = now();
    AMG_data* mgl = fasp_amg_data_create(amgparam.max_levels);
    mgl[0].A = fasp_dcsr_create(A.row, A.col, A.nnz);
    fasp_dcsr_cp(&A, &mgl[0].A);
    mgl[0].b = fasp_dvec_create(N);
    mgl[0].x = fasp_dvec_create(N);
    if (fasp_amg_setup_rs(mgl, &amgparam) < 0) { fprintf(stderr, "setup failed\n"); return 2; }
    const double setup_s = now() - t0;
    precond_data pcdata;
    fasp_param_amg_to_prec(&pcdata, &amgparam);
    pcdata.max_levels = mgl[0].num_levels;
    pcdata.mgl_data   = mgl;
    precond pc;
    pc.data = &pcdata;
    pc.fct  = fasp_precond_amg;

    /* one untimed full solve: the reference's own iteration count */
    fasp_dvec_set(N, &x, 0.0);
    t0 = now();
    INT iters = fasp_solver_dcsr_pcg(&A, &b, &x, &pc, itparam.tol, itparam.abstol, itparam.maxit,
                                     itparam.stop_type, 0);
    const double first_s = now() - t0;
    if (iters <= 0) { fprintf(stderr, "reference solve failed: %d\n", iters); return 3; }
    if (sample > iters) sample = iters;

    double sum = 0.0, tmin = 1e300, tmax = 0.0;
    const int maxit_step = sample > 0 ? sample : itparam.maxit;
    for (int k = 0; k < warmup + steps; ++k) {
        fasp_dvec_set(N, &x, 0.0);
        t0 = now();
        fasp_solver_dcsr_pcg(&A, &b, &x, &pc, itparam.tol, itparam.abstol, maxit_step, itparam.stop_type, 0);
        const double dt = now() - t0;
        if (k >= warmup) {
            sum += dt;
            if (dt < tmin) tmin = dt;
            if (dt > tmax) tmax = dt;
        }
    }
    double per_step = steps > 0 ? sum / steps : first_s;
    if (steps <= 0) tmin = tmax = first_s;
    /* sampled mode only: a k-iteration solve applies the preconditioner and A k+1 times (KryPcg.c:125-131) */
    const double ms = (sample > 0 ? per_step * (double)(iters + 1) / (sample + 1) : per_step) * 1e3;
    char what[256];
    if (sample > 0)
        snprintf(what, sizeof(what), "%d of %d PCG iterations per step, scaled by (%d+1)/(%d+1)", sample, (int)iters,
                 (int)iters, sample);
    else
        snprintf(what, sizeof(what), "full solves, nothing extrapolated (%d iterations each; min %.1f max %.1f ms)",
                 (int)iters, tmin * 1e3, tmax * 1e3);
    printf("{\"ms_per_solve\": %.3f, \"extrapolated\": %s, \"ms_first_solve\": %.3f, \"iterations\": %d, "
           "\"levels\": %d, \"setup_s\": %.2f, \"threads\": %d, \"n\": %d, "
           "\"sample\": \"%s; OpenMP FASP, %d threads, multicolour GS\"}\n",
           ms, sample > 0 ? "true" : "false", first_s * 1e3, (int)iters, (int)mgl[0].num_levels, setup_s, threads, n,
           what, threads);
    return 0;
}
