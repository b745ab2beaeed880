// gmres.cu — right-preconditioned restarted GMRES on the device: fixed restart
// (fasp_solver_dcsr_pgmres, KryPgmres.c:66), variable restart (fasp_solver_dcsr_pvgmres,
// KryPvgmres.c:66-387, Baker/Jessup/Kolev adaptation :200-210) and the flexible variant
// (fasp_solver_dcsr_pvfgmres, KryPvfgmres.c:67-358: keeps z_j = B p_j, stops on the absolute
// residual against tol * ||b||); BSR twins share the code.
//
// Device residency: basis vectors, Hessenberg matrix, Givens rotations, rs[] and all
// norms live in HBM. Inside a restart cycle nothing returns to the host: every inner step is
// a fixed kernel sequence gated by device flags (a step that follows the convergence break is
// skipped on the device). The host reads one small status struct per RESTART CYCLE to choose
// the next restart length, exactly where the CPU code re-enters its outer loop.
//
// Modified Gram-Schmidt is fused: step j of the orthogonalisation applies the previous axpy
// and accumulates the next dot product in one pass (i+2 vector kernels per inner step instead
// of the reference's 2i+2 BLAS-1 calls), with the same sequence of floating-point operations
// per entry.
#include "krylov.cuh"
#include "reduce.cuh"
#include "comm.cuh"
#include "p2p.cuh"

namespace fc {

struct GmState {
    double r_norm, r_norm_old, absres0, absres, relres, normu, cr, tol, abstol;
    double rr, t2, xx, scale;
    double den_norm, epsilon;   // flexible variant: ||b|| (or ||r0||) and tol * den_norm
    double b_norm;              // flexible variant: ||b|| as printed by ITS_PUTNORM
    int    iter, maxit, i, stop_type, variable, flexible;
    int    done, converged, silent;
    int    notable;   // converged before the loop: the reference jumps to FINISHED without printing the table
    int    skip_inner, skip_scale, skip_true, skip_copy;
    int    R;   // leading dimension of hh: hh[j][k] = H[j * R + k]
};

struct GmPinned {
    int    done, iter, converged, pad;
    double cr, relres;
};

// p0 = b - A x done; rr = ||p0||^2, xx = ||x||^2 (MOD only)     (KryPvgmres.c:149-182)
__global__ void k_gm_init(GmState* st, double* norms, double* habs)
{
    st->r_norm = sqrt(st->rr);
    if (st->flexible) {   // KryPvfgmres.c:147-167, xx = ||b||^2 here
        const double b_norm = sqrt(st->xx);
        st->b_norm   = b_norm;
        st->den_norm = (b_norm > 0.0) ? b_norm : st->r_norm;
        st->epsilon  = st->tol * st->den_norm;
        st->absres0 = st->absres = st->r_norm;
        st->relres  = (b_norm > 0.0) ? st->r_norm / b_norm : st->r_norm;
        norms[0]    = st->relres;
        habs[0]     = st->r_norm;
        if (st->r_norm < st->epsilon || st->r_norm < st->abstol || st->r_norm == 0.0) {
            st->converged = 1;
            st->done      = 1;
            st->silent    = 1;   // "goto FINISHED" / early return: no final line
        }
        return;
    }
    if (st->stop_type == STOP_MOD_REL_RES) {
        st->normu   = fmax(SMALLREAL, sqrt(st->xx));
        st->absres0 = st->r_norm;
        st->relres  = st->absres0 / st->normu;
    } else if (st->stop_type == STOP_REL_PRECRES) {   // xx = (p0, B p0), KryPvgmres.c:160-167
        const double r_normb = sqrt(st->xx);
        st->absres0 = fmax(SMALLREAL, r_normb);
        st->relres  = r_normb / st->absres0;
    } else {
        st->absres0 = fmax(SMALLREAL, st->r_norm);
        st->relres  = st->r_norm / st->absres0;
    }
    st->absres = st->absres0;
    norms[0]   = st->relres;
    if (st->relres < st->tol || st->absres0 < st->abstol) {
        st->converged = 1;
        st->done      = 1;
        st->notable   = 1;
    }
}

// rs[0] = r_norm_old = r_norm ; p0 *= 1/r_norm                    (:190-195)
__global__ void k_gm_cycle_start(GmState* st, double* rs)
{
    if (!st->done && st->flexible && st->r_norm == 0.0) {   // KryPvfgmres.c:181-188
        st->done = st->converged = st->silent = 1;
    }
    if (st->done) {
        st->skip_inner = 1;
        return;
    }
    rs[0]          = st->r_norm;
    st->r_norm_old = st->r_norm;
    st->scale      = 1.0 / st->r_norm;
    st->i          = 0;
    st->skip_inner = 0;
    st->skip_scale = 1;
    st->skip_true  = 1;
    st->skip_copy  = 1;
}

// x *= st->scale (fasp_blas_darray_ax, BlaArray.c:43: no-op for a == 1)
__global__ void __launch_bounds__(256)
k_gm_scale(const GmState* st, const int* gate, double* __restrict__ x, size_t n)
{
    if (*gate) return;
    const double a = st->scale;
    if (a == 1.0) return;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        x[i] = __dmul_rn(a, x[i]);
}

// fused modified Gram-Schmidt step j of inner iteration i (1-based), KryPvgmres.c:228-231:
//   if j > 0 : p_i -= hh[j-1][i-1] * p_{j-1}
//   if j < i : hh[j][i-1] = (p_j, p_i)        else (j == i): t2 = ||p_i||^2
__global__ void __launch_bounds__(256)
k_gm_mgs(GmState* st, double* hh, int j, int i, const double* __restrict__ pjm1,
         const double* __restrict__ pj, double* __restrict__ pi, size_t n, double* partials,
         unsigned int* ticket)
{
    if (st->skip_inner) return;
    const int    R    = st->R;
    const double hneg = (j > 0) ? -hh[(size_t)(j - 1) * R + (i - 1)] : 0.0;
    double       v[1] = {0.0};
    for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (size_t)gridDim.x * 256) {
        double x = pi[k];
        if (j > 0) {
            // fasp_blas_darray_axpy(n, -h, p_{j-1}, p_i): y += a*x with the a == +-1 shortcuts
            x     = __dadd_rn(x, __dmul_rn(hneg, pjm1[k]));
            pi[k] = x;
        }
        v[0] += (j < i) ? pj[k] * x : x * x;
    }
    grid_reduce<1, 0>(v, partials, ticket, [&](const double* s) {
        if (j < i) hh[(size_t)j * R + (i - 1)] = s[0];
        else st->t2 = s[0];
    });
}

// scalar part of inner step i: Hessenberg column, Givens rotations, residual estimate
// (KryPvgmres.c:232-263 ; fixed-restart variants KryPgmres.c:205-230)
__global__ void k_gm_givens(GmState* st, double* hh, double* c, double* s, double* rs,
                            double* norms, double* hfac, int i)
{
    if (st->skip_inner) {
        st->skip_scale = 1;
        return;
    }
    const int R = st->R;
    st->i       = i;
    st->iter += 1;
    double t                      = sqrt(st->t2);
    hh[(size_t)i * R + (i - 1)]   = t;
    const bool do_scale           = (st->variable || st->flexible) ? (t != 0.0) : (fabs(t) > SMALLREAL);
    st->scale                     = do_scale ? 1.0 / t : 1.0;
    st->skip_scale                = 0;
    for (int j = 1; j < i; ++j) {
        t                               = hh[(size_t)(j - 1) * R + (i - 1)];
        hh[(size_t)(j - 1) * R + (i - 1)] = s[j - 1] * hh[(size_t)j * R + (i - 1)] + c[j - 1] * t;
        hh[(size_t)j * R + (i - 1)]       = -s[j - 1] * t + c[j - 1] * hh[(size_t)j * R + (i - 1)];
    }
    const double hi  = hh[(size_t)i * R + (i - 1)];
    const double hd  = hh[(size_t)(i - 1) * R + (i - 1)];
    t                = hi * hi;
    t += hd * hd;
    double gamma = sqrt(t);
    if (st->variable || st->flexible) {
        if (gamma == 0.0) gamma = SMALLREAL;
    } else {
        gamma = fmax(gamma, SMALLREAL);
    }
    c[i - 1]  = hd / gamma;
    s[i - 1]  = hi / gamma;
    rs[i]     = -s[i - 1] * rs[i - 1];
    rs[i - 1] = c[i - 1] * rs[i - 1];
    hh[(size_t)(i - 1) * R + (i - 1)] = s[i - 1] * hi + c[i - 1] * hd;
    st->absres = st->r_norm = fabs(rs[i]);
    if (st->flexible) {   // KryPvfgmres.c:258-273: table relative to ||b||, exit on r_norm <= epsilon
        st->relres      = (st->den_norm > 0.0) ? st->r_norm / st->den_norm : st->r_norm;
        norms[st->iter] = st->relres;
        hfac[st->iter]  = st->absres;
        if (st->r_norm <= st->epsilon || st->iter >= st->maxit) st->skip_inner = 1;
        return;
    }
    st->relres              = st->absres / st->absres0;
    norms[st->iter]         = st->relres;
    hfac[st->iter]          = st->absres;
    if (st->relres < st->tol || st->iter >= st->maxit) st->skip_inner = 1;
}

// back substitution for the i x i triangular system (KryPvgmres.c:271-278)
__global__ void k_gm_backsolve(GmState* st, const double* hh, double* rs)
{
    if (st->done) return;
    const int R = st->R, i = st->i;
    rs[i - 1] = rs[i - 1] / hh[(size_t)(i - 1) * R + (i - 1)];
    for (int k = i - 2; k >= 0; --k) {
        double t = 0.0;
        for (int j = k + 1; j < i; ++j) t -= hh[(size_t)k * R + j] * rs[j];
        t += rs[k];
        rs[k] = t / hh[(size_t)k * R + k];
    }
}

// w = rs[i-1] p_{i-1} + sum_{j=i-2..0} rs[j] p_j        (KryPvgmres.c:280-284)
__global__ void __launch_bounds__(256)
k_gm_form_w(const GmState* st, const double* __restrict__ rs, const double* __restrict__ P,
            size_t ldp, double* __restrict__ w, size_t n)
{
    if (st->done) return;
    const int i = st->i;
    for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (size_t)gridDim.x * 256) {
        double a = rs[i - 1];
        double x = P[(size_t)(i - 1) * ldp + k];
        x        = (a == 1.0) ? x : __dmul_rn(a, x);
        for (int j = i - 2; j >= 0; --j) x = __dadd_rn(x, __dmul_rn(rs[j], P[(size_t)j * ldp + k]));
        w[k] = x;
    }
}

// x += r
__global__ void __launch_bounds__(256)
k_gm_add(const GmState* st, const double* __restrict__ r, double* __restrict__ x, size_t n)
{
    if (st->done) return;
    for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (size_t)gridDim.x * 256)
        x[k] = __dadd_rn(x[k], r[k]);
}

__global__ void k_gm_after_update(GmState* st)
{
    if (st->done) return;
    st->skip_true = st->flexible ? !(st->r_norm <= st->epsilon) : !(st->relres < st->tol);
}

// false-convergence check with the true residual (KryPvgmres.c:295-340)
__global__ void k_gm_truecheck(GmState* st, double* norms)
{
    if (st->done || st->skip_true) return;
    st->skip_true = 1;
    st->r_norm    = sqrt(st->rr);
    st->absres    = st->r_norm;
    if (st->flexible) {   // KryPvfgmres.c:294-325
        double relres;
        if (st->stop_type == STOP_MOD_REL_RES) {
            st->normu = fmax(SMALLREAL, sqrt(st->xx));
            relres    = st->r_norm / st->normu;
        } else if (st->stop_type == STOP_REL_PRECRES) {   // xx = (B r, r), KryPvfgmres.c:288-293
            relres = sqrt(st->xx) / st->den_norm;
        } else {
            relres = st->r_norm / st->den_norm;
        }
        st->relres = st->r_norm / st->den_norm;   // what ITS_FINAL reports (:358)
        if (relres <= st->tol) {
            st->converged = 1;
            st->done      = 1;
        } else {
            st->skip_copy = 0;
            st->i         = 0;
        }
        return;
    }
    if (st->stop_type == STOP_MOD_REL_RES) {
        st->normu  = fmax(SMALLREAL, sqrt(st->xx));
        st->relres = st->absres / st->normu;
    } else if (st->stop_type == STOP_REL_PRECRES) {   // xx = (B r, r), KryPvgmres.c:312-319
        st->absres = sqrt(st->xx);
        st->relres = st->absres / st->absres0;
    } else {
        st->relres = st->absres / st->absres0;
    }
    norms[st->iter] = st->relres;
    if (st->relres < st->tol) {
        st->converged = 1;
        st->done      = 1;
    } else {
        st->skip_copy = 0;   // p0 = r, i = 0
        st->i         = 0;
    }
}

// rs[] of the residual vector in the Krylov basis (KryPvgmres.c:343-346)
__global__ void k_gm_resvec_scalars(GmState* st, double* rs, const double* c, const double* s)
{
    if (st->done) return;
    for (int j = st->i; j > 0; --j) {
        rs[j - 1] = -s[j - 1] * rs[j];
        rs[j]     = c[j - 1] * rs[j];
    }
}

// p0 <- residual vector rebuilt from the basis, no SpMV (KryPvgmres.c:348-355):
//   p_i += (rs_i - 1) p_i ; p_i += rs_j p_j (j = i-1..1) ; p_0 += (rs_0 - 1) p_0 ; p_0 += p_i
__global__ void __launch_bounds__(256)
k_gm_resvec(const GmState* st, const double* __restrict__ rs, double* __restrict__ P,
            size_t ldp, size_t n)
{
    if (st->done) return;
    const int i = st->i;
    if (i == 0) return;
    const double ai = rs[i] - 1.0, a0 = rs[0] - 1.0;
    for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (size_t)gridDim.x * 256) {
        double pi = P[(size_t)i * ldp + k];
        pi        = __dadd_rn(pi, __dmul_rn(ai, pi));
        for (int j = i - 1; j > 0; --j) pi = __dadd_rn(pi, __dmul_rn(rs[j], P[(size_t)j * ldp + k]));
        double p0 = P[k];
        p0        = __dadd_rn(p0, __dmul_rn(a0, p0));
        P[k]      = __dadd_rn(p0, pi);
    }
}

__global__ void k_gm_cycle_end(GmState* st, GmPinned* out)
{
    if (!st->done) {
        st->cr = st->r_norm / st->r_norm_old;
        if (st->iter >= st->maxit) st->done = 1;
    }
    out->done      = st->done;
    out->iter      = st->iter;
    out->converged = st->converged;
    out->cr        = st->cr;
    out->relres    = st->relres;
}

static int ggrid(size_t n)
{
    size_t g   = (n + 1023) / 1024;
    size_t cap = (size_t)ctx().sm_count * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

GmresCache::~GmresCache() { release(); }
void GmresCache::release()
{
    for (auto& g : step_graph) g.reset();
    step_graph.clear();
    start_graph.reset();
    end_graph.reset();
    init_graph.reset();
    if (t0) cudaEventDestroy(t0);
    if (t1) cudaEventDestroy(t1);
    t0 = t1 = nullptr;
    if (pin_h) cudaFreeHost(pin_h);
    pin_h = nullptr;
    if (pin_flags) cudaFreeHost(pin_flags);
    pin_flags = nullptr;
    for (auto& e : ev) cudaEventDestroy(e);
    ev.clear();
    dfree(pin_d);
    pin_d = nullptr;
    if (registered && work) p2p_unregister(work);
    registered = false;
    dfree(work);
    dfree(st);
    work = nullptr, st = nullptr;
    n = 0, ldp = 0, R = 0, hcap = 0, kind = -1;
}

int gmres_solve(LinOp& A, const double* b, double* x, Prec& pc, double tol, double abstol,
                int MaxIt, int restart, int StopType, int PrtLvl, int kind,
                SolveStats* stats, GmresCache* cache)
{
    const bool flexible = (kind == GM_FLEXIBLE);
    const bool variable = (kind == GM_VARIABLE) || flexible;   // both adapt the restart length
    ensure_init();
    Ctx&         c = ctx();
    const size_t n = (size_t)A.n;
    const bool   precres = (StopType == STOP_REL_PRECRES);
    if (StopType != STOP_REL_RES && StopType != STOP_MOD_REL_RES && !precres)
        fail(ERROR_INPUT_PAR, "device GMRES: unknown stop_type %d", StopType);
    if (restart < 1) fail(ERROR_INPUT_PAR, "GMRES restart must be positive");
    if (PrtLvl > PRINT_NONE)
        printf("\nCalling %s solver (%s) ...\n", flexible ? "VFGMRes" : (variable ? "VGMRes" : "GMRes"),
               A.format());

    const long long launches0   = c.launches;
    const int       restart_max = variable ? restart : (restart < MaxIt ? restart : (MaxIt > 0 ? MaxIt : 1));
    const int       R           = restart_max;
    const size_t    ldp         = (A.vec_capacity() + 1) & ~(size_t)1;   // room for ghosts
    const int       hcap        = MaxIt + 2;

    GmresCache  local;
    GmresCache& W = cache ? *cache : local;
    // basis p[0..R], w, r, (flexible: z[0..R-1]) ; hh (R+1) x R, c, s, rs ; norms + absres history
    const size_t nsmall = (size_t)(R + 1) * R + 2 * (size_t)R + (R + 1) + 2 * (size_t)hcap;
    const size_t nvec   = (size_t)(R + 3) + (flexible ? (size_t)R : 0);
    const int look = c.opt.lookahead < 1 ? 1 : c.opt.lookahead;
    if (W.n != n || W.ldp != ldp || W.R != R || W.hcap != hcap || W.kind != kind || W.look != look) {
        W.release();
        W.look = look;
        FC_CUDA(cudaMallocHost(&W.pin_flags, sizeof(int) * (look + 1)));
        W.ev.resize(look + 1);
        for (auto& e : W.ev) FC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        W.work = dalloc<double>(nvec * ldp + nsmall);
        if (p2p_active() && A.distributed()) {   // basis vectors are gathered by the peers' kernels
            p2p_register(W.work, sizeof(double) * (nvec * ldp + nsmall));
            W.registered = true;
        }
        W.st    = static_cast<void*>(dalloc<GmState>(1));
        W.pin_d = static_cast<void*>(dalloc<GmPinned>(1));
        FC_CUDA(cudaMallocHost(&W.pin_h, sizeof(GmPinned)));
        FC_CUDA(cudaEventCreate(&W.t0));
        FC_CUDA(cudaEventCreate(&W.t1));
        W.step_graph.clear();
        W.step_graph.resize(R + 1);
        W.n = n, W.ldp = ldp, W.R = R, W.hcap = hcap, W.kind = kind;
    }
    // graphs bake in pointers and kernel choices: re-capture when any of them changed
    if (W.kA != A.key() || W.kb != b || W.kx != x || W.kpc != pc.key() || W.kstop != StopType ||
        W.kepoch != c.opt_epoch) {
        for (auto& g : W.step_graph) g.reset();
        W.start_graph.reset();
        W.end_graph.reset();
        W.init_graph.reset();
        W.kA = A.key(), W.kb = b, W.kx = x, W.kpc = pc.key(), W.kstop = StopType;
        W.kepoch = c.opt_epoch;
    }
    double*   work  = W.work;
    GmState*  st    = static_cast<GmState*>(W.st);
    GmPinned* pin_d = static_cast<GmPinned*>(W.pin_d);
    GmPinned* pin_h = static_cast<GmPinned*>(W.pin_h);
    cudaEvent_t t0 = W.t0, t1 = W.t1;
    std::vector<CapturedGraph>& step_graph = W.step_graph;
    CapturedGraph&              start_graph = W.start_graph;
    CapturedGraph&              end_graph   = W.end_graph;
    int                         ret = 0;

    auto cleanup = [&]() {};

    try {
        double* P  = work;
        double* w  = P + (size_t)(R + 1) * ldp;
        double* r  = w + ldp;
        double* Z  = r + ldp;
        double* hh = Z + (flexible ? (size_t)R * ldp : 0);
        double* cc = hh + (size_t)(R + 1) * R;
        double* ss = cc + R;
        double* rs = ss + R;
        double* norms = rs + (R + 1);
        double* habs  = norms + hcap;
        FC_CUDA(cudaMemsetAsync(hh, 0, sizeof(double) * nsmall, c.stream));
        GmState h0;
        memset(&h0, 0, sizeof(h0));
        h0.tol = tol, h0.abstol = abstol, h0.maxit = MaxIt, h0.stop_type = StopType;
        h0.variable = (variable && !flexible) ? 1 : 0, h0.flexible = flexible ? 1 : 0, h0.R = R, h0.cr = 1.0;
        h0.skip_inner = h0.skip_scale = h0.skip_true = h0.skip_copy = 1;
        h0.absres0 = h0.absres = h0.relres = h0.normu = BIGREAL;
        FC_CUDA(cudaMemcpyAsync(st, &h0, sizeof(h0), cudaMemcpyHostToDevice, c.stream));
        red_partials((size_t)c.sm_count * 8);
        const int  g         = ggrid(n);
        const bool use_graph = c.opt.graph && !c.opt.profile && pc.capturable();
        const bool glob      = A.distributed();   // sums over all ranks' rows
        auto       pvec      = [&](int k) { return P + (size_t)k * ldp; };

        auto initial = [&]() {
            Reduce red;
            red.global = glob;
            red.nrm2_out = &st->rr;
            A.apply(CSR_RESID, 1.0, x, b, pvec(0), red, nullptr);
            if (flexible || StopType == STOP_MOD_REL_RES) {
                Reduce rx;
                rx.global   = glob;
                rx.nrm2_out = &st->xx;
                vec_reduce(flexible ? b : x, n, rx, nullptr);   // flexible: ||b|| (KryPvfgmres.c:150)
            }
            if (precres && !flexible) {   // r_normb = sqrt((p0, B p0)), KryPvgmres.c:160-167
                Reduce rz;
                rz.global   = glob;
                rz.dot_with = pvec(0);
                rz.dot_out  = &st->xx;
                pc.apply(pvec(0), r, rz, nullptr);
            }
            FC_LAUNCH(k_gm_init, 1, 1, 0, st, norms, habs);
            FC_LAUNCH(k_gm_cycle_end, 1, 1, 0, st, pin_d);
        };
        auto inner_step = [&](int i) {
            const int* gate = &st->skip_inner;
            double* zi = flexible ? Z + (size_t)(i - 1) * ldp : r;   // flexible keeps z_{i-1} (:226-231)
            pc.apply(pvec(i - 1), zi, Reduce(), gate);
            A.apply(CSR_MXV, 1.0, zi, nullptr, pvec(i), Reduce(), gate);
            for (int j = 0; j <= i; ++j)
            {
                FC_LAUNCH(k_gm_mgs, g, 256, 0, st, hh, j, i, j > 0 ? pvec(j - 1) : nullptr,
                          j < i ? pvec(j) : nullptr, pvec(i), n, red_partials(g), red_ticket());
                if (glob) comm_allreduce(j < i ? hh + (size_t)j * R + (i - 1) : &st->t2, 1, 0, gate);
            }
            FC_LAUNCH(k_gm_givens, 1, 1, 0, st, hh, cc, ss, rs, norms, habs, i);
            FC_LAUNCH(k_gm_scale, g, 256, 0, st, &st->skip_scale, pvec(i), n);
        };
        auto cycle_start = [&]() {
            FC_LAUNCH(k_gm_cycle_start, 1, 1, 0, st, rs);
            FC_LAUNCH(k_gm_scale, g, 256, 0, st, &st->skip_inner, pvec(0), n);
        };
        auto cycle_end = [&]() {
            const int* done = &st->done;
            FC_LAUNCH(k_gm_backsolve, 1, 1, 0, st, hh, rs);
            if (flexible) {   // x += sum_j rs_j z_j (KryPvfgmres.c:286-291), no preconditioner call
                FC_LAUNCH(k_gm_form_w, g, 256, 0, st, rs, Z, ldp, r, n);
            } else {
                FC_LAUNCH(k_gm_form_w, g, 256, 0, st, rs, P, ldp, w, n);
                pc.apply(w, r, Reduce(), done);
            }
            FC_LAUNCH(k_gm_add, g, 256, 0, st, r, x, n);
            FC_LAUNCH(k_gm_after_update, 1, 1, 0, st);
            Reduce red;
            red.global = glob;
            red.nrm2_out = &st->rr;
            A.apply(CSR_RESID, 1.0, x, b, r, red, &st->skip_true, true);
            if (StopType == STOP_MOD_REL_RES) {
                Reduce rx;
            rx.global = glob;
                rx.nrm2_out = &st->xx;
                vec_reduce(x, n, rx, &st->skip_true);
            }
            if (precres) {   // (B r, r) of the true residual; the preconditioner call is gated like the residual
                Reduce rz;
                rz.global   = glob;
                rz.dot_with = r;
                rz.dot_out  = &st->xx;
                pc.apply(r, w, rz, &st->skip_true);
            }
            FC_LAUNCH(k_gm_truecheck, 1, 1, 0, st, norms);
            vec_copy(pvec(0), r, n, &st->skip_copy);
            FC_LAUNCH(k_gm_resvec_scalars, 1, 1, 0, st, rs, cc, ss);
            FC_LAUNCH(k_gm_resvec, g, 256, 0, st, rs, P, ldp, n);
            FC_LAUNCH(k_gm_cycle_end, 1, 1, 0, st, pin_d);
        };

        // graphs are built before the timed region (nothing executes during a capture)
        W.init_graph.prepare(use_graph, initial);
        start_graph.prepare(use_graph, cycle_start);
        end_graph.prepare(use_graph, cycle_end);
        FC_CUDA(cudaEventRecord(t0, c.stream));
        W.init_graph.run(use_graph, initial);
        FC_CUDA(cudaMemcpyAsync(pin_h, pin_d, sizeof(GmPinned), cudaMemcpyDeviceToHost, c.stream));
        FC_CUDA(cudaStreamSynchronize(c.stream));

        int    Restart = restart_max;
        int    iter    = 0;
        double cr      = 1.0;
        bool   first   = true;
        while (!pin_h->done && iter < MaxIt) {
            if (variable) {   // KryPvgmres.c:200-210
                const double cr_max = 0.99, cr_min = 0.174;
                const int    d = 3, restart_min = 3;
                if (cr > cr_max || first) Restart = restart_max;
                else if (cr < cr_min) { /* keep */ }
                else if (Restart - d > restart_min) Restart -= d;
                else Restart = restart_max;
            }
            first = false;
            start_graph.run(use_graph, cycle_start);
            int steps = Restart;
            if (steps > MaxIt - iter) steps = MaxIt - iter;
            // Inner steps are enqueued `look` ahead of an asynchronous read of the device's skip_inner flag: once
            // the restart cycle has met its exit test (KryPvgmres.c:263) the remaining steps would all return at
            // once, so they are not launched at all (a converged 11-iteration GMRES(30) solve used to enqueue 19
            // gated V-cycles, ~200 empty kernels each).
            bool exited = false;
            for (int i = 1; i <= steps && !exited; ++i) {
                step_graph[i].run(use_graph, [&]() { inner_step(i); });
                const int slot = i % (look + 1);
                FC_CUDA(cudaMemcpyAsync(&W.pin_flags[slot], &st->skip_inner, sizeof(int), cudaMemcpyDeviceToHost,
                                        c.stream));
                FC_CUDA(cudaEventRecord(W.ev[slot], c.stream));
                if (i > look) {
                    const int old = (i - look) % (look + 1);
                    FC_CUDA(cudaEventSynchronize(W.ev[old]));
                    if (W.pin_flags[old]) exited = true;
                }
            }
            end_graph.run(use_graph, cycle_end);
            FC_CUDA(cudaMemcpyAsync(pin_h, pin_d, sizeof(GmPinned), cudaMemcpyDeviceToHost, c.stream));
            FC_CUDA(cudaStreamSynchronize(c.stream));
            iter = pin_h->iter;
            cr   = pin_h->cr;
        }
        FC_CUDA(cudaEventRecord(t1, c.stream));
        FC_CUDA(cudaStreamSynchronize(c.stream));

        GmState hs;
        FC_CUDA(cudaMemcpy(&hs, st, sizeof(hs), cudaMemcpyDeviceToHost));
        if (stats || PrtLvl >= PRINT_SOME) {
            const int           nh = (hs.iter + 1 < hcap) ? hs.iter + 1 : hcap;
            std::vector<double> h2(2 * (size_t)hcap);
            FC_CUDA(cudaMemcpy(h2.data(), norms, sizeof(double) * 2 * hcap, cudaMemcpyDeviceToHost));
            if (PrtLvl >= PRINT_SOME && flexible) {   // ITS_PUTNORM, KryPvfgmres.c:153-156
                printf("L2 norm of right-hand side = %.10e.\n", hs.b_norm);
                printf("L2 norm of residual = %.10e.\n", hs.absres0);
            }
            if (PrtLvl >= PRINT_SOME && !hs.silent && !hs.notable) {   // a solve that stops before the loop prints no table
                print_itinfo(PrtLvl, StopType, 0, h2[0], hs.absres0, 0.0);
                for (int i = 1; i < nh; ++i)
                    print_itinfo(PrtLvl, StopType, i, h2[i], h2[hcap + i], h2[i] / h2[i - 1]);
            }
            if (stats) {
                stats->hist_relres.assign(h2.begin(), h2.begin() + nh);
                stats->hist_absres.assign(h2.begin() + hcap, h2.begin() + hcap + nh);
                stats->hist_factor.clear();
            }
        }
        if (PrtLvl > PRINT_NONE && !hs.silent) print_final(hs.iter, MaxIt, hs.relres);
        float ms = 0.f;
        FC_CUDA(cudaEventElapsedTime(&ms, t0, t1));
        if (stats) {
            stats->iters    = hs.iter;
            stats->relres   = hs.relres;
            stats->ms       = ms;
            stats->launches = c.launches - launches0;
        }
        ret = (hs.iter >= MaxIt) ? ERROR_SOLVER_MAXIT : hs.iter;
    } catch (...) {
        cudaStreamSynchronize(c.stream);
        cleanup();
        throw;
    }
    cleanup();
    return ret;
}

} // namespace fc
