// amg.cu — AMG hierarchy upload and the multigrid cycle on the device.
//
// Replaces, on the solve path:
//   fasp_solver_mgcycle      PreMGCycle.c:48-274   (V / W / VW / WV, non-recursive)
//   fasp_dcsr_presmoothing / _postsmoothing  PreMGSmoother.inl:49,155 (Jacobi, L1, poly)
//   fasp_coarse_itsolver     PreMGUtil.inl:37      (-> dense inverse, dense.cu)
//   fasp_precond_amg         PreCSR.c:416-435
// The hierarchy itself (A_l, P_l, R_l = P_l^T) comes from FASP's host setup and is uploaded
// once; nothing here allocates per cycle (the CPU smoothers calloc two N-vectors per call).
//
// Per level of a V(1,1) cycle the device runs, with x_l = 0 on entry (PreCSR.c:430,
// PreMGCycle.c:151):
//   pre-smooth   x = D~^-1 b              vector kernel, no pass over A (b - A*0 == b exactly)
//   residual     w = b - A x              1 pass over A_l
//   restrict     b_{l+1} = R w            1 pass over R_l
//   ...coarse...
//   prolongate   x += P x_{l+1}           1 pass over P_l
//   post-smooth  x = S(x)                 1 pass over A_l
#include "amg.cuh"
#include "comm.cuh"
#include "dist.cuh"
#include "p2p.cuh"

namespace fc {

static bool smoother_supported(short s)
{
    return s == SMOOTHER_JACOBI || s == SMOOTHER_L1DIAG || s == SMOOTHER_POLY ||
           (s == SMOOTHER_GS && ctx().opt.gs_multicolor);
}

void amg_set_params(Amg& h, const AMG_param* p)
{
    h.amg_type       = p->AMG_type;
    h.smoother       = p->smoother;
    h.cycle_type     = p->cycle_type;
    h.presmooth      = p->presmooth_iter;
    h.postsmooth     = p->postsmooth_iter;
    h.ndeg           = p->polynomial_degree;
    h.coarse_scaling = p->coarse_scaling;
    h.coarse_solver  = p->coarse_solver;
    h.smooth_order   = p->smooth_order;
    h.relax          = p->relaxation;
    h.tol            = p->tol;
    h.maxit          = p->maxit;
}

void amg_level_vectors(Amg& h, Level& L)
{
    if (L.cap < L.n) L.cap = L.n;
    const size_t c = (size_t)L.cap + 8;
    L.b  = dalloc<double>(c);
    L.xa = dalloc<double>(c);
    L.xb = dalloc<double>(c);
    L.w  = dalloc<double>(c);
    h.bytes += 4 * sizeof(double) * c;
}

// work vectors of the polynomial smoother: as long as the level's other vectors (multi-GPU: L.cap must already be
// the capacity agreed by all ranks, the peers push ghost entries behind the owned part)
void amg_level_poly_vectors(Amg& h, Level& L)
{
    for (int i = 0; i < 3; ++i)
        if (!L.pv[i]) {
            L.pv[i] = dalloc<double>((size_t)(L.cap > L.n ? L.cap : L.n) + 8);
            h.bytes += sizeof(double) * (size_t)(L.cap > L.n ? L.cap : L.n);
        }
}

void amg_level_smoother_data(Amg& h, Level& L, const dCSRmat* hostA)
{
    switch (h.smoother) {
        case SMOOTHER_GS: {
            std::vector<int> ic, icmap;
            gs_multicolor_host(hostA->row, hostA->IA, hostA->JA, ic, icmap);
            L.color_ptr  = ic;
            L.ncolors    = (int)ic.size() - 1;
            L.color_rows = dalloc<int>(icmap.size() ? icmap.size() : 1);
            FC_CUDA(cudaMemcpyAsync(L.color_rows, icmap.data(), sizeof(int) * icmap.size(),
                                    cudaMemcpyHostToDevice, ctx().stream));
            FC_CUDA(cudaStreamSynchronize(ctx().stream));
            h.bytes += sizeof(int) * icmap.size();
            break;
        }
        case SMOOTHER_JACOBI: csr_ensure_diag(L.A); break;
        case SMOOTHER_L1DIAG: csr_ensure_l1(L.A); break;
        case SMOOTHER_POLY: {
            // constants of fasp_smoother_dcsr_poly, ItrSmootherCSRpoly.c:94-107
            double nrm = csr_dinv_a_norminf(L.A);
            if (L.dist && comm_active()) {   // the norm is a maximum over ALL rows
                double* d = dalloc<double>(1);
                FC_CUDA(cudaMemcpyAsync(d, &nrm, sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
                comm_allreduce(d, 1, 2);
                FC_CUDA(cudaMemcpyAsync(&nrm, d, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
                FC_CUDA(cudaStreamSynchronize(ctx().stream));
                dfree(d);
            }
            double mu0        = 1.0 / nrm;
            const double mu1  = 4.0 * mu0;
            const double smu0 = sqrt(mu0), smu1 = sqrt(mu1);
            L.pk[1] = (mu0 + mu1) / 2.0;
            L.pk[2] = (smu0 + smu1) * (smu0 + smu1) / 2.0;
            L.pk[3] = mu0 * mu1;
            L.pk[4] = 2.0 * L.pk[3] / L.pk[2];
            L.pk[5] = (mu1 - 2.0 * smu0 * smu1 + mu0) / (mu1 + 2.0 * smu0 * smu1 + mu0);
            amg_level_poly_vectors(h, L);
            break;
        }
        default: break;
    }
}

// Coarsest level: dense inverse (the default up to coarse_dense_max rows) or, when the level is too large,
// option coarse_dense = 0, or the factorisation meets a vanishing pivot, CG in one cooperative kernel to the
// reference's coarse tolerance param->tol * 1e-4 (PreMGCycle.c:56, PreMGUtil.inl:37-58).
void amg_setup_coarse(Amg& h)
{
    Level& C = h.lv[h.nl - 1];
    h.coarse_iterative = true;
    if (ctx().opt.coarse_dense && C.n <= ctx().opt.coarse_dense_max) {
        if (dense_invert_csr(h.coarse, C.A)) {
            h.coarse_iterative = false;
            h.bytes += sizeof(double) * (size_t)C.n * C.n;
        }
    }
    if (h.coarse_iterative) {
        coarse_cg_setup(h.coarse_cg, C.A, h.tol * 1e-4);
        h.bytes += sizeof(double) * 3 * (size_t)C.n;
    }
}

Amg* amg_upload(AMG_data* mgl, AMG_param* param)
{
    ensure_init();
    if (mgl == nullptr || param == nullptr) fail(ERROR_INPUT_PAR, "amg_upload: null argument");
    const int nl = mgl[0].num_levels;
    if (nl < 1 || nl > MAX_AMG_LVL) fail(ERROR_DATA_STRUCTURE, "amg_upload: num_levels = %d", nl);
    if (mgl[0].ILU_levels > 0 || mgl[0].SWZ_levels > 0)
        fail(ERROR_AMG_SMOOTH_TYPE, "ILU / Schwarz smoothing levels are not on the device path");
    if (param->cycle_type == AMLI_CYCLE || param->cycle_type == NL_AMLI_CYCLE)
        fail(ERROR_INPUT_PAR, "AMLI cycles are not on the device path (cycle_type %d)",
             (int)param->cycle_type);
    if (nl > 1 && !smoother_supported(param->smoother))
        fail(ERROR_AMG_SMOOTH_TYPE,
             "smoother %d has no data-parallel device form (supported: Jacobi 1, poly 9, L1 10; "
             "GS 2 as multicolour GS with option gs_multicolor=1, as FASP's OpenMP build does)",
             (int)param->smoother);

    Amg* h = new Amg();
    try {
        amg_set_params(*h, param);
        h->nl = nl;
        h->lv.resize(nl);
        const bool ua = (param->AMG_type == UA_AMG);
        for (int l = 0; l < nl; ++l) {
            Level&         L = h->lv[l];
            const dCSRmat& A = mgl[l].A;
            csr_upload(L.A, A.row, A.col, A.nnz, A.IA, A.JA, A.val);
            L.n = A.row;
            if (l < nl - 1) {
                const dCSRmat& P = mgl[l].P;
                const dCSRmat& R = mgl[l].R;
                csr_upload(L.P, P.row, P.col, P.nnz, P.IA, P.JA, P.val, ua);
                csr_upload(L.R, R.row, R.col, R.nnz, R.IA, R.JA, R.val, ua);
                amg_level_smoother_data(*h, L, &A);
            }
            L.nglobal = L.n;
            amg_level_vectors(*h, L);
            h->bytes += L.A.bytes + L.P.bytes + L.R.bytes;
        }
        h->scal = dalloc<double>(4);
        FC_CUDA(cudaMemsetAsync(h->scal, 0, 4 * sizeof(double), ctx().stream));
        amg_setup_coarse(*h);
        FC_CUDA(cudaStreamSynchronize(ctx().stream));
    } catch (...) {
        amg_free(h);
        throw;
    }
    return h;
}

void amg_free(Amg* h)
{
    if (!h) return;
    for (Level& L : h->lv) {
        if (L.p2p_registered) {
            p2p_unregister(L.b), p2p_unregister(L.xa), p2p_unregister(L.xb), p2p_unregister(L.w);
            if (L.dscale_ext) p2p_unregister(L.dscale_ext);
            for (int i = 0; i < 3; ++i)
                if (L.pv[i]) p2p_unregister(L.pv[i]);
        }
        csr_free(L.A);
        csr_free(L.P);
        csr_free(L.R);
        dfree(L.b);
        dfree(L.xa);
        dfree(L.xb);
        dfree(L.w);
        for (int i = 0; i < 3; ++i) dfree(L.pv[i]);
        dfree(L.dscale_ext);
        dfree(L.color_rows);
        halo_free(L.hA);
        halo_free(L.hP);
        halo_free(L.hR);
    }
    dense_free(h->coarse);
    coarse_cg_free(h->coarse_cg);
    dfree(h->scal);
    delete h;
}

// ------------------------------------------------------------------------------------
// smoothing
// ------------------------------------------------------------------------------------
namespace {

struct CycleState {
    Amg&          h;
    const int*    done;
    const double* b0;       // level-0 right-hand side
    double*       x_out;    // where the final level-0 iterate must land
    Reduce        red;      // fused into the last kernel writing x_out when possible
    bool          red_done = false;
    int           gs_order = 1;   // +1 pre-smoothing, -1 post-smoothing
    std::vector<double*> cur;     // current iterate buffer per level
    std::vector<bool>    xzero;   // iterate known to be identically zero
    // multi-GPU, redundant ghost rows: the ghost entries (in A_l's ghost layout) of the level's right-hand side /
    // of the current iterate buffer are up to date, so the next gather of that vector needs no exchange
    std::vector<bool>    bfresh, xfresh;
    // the restriction that produced this level's right-hand side already wrote the zero-guess sweep x = s b / d
    // into other(l) (CsrArgs::div_out): the first pre-smoothing sweep only does its bookkeeping
    std::vector<bool>    prefused;
    CycleState(Amg& h_, const int* d)
        : h(h_), done(d), cur(h_.nl), xzero(h_.nl, false), bfresh(h_.nl, false), xfresh(h_.nl, false),
          prefused(h_.nl, false)
    {
    }
    const double* rhs(int l) const { return l == 0 ? b0 : h.lv[l].b; }
    double*       other(int l) const
    {
        Level& L = h.lv[l];
        return cur[l] == L.xa ? L.xb : L.xa;
    }
};

// nsweeps of the configured smoother on level l. `last` marks the final operation of the
// whole cycle: its output goes to x_out and carries the fused reduction.
void smooth(CycleState& s, int l, int nsweeps, bool last)
{
    Amg&          h = s.h;
    Level&        L = h.lv[l];
    const double* b = s.rhs(l);
    const size_t  n = L.n;
    for (int sw = 0; sw < nsweeps; ++sw) {
        const bool final_sweep = last && sw == nsweeps - 1;
        switch (h.smoother) {
            case SMOOTHER_JACOBI:
            case SMOOTHER_L1DIAG: {
                const bool jac = (h.smoother == SMOOTHER_JACOBI);
                double*    out = final_sweep ? s.x_out : s.other(l);
                Reduce     red = final_sweep ? s.red : Reduce();
                if (s.xzero[l] && ctx().opt.zero_guess) {
                    // with the ghost rows of b at hand (R computed them) the sweep covers them as well: the
                    // residual that follows then needs no exchange of x
                    const bool ext = s.bfresh[l] && L.dscale_ext != nullptr && !red.dot_out && !red.nrm2_out;
                    if (!s.prefused[l])
                        vec_scale_div(out, jac ? h.relax : 1.0, b, ext ? L.dscale_ext : (jac ? L.A.diag : L.A.l1),
                                      ext ? n + (size_t)L.A.nghost : n, red, s.done);
                    s.prefused[l] = false;
                    s.xfresh[l]   = ext;
                } else {
                    if (s.xzero[l]) {
                        vec_set(s.cur[l], 0.0, n, s.done);
                        s.xfresh[l] = false;
                    }
                    CsrArgs a;
                    a.mode      = jac ? CSR_JACOBI : CSR_L1;
                    a.alpha     = h.relax;
                    a.x         = s.cur[l];
                    a.b         = b;
                    a.y         = out;
                    a.red       = red;
                    a.done      = s.done;
                    a.skip_halo = s.xfresh[l];
                    csr_launch(L.A, a);
                    s.xfresh[l] = false;   // the sweep writes another buffer: owned rows only
                }
                s.bfresh[l] = false;
                if (final_sweep) s.red_done = true;
                s.cur[l]   = out;
                s.xzero[l] = false;
                break;
            }
            case SMOOTHER_GS: {
                // multicolour GS in place; pre-smoothing ascends the colours, post-smoothing
                // descends (PreMGCycle.c:126-127, 253-254)
                if (s.xzero[l]) vec_set(s.cur[l], 0.0, n, s.done);
                gs_multicolor_sweeps(L.A, L.color_rows, L.color_ptr, b, s.cur[l], 1, s.gs_order, s.done);
                s.xzero[l] = false;
                s.xfresh[l] = s.bfresh[l] = false;
                break;
            }
            case SMOOTHER_POLY: {
                // fasp_smoother_dcsr_poly, ItrSmootherCSRpoly.c:114-125 + Rr :551-609
                double*       u    = s.cur[l];
                const double* r    = L.w;
                double*       rbar = L.pv[0];
                if (s.xzero[l]) {
                    vec_set(u, 0.0, n, s.done);
                    if (ctx().opt.zero_guess) {
                        r = b;   // b - A*0
                        vec_mul(rbar, L.A.dinv, b, n, s.done);
                    }
                }
                if (s.xzero[l]) s.xfresh[l] = false;
                if (r == L.w) {
                    CsrArgs a;
                    a.mode      = CSR_RESID_DINV;
                    a.x         = u;
                    a.b         = b;
                    a.y         = L.w;
                    a.v0_out    = rbar;
                    a.done      = s.done;
                    a.skip_halo = s.xfresh[l];
                    csr_launch(L.A, a);
                }
                s.xfresh[l] = s.bfresh[l] = false;   // u += error below touches owned rows only
                double* v0 = L.pv[1];
                double* v1 = L.pv[2];
                {
                    CsrArgs a;
                    a.mode   = CSR_POLY1;
                    a.x      = rbar;
                    a.y      = v1;
                    a.v0_out = v0;
                    a.k1 = L.pk[1], a.k2 = L.pk[2], a.k3 = L.pk[3];
                    a.done = s.done;
                    csr_launch(L.A, a);
                }
                double* vn = rbar;   // rbar is dead after POLY1
                for (int j = 1; j < h.ndeg; ++j) {
                    CsrArgs a;
                    a.mode  = CSR_POLYJ;
                    a.x     = v1;
                    a.v0    = v0;
                    a.b     = r;
                    a.y     = vn;
                    a.k4 = L.pk[4], a.k5 = L.pk[5];
                    a.u_acc = (j == h.ndeg - 1) ? u : nullptr;   // u += error (:125)
                    a.done  = s.done;
                    csr_launch(L.A, a);
                    double* t = v0;
                    v0        = v1;
                    v1        = vn;
                    vn        = t;
                }
                s.xzero[l] = false;
                break;
            }
            default: fail(ERROR_AMG_SMOOTH_TYPE, "smoother %d not on the device path", (int)h.smoother);
        }
    }
}

void coarse_solve(CycleState& s, bool last)
{
    Amg& h      = s.h;
    const int l = h.nl - 1;
    double* out = last ? s.x_out : s.cur[l];
    if (h.coarse_iterative) coarse_cg_apply(h.coarse_cg, h.lv[l].A, s.rhs(l), out, s.done);
    else dense_apply(h.coarse, s.rhs(l), out, s.done);
    s.cur[l]   = out;
    s.xzero[l] = false;
}

// the cycle of PreMGCycle.c:95-269; level-0 iterate starts in s.cur[0]
void run_cycle(CycleState& s)
{
    Amg&      h  = s.h;
    const int nl = h.nl;
    int       num_lvl[MAX_AMG_LVL] = {0};
    int       ncycles[MAX_AMG_LVL];
    for (int i = 0; i < MAX_AMG_LVL; ++i) ncycles[i] = 1;
    switch (h.cycle_type) {
        case VW_CYCLE:
            for (int i = MAX_AMG_LVL - 2; i > 0; i -= 2) ncycles[i] = 2;
            break;
        case WV_CYCLE:
            for (int i = MAX_AMG_LVL - 1; i > 0; i -= 2) ncycles[i] = 2;
            break;
        default:
            for (int i = 0; i < MAX_AMG_LVL; ++i) ncycles[i] = h.cycle_type;
    }

    int l = 0;
    if (nl == 1) coarse_solve(s, true);
    while (nl > 1) {
        // ForwardSweep
        while (l < nl - 1) {
            Level& L = h.lv[l];
            num_lvl[l]++;
            s.gs_order = 1;
            smooth(s, l, h.presmooth, false);
            if (s.xzero[l]) {   // no pre-smoothing at all: x is still zero, residual = b
                vec_set(s.cur[l], 0.0, L.n, s.done);
                s.xzero[l] = false;
            }
            CsrArgs a;   // w = b - A x   (copy + aAxpy(-1), PreMGCycle.c:136-137)
            a.mode      = CSR_RESID;
            a.x         = s.cur[l];
            a.b         = s.rhs(l);
            a.y         = L.w;
            a.done      = s.done;
            a.skip_halo = s.xfresh[l];
            csr_launch(L.A, a);
            if (L.dist) s.xfresh[l] = true;   // exchanged or computed: the ghosts of this buffer are current
            CsrArgs r;   // b_{l+1} = R w  (:140-147; pattern-only for UA)
            r.mode = CSR_MXV;
            r.x    = L.w;
            r.y    = h.lv[l + 1].b;
            r.done = s.done;
            const bool gather_next = L.dist && !h.lv[l + 1].dist;   // agglomeration boundary
            if (gather_next) r.y += L.gdispls[comm_rank()];
            // Fuse the next level's zero-guess pre-smoothing sweep x = s b / d (a pass over b_{l+1} of its own) into
            // this kernel's epilogue: same operations on the same values, one launch and 8 B/row less per level.
            // Not across the agglomeration boundary (every rank needs the whole replicated vector).
            Level&     Ln   = h.lv[l + 1];
            const bool jl1  = (h.smoother == SMOOTHER_JACOBI || h.smoother == SMOOTHER_L1DIAG);
            // Default: in the multi-GPU solve only. On one GPU it is a wash (R_0 +32 us, the saved sweep 33 us: 60.9 vs
            // 59.8 ms per solve); across GPUs the solve is bound by the number of graph nodes and every launch counts.
            const bool want = ctx().opt.fuse_restrict == 2 || (ctx().opt.fuse_restrict == 1 && h.dist);
            bool       fuse = want && jl1 && ctx().opt.zero_guess && h.presmooth >= 1 && l + 1 < nl - 1 && !gather_next;
            if (fuse) {
                const bool jac  = (h.smoother == SMOOTHER_JACOBI);
                const bool extn = L.r_ext && Ln.dscale_ext != nullptr;   // R also computes the ghost rows of b_{l+1}
                if (L.r_ext && !extn) fuse = false;                      // extra rows without a divisor for them
                if (fuse) {
                    r.div_out = Ln.xb;   // = other(l+1) once cur[l+1] = xa (below)
                    r.div_d   = extn ? Ln.dscale_ext : (jac ? Ln.A.diag : Ln.A.l1);
                    r.div_s   = jac ? h.relax : 1.0;
                    if (!r.div_d) fuse = false, r.div_out = nullptr;
                    else r.mode = CSR_MXV_DIV;
                }
            }
            csr_launch(L.R, r);
            if (gather_next)
                comm_allgatherv(r.y, L.gcounts[comm_rank()], h.lv[l + 1].b, L.gcounts, L.gdispls, s.done);
            ++l;
            s.cur[l]    = h.lv[l].xa;
            s.xzero[l]  = true;   // fasp_dvec_set(..., 0.0) (:151) is folded into the next writer
            s.xfresh[l] = false;
            s.bfresh[l] = h.lv[l - 1].r_ext;   // R also produced the ghost rows of this level's right-hand side
            s.prefused[l] = fuse;
        }

        coarse_solve(s, false);

        // BackwardSweep
        while (l > 0) {
            --l;
            Level& L  = h.lv[l];
            Level& Lc = h.lv[l + 1];
            CsrArgs p;   // x_l += alpha P x_{l+1}   (:219-227)
            p.mode = CSR_AXPY;
            p.x    = s.cur[l + 1];
            p.y    = s.cur[l];
            p.done = s.done;
            if (h.coarse_scaling == ON) {   // (:210-216)
                vec_dot(s.cur[l + 1], Lc.b, Lc.n, h.scal + 1, s.done);
                if (Lc.dist) comm_allreduce(h.scal + 1, 1, 0, s.done);
                CsrArgs v;
                v.mode         = CSR_MXV;
                v.x            = s.cur[l + 1];
                v.y            = Lc.w;
                v.red.dot_with = s.cur[l + 1];
                v.red.dot_out  = h.scal + 2;
                v.red.global   = Lc.dist;
                v.done         = s.done;
                csr_launch(Lc.A, v);
                scaling_alpha(h.scal, s.done);
                p.alpha_dev = h.scal;
            }
            csr_launch(L.P, p);
            // P with ghost rows appended also updated the ghost entries of x_l (they were current before)
            s.xfresh[l] = L.p_ext && s.xfresh[l];
            // the cycle ends when the backward sweep reaches level 0
            const bool last_level_visit = (l == 0);
            s.gs_order = -1;
            smooth(s, l, h.postsmooth, last_level_visit);
            if (num_lvl[l] < ncycles[l]) break;
            num_lvl[l] = 0;
        }
        if (l == 0) break;
    }

    // result placement + reduction when the last kernel could not do it
    if (s.cur[0] != s.x_out) {
        vec_copy(s.x_out, s.cur[0], h.lv[0].n, s.done);
        s.cur[0] = s.x_out;
    }
    if (!s.red_done) vec_reduce(s.x_out, h.lv[0].n, s.red, s.done);
}

} // namespace

// nsweeps of the level-0 smoother on (b, u), u in/out: the host-pointer smoother drop-ins
void amg_smooth_only(Amg& h, const double* b, double* u, int nsweeps)
{
    CycleState s(h, nullptr);
    s.b0       = b;
    s.x_out    = u;
    s.cur[0]   = u;
    s.xzero[0] = false;
    smooth(s, 0, nsweeps, false);
    if (s.cur[0] != u) vec_copy(u, s.cur[0], h.lv[0].n, nullptr);
}

void amg_apply(Amg& h, const double* r, double* z, const Reduce& red, const int* done)
{
    CycleState s(h, done);
    s.b0    = r;
    s.x_out = z;
    const int ncyc = h.maxit < 1 ? 1 : h.maxit;
    for (int c = 0; c < ncyc; ++c) {
        s.red      = (c == ncyc - 1) ? red : Reduce();
        s.red_done = false;
        if (c == 0) {
            s.cur[0]   = h.lv[0].xa;
            s.xzero[0] = true;
            if (h.nl == 1) s.xzero[0] = false;
        } else {
            // continue from the previous cycle's iterate, which sits in z: move it to a
            // work buffer so the final sweep can write z without aliasing its own input
            vec_copy(h.lv[0].xa, z, h.lv[0].n, done);
            s.cur[0]   = h.lv[0].xa;
            s.xzero[0] = false;
        }
        run_cycle(s);
    }
}

void amg_cycle_inplace(Amg& h, const double* b, double* x, bool x_is_zero, const Reduce& red,
                       const int* done)
{
    CycleState s(h, done);
    s.b0    = b;
    s.x_out = x;
    s.red   = red;
    if (x_is_zero) {
        s.cur[0]   = h.lv[0].xa;
        s.xzero[0] = (h.nl > 1);
    } else {
        vec_copy(h.lv[0].xa, x, h.lv[0].n, done);
        s.cur[0]   = h.lv[0].xa;
        s.xzero[0] = false;
    }
    run_cycle(s);
}

} // namespace fc
