// bamg.cu — BSR AMG hierarchy upload and multigrid cycle on the device.
//
// Replaces fasp_solver_mgcycle_bsr (PreMGCycle.c:287-566) and fasp_precond_dbsr_amg
// (PreBSR.c:1149) for the block-Jacobi smoother (the only data-parallel smoother the BSR
// cycle offers, PreMGCycle.c:328-365). The hierarchy (A_l, P_l, R_l as BSR with identity
// transfer blocks, diaginv_l) comes from FASP's host setup fasp_amg_setup_ua_bsr /
// fasp_amg_setup_sa_bsr. Both pre- and post-smoothing run `presmooth_iter` sweeps, as the
// reference does (PreMGCycle.c:297,497). The coarsest level is solved with a dense inverse
// instead of the reference's unpreconditioned inner GMRES (PreMGCycle.c:443-459).
#include "amg.cuh"

namespace fc {

BAmg* bamg_upload(AMG_data_bsr* mgl, AMG_param* param)
{
    ensure_init();
    if (!mgl || !param) fail(ERROR_INPUT_PAR, "bamg_upload: null argument");
    const int nl = mgl[0].num_levels;
    if (nl < 1 || nl > MAX_AMG_LVL) fail(ERROR_DATA_STRUCTURE, "bamg_upload: num_levels = %d", nl);
    if (mgl[0].ILU_levels > 0) fail(ERROR_AMG_SMOOTH_TYPE, "ILU smoothing levels are not on the device path");
    if (mgl[0].A_nk != nullptr) fail(ERROR_INPUT_PAR, "near-kernel correction (A_nk) is not on the device path");
    if (nl > 1 && param->smoother != SMOOTHER_JACOBI)
        fail(ERROR_AMG_SMOOTH_TYPE, "BSR cycle: only the block-Jacobi smoother (1) is data-parallel, got %d",
             (int)param->smoother);
    if (param->cycle_type != V_CYCLE && param->cycle_type != W_CYCLE)
        fail(ERROR_INPUT_PAR, "BSR cycle: cycle_type %d is not on the device path", (int)param->cycle_type);

    BAmg* h = new BAmg();
    try {
        h->nl = nl, h->nb = mgl[0].A.nb;
        h->smoother = param->smoother, h->cycle_type = param->cycle_type;
        h->presmooth = param->presmooth_iter, h->postsmooth = param->presmooth_iter;
        h->relax = param->relaxation, h->tol = param->tol, h->maxit = param->maxit;
        h->coarse_scaling = param->coarse_scaling;
        h->scal = dalloc<double>(4);
        FC_CUDA(cudaMemsetAsync(h->scal, 0, 4 * sizeof(double), ctx().stream));
        h->lv.resize(nl);
        for (int l = 0; l < nl; ++l) {
            BLevel&        L = h->lv[l];
            const dBSRmat& A = mgl[l].A;
            bsr_upload(L.A, A.ROW, A.COL, A.NNZ, A.nb, A.IA, A.JA, A.val);
            L.n = A.ROW * A.nb;
            if (l < nl - 1) {
                const dBSRmat& P = mgl[l].P;
                const dBSRmat& R = mgl[l].R;
                // UA-AMG: identity transfer blocks -> pattern-only gather / segmented sum (4 B per block)
                bsr_upload(L.P, P.ROW, P.COL, P.NNZ, P.nb, P.IA, P.JA, P.val, true);
                bsr_upload(L.R, R.ROW, R.COL, R.NNZ, R.nb, R.IA, R.JA, R.val, true);
                if (h->coarse_scaling == ON) {
                    L.peh = dalloc<double>((size_t)A.ROW * A.nb);
                    L.aeh = dalloc<double>((size_t)A.ROW * A.nb);
                    h->bytes += 2 * sizeof(double) * (size_t)A.ROW * A.nb;
                }
                const size_t nd = (size_t)A.ROW * A.nb * A.nb;
                if (mgl[l].diaginv.val == nullptr || (size_t)mgl[l].diaginv.row < nd)
                    fail(ERROR_DATA_STRUCTURE, "level %d has no diaginv (host setup incomplete)", l);
                L.diaginv = dalloc<double>(nd);
                FC_CUDA(cudaMemcpyAsync(L.diaginv, mgl[l].diaginv.val, sizeof(double) * nd,
                                        cudaMemcpyHostToDevice, ctx().stream));
                h->bytes += sizeof(double) * nd;
            }
            L.b  = dalloc<double>(L.n);
            L.xa = dalloc<double>(L.n);
            L.xb = dalloc<double>(L.n);
            L.w  = dalloc<double>(L.n);
            h->bytes += L.A.bytes + L.P.bytes + L.R.bytes + 4 * sizeof(double) * (size_t)L.n;
        }
        BLevel& C = h->lv[nl - 1];
        if (C.n > ctx().opt.coarse_dense_max)
            fail(ERROR_AMG_SETUP, "coarsest BSR level has %d unknowns > coarse_dense_max = %d", C.n,
                 ctx().opt.coarse_dense_max);
        if (!dense_invert_bsr(h->coarse, C.A))
            fail(ERROR_AMG_SETUP, "coarsest BSR level is numerically singular (pivot below 1e-14 of the largest entry)");
        h->bytes += sizeof(double) * (size_t)C.n * C.n;
        FC_CUDA(cudaStreamSynchronize(ctx().stream));
    } catch (...) {
        bamg_free(h);
        throw;
    }
    return h;
}

void bamg_free(BAmg* h)
{
    if (!h) return;
    for (BLevel& L : h->lv) {
        bsr_free(L.A);
        bsr_free(L.P);
        bsr_free(L.R);
        dfree(L.b);
        dfree(L.xa);
        dfree(L.xb);
        dfree(L.w);
        dfree(L.diaginv);
        dfree(L.peh);
        dfree(L.aeh);
    }
    dense_free(h->coarse);
    dfree(h->scal);
    delete h;
}

namespace {

struct BCycle {
    BAmg&         h;
    const int*    done;
    const double* b0;
    double*       x_out;
    Reduce        red;
    bool          red_done = false;
    std::vector<double*> cur;
    std::vector<bool>    xzero;
    BCycle(BAmg& h_, const int* d) : h(h_), done(d), cur(h_.nl), xzero(h_.nl, false) {}
    const double* rhs(int l) const { return l == 0 ? b0 : h.lv[l].b; }
    double*       other(int l) const { return cur[l] == h.lv[l].xa ? h.lv[l].xb : h.lv[l].xa; }
};

void bsmooth(BCycle& s, int l, int nsweeps, bool last)
{
    BLevel& L = s.h.lv[l];
    for (int sw = 0; sw < nsweeps; ++sw) {
        const bool final_sweep = last && sw == nsweeps - 1;
        double*    out         = final_sweep ? s.x_out : s.other(l);
        BsrArgs a;
        if (s.xzero[l] && ctx().opt.zero_guess) {
            // first sweep from x = 0: b - sum A_IJ 0 == b exactly, so u_I = Dinv_I b_I without a pass over A
            a.mode    = BSR_DINV;
            a.b       = s.rhs(l);
            a.y       = out;
            a.diaginv = L.diaginv;
            a.red     = final_sweep ? s.red : Reduce();
            a.done    = s.done;
            bsr_launch(L.A, a);
            if (final_sweep) s.red_done = true;
            s.cur[l]   = out;
            s.xzero[l] = false;
            continue;
        }
        if (s.xzero[l]) {
            vec_set(s.cur[l], 0.0, L.n, s.done);
            s.xzero[l] = false;
        }
        a.mode    = BSR_JACOBI;
        a.x       = s.cur[l];
        a.b       = s.rhs(l);
        a.y       = out;
        a.diaginv = L.diaginv;
        a.red     = final_sweep ? s.red : Reduce();
        a.done    = s.done;
        bsr_launch(L.A, a);
        if (final_sweep) s.red_done = true;
        s.cur[l] = out;
    }
}

void brun_cycle(BCycle& s)
{
    BAmg&     h  = s.h;
    const int nl = h.nl;
    int       nu_l[MAX_AMG_LVL] = {0};
    int       l = 0;
    if (nl == 1) {
        dense_apply(h.coarse, s.rhs(0), s.x_out, s.done);
        s.cur[0] = s.x_out;
    }
    while (nl > 1) {
        while (l < nl - 1) {   // ForwardSweep (PreMGCycle.c:319-403)
            BLevel& L = h.lv[l];
            nu_l[l]++;
            bsmooth(s, l, h.presmooth, false);
            if (s.xzero[l]) {
                vec_set(s.cur[l], 0.0, L.n, s.done);
                s.xzero[l] = false;
            }
            BsrArgs a;
            a.mode = BSR_RESID;
            a.x    = s.cur[l];
            a.b    = s.rhs(l);
            a.y    = L.w;
            a.done = s.done;
            bsr_launch(L.A, a);
            BsrArgs r;
            r.mode = BSR_MXV;
            r.x    = L.w;
            r.y    = h.lv[l + 1].b;
            r.done = s.done;
            bsr_launch(L.R, r);
            ++l;
            s.cur[l]   = h.lv[l].xa;
            s.xzero[l] = true;
        }
        dense_apply(h.coarse, s.rhs(l), s.cur[l], s.done);
        s.xzero[l] = false;
        while (l > 0) {   // BackwardSweep (:462-560)
            --l;
            BLevel& L = h.lv[l];
            if (h.coarse_scaling == ON) {
                // alpha = min((A P e, w) / (A P e, A P e), 1) ; x += alpha P e   (PreMGCycle.c:465-480)
                BsrArgs pe;
                pe.mode = BSR_MXV;
                pe.x    = s.cur[l + 1];
                pe.y    = L.peh;
                pe.done = s.done;
                bsr_launch(L.P, pe);
                BsrArgs ae;
                ae.mode         = BSR_MXV;
                ae.x            = L.peh;
                ae.y            = L.aeh;
                ae.red.dot_with = L.w;
                ae.red.dot_out  = h.scal + 1;
                ae.red.nrm2_out = h.scal + 2;
                ae.done         = s.done;
                bsr_launch(L.A, ae);
                scaling_alpha(h.scal, s.done);
                vec_axpy_dev(h.scal, L.peh, s.cur[l], L.n, s.done);
            } else {
                BsrArgs p;
                p.mode  = BSR_AXPY;
                p.alpha = 1.0;
                p.x     = s.cur[l + 1];
                p.y     = s.cur[l];
                p.done  = s.done;
                bsr_launch(L.P, p);
            }
            bsmooth(s, l, h.postsmooth, l == 0);
            if (nu_l[l] < h.cycle_type) break;
            nu_l[l] = 0;
        }
        if (l == 0) break;
    }
    if (s.cur[0] != s.x_out) {
        vec_copy(s.x_out, s.cur[0], h.lv[0].n, s.done);
        s.cur[0] = s.x_out;
    }
    if (!s.red_done) vec_reduce(s.x_out, h.lv[0].n, s.red, s.done);
}

} // namespace

void bamg_apply(BAmg& h, const double* r, double* z, const Reduce& red, const int* done)
{
    BCycle s(h, done);
    s.b0    = r;
    s.x_out = z;
    const int ncyc = h.maxit < 1 ? 1 : h.maxit;
    for (int c = 0; c < ncyc; ++c) {
        s.red      = (c == ncyc - 1) ? red : Reduce();
        s.red_done = false;
        if (c == 0) {
            s.cur[0]   = h.lv[0].xa;
            s.xzero[0] = (h.nl > 1);
        } else {
            vec_copy(h.lv[0].xa, z, h.lv[0].n, done);
            s.cur[0]   = h.lv[0].xa;
            s.xzero[0] = false;
        }
        brun_cycle(s);
    }
}

void bamg_cycle_inplace(BAmg& h, const double* b, double* x, bool x_is_zero, const Reduce& red,
                        const int* done)
{
    BCycle s(h, done);
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
    brun_cycle(s);
}

} // namespace fc
