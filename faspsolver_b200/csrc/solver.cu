// solver.cu — solver objects: an uploaded hierarchy plus the device Krylov loop.
// Mirrors the wiring of fasp_solver_dcsr_krylov_amg (SolCSR.c:476-569) and
// fasp_solver_dcsr_itsolver (SolCSR.c:56-140) with the setup phase factored out, and the
// AMG-as-solver loop fasp_amg_solve (PreMGSolve.c:49-135).
#include "solver.cuh"
#include "reduce.cuh"
#include "p2p.cuh"
#include "comm.cuh"
#include <thread>
#include <chrono>

namespace fc {

// pageable caller memory <-> pinned staging, split over a few host threads (a single-threaded
// memcpy of a 134 MB vector costs more than its PCIe transfer)
static void par_memcpy(void* dst, const void* src, size_t bytes)
{
    const size_t chunk = (size_t)8 << 20;
    unsigned     nt    = std::thread::hardware_concurrency();
    if (nt > 8) nt = 8;
    if (nt < 2 || bytes < 2 * chunk) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (bytes / nt + 63) & ~(size_t)63;
    for (unsigned t = 0; t < nt; ++t) {
        const size_t off = (size_t)t * per;
        if (off >= bytes) break;
        const size_t len = (off + per < bytes) ? per : bytes - off;
        th.emplace_back([=]() { memcpy((char*)dst + off, (const char*)src + off, len); });
    }
    for (auto& t : th) t.join();
}

fasp_cuda_solver_s* solver_create_csr(AMG_data* mgl, AMG_param* amgparam)
{
    ensure_init();
    fasp_cuda_solver_s* s = new fasp_cuda_solver_s();
    try {
        AMG_param p = *amgparam;
        p.tol       = 1e-6;   // fasp_precond_amg re-initialises tol (PreCSR.c:425, AuxParam.c:437)
        s->amg      = amg_upload(mgl, &p);
        s->n        = (size_t)s->amg->lv[0].n;
        s->d_b      = dalloc<double>(s->n);
        s->d_x      = dalloc<double>(s->n);
        FC_CUDA(cudaMallocHost(&s->pin, sizeof(double) * 2 * s->n));
    } catch (...) {
        solver_destroy(s);
        throw;
    }
    return s;
}

fasp_cuda_solver_s* solver_create_dist(AMG_data* mgl, AMG_param* amgparam, int agg_rows)
{
    ensure_init();
    fasp_cuda_solver_s* s = new fasp_cuda_solver_s();
    try {
        AMG_param p = *amgparam;
        p.tol       = 1e-6;
        s->amg      = dist_amg_upload(mgl, &p, agg_rows);
        s->n        = (size_t)s->amg->lv[0].n;                 // local rows
        const size_t cap = (size_t)s->amg->lv[0].cap + 8;      // + ghosts of the level-0 operator
        s->d_b      = dalloc<double>(cap);
        s->d_x      = dalloc<double>(cap);
        if (p2p_active()) {   // x is gathered by the peers when the true residual is formed
            p2p_register(s->d_x, sizeof(double) * cap);
            s->x_registered = true;
        }
        FC_CUDA(cudaMallocHost(&s->pin, sizeof(double) * 2 * s->n));
    } catch (...) {
        solver_destroy(s);
        throw;
    }
    return s;
}

fasp_cuda_solver_s* solver_create_dist_slabs(int nlev, const fasp_cuda_slab_level* sl, const int* tail_off,
                                             AMG_data* tail, AMG_param* amgparam)
{
    ensure_init();
    fasp_cuda_solver_s* s = new fasp_cuda_solver_s();
    try {
        AMG_param p = *amgparam;
        p.tol       = 1e-6;
        s->amg      = dist_amg_upload_slabs(nlev, sl, tail_off, tail, &p);
        s->n        = (size_t)s->amg->lv[0].n;
        const size_t cap = (size_t)s->amg->lv[0].cap + 8;
        s->d_b      = dalloc<double>(cap);
        s->d_x      = dalloc<double>(cap);
        if (p2p_active()) {
            p2p_register(s->d_x, sizeof(double) * cap);
            s->x_registered = true;
        }
        FC_CUDA(cudaMallocHost(&s->pin, sizeof(double) * 2 * s->n));
    } catch (...) {
        solver_destroy(s);
        throw;
    }
    return s;
}

fasp_cuda_solver_s* solver_create_bsr(AMG_data_bsr* mgl, AMG_param* amgparam)
{
    ensure_init();
    fasp_cuda_solver_s* s = new fasp_cuda_solver_s();
    try {
        AMG_param p = *amgparam;
        p.tol       = 1e-6;   // fasp_precond_dbsr_amg re-initialises the parameters (PreBSR.c:1159-1169)
        s->bamg     = bamg_upload(mgl, &p);
        s->n        = (size_t)s->bamg->lv[0].n;
        s->d_b      = dalloc<double>(s->n);
        s->d_x      = dalloc<double>(s->n);
        FC_CUDA(cudaMallocHost(&s->pin, sizeof(double) * 2 * s->n));
    } catch (...) {
        solver_destroy(s);
        throw;
    }
    return s;
}

void solver_destroy(fasp_cuda_solver_s* s)
{
    if (!s) return;
    for (auto& r : s->hostreg) {
        cudaHostUnregister(const_cast<void*>(r.p));
        cudaGetLastError();
    }
    s->hostreg.clear();
    s->pcg_cache.release();
    s->gmres_cache.release();
    if (s->x_registered) p2p_unregister(s->d_x);
    amg_free(s->amg);
    bamg_free(s->bamg);
    dfree(s->d_b);
    dfree(s->d_x);
    if (s->pin) cudaFreeHost(s->pin);
    delete s;
}

static int run_krylov(fasp_cuda_solver_s* s, LinOp& op, Prec& pc, const double* b_dev, double* x_dev,
                      ITS_param* it)
{
    switch (it->itsolver_type) {
        case SOLVER_CG:
            return pcg_solve(op, b_dev, x_dev, pc, it->tol, it->abstol, it->maxit, it->stop_type,
                             it->print_level, &s->stats, &s->pcg_cache);
        case SOLVER_GMRES:
            return gmres_solve(op, b_dev, x_dev, pc, it->tol, it->abstol, it->maxit, it->restart,
                               it->stop_type, it->print_level, GM_FIXED, &s->stats, &s->gmres_cache);
        case SOLVER_VGMRES:
            return gmres_solve(op, b_dev, x_dev, pc, it->tol, it->abstol, it->maxit, it->restart,
                               it->stop_type, it->print_level, GM_VARIABLE, &s->stats, &s->gmres_cache);
        case SOLVER_VFGMRES:
            return gmres_solve(op, b_dev, x_dev, pc, it->tol, it->abstol, it->maxit, it->restart,
                               it->stop_type, it->print_level, GM_FLEXIBLE, &s->stats, &s->gmres_cache);
        default:
            fail(ERROR_SOLVER_TYPE,
                 "itsolver_type %d not on the device path (supported: CG 1, GMRES 4, VGMRES 5, VFGMRES 6)",
                 (int)it->itsolver_type);
    }
}

int solver_solve_dev(fasp_cuda_solver_s* s, const double* b_dev, double* x_dev, ITS_param* it)
{
    if (!s || (!s->amg && !s->bamg)) fail(ERROR_INPUT_PAR, "null solver");
    if (s->amg) {
        CsrOp     op(&s->amg->lv[0].A);
        AmgPrec   pc(s->amg);
        // Row-partitioned hierarchy: x is gathered by the level-0 kernels (true residual), so it needs
        // room for the ghost entries behind its owned part and must be peer-mapped. A caller vector of
        // exactly n_local doubles has neither: stage it through the solver's own x buffer.
        const bool stage_x = s->amg->dist && x_dev != s->d_x;
        double*    x_run   = stage_x ? s->d_x : x_dev;
        if (stage_x)
            FC_CUDA(cudaMemcpyAsync(s->d_x, x_dev, sizeof(double) * s->n, cudaMemcpyDeviceToDevice, ctx().stream));
        const int ret = run_krylov(s, op, pc, b_dev, x_run, it);
        if (stage_x)
            FC_CUDA(cudaMemcpyAsync(x_dev, s->d_x, sizeof(double) * s->n, cudaMemcpyDeviceToDevice, ctx().stream));
        if (p2p_active() && p2p_error())
            fail(ERROR_SOLVER_MISC, "multi-GPU barrier timed out (a rank fell out of step)");
        return ret;
    }
    BsrOp    op(&s->bamg->lv[0].A);
    BAmgPrec pc(s->bamg);
    return run_krylov(s, op, pc, b_dev, x_dev, it);
}

int solver_solve_host(fasp_cuda_solver_s* s, const double* b, double* x, ITS_param* it)
{
    if (!s || (!s->amg && !s->bamg)) fail(ERROR_INPUT_PAR, "null solver");
    Ctx&         c  = ctx();
    const size_t n  = s->n;
    const auto   w0 = std::chrono::steady_clock::now();
    // A caller buffer that is already page-locked (fasp_cuda_host_pin / cudaHostRegister /
    // cudaMallocHost: the application solves many right-hand sides with the same arrays) is copied
    // by plain DMA. Anything else goes through the library's pinned staging area (threaded memcpy
    // overlapping the DMA). host_register=2 makes the library page-lock unknown buffers itself and
    // REMEMBER them: only for callers that promise not to free or remap those arrays while the
    // solver lives (a stale registration would DMA into pages the process no longer sees).
    const size_t bytes = sizeof(double) * n;
    auto pinned = [&](const void* p) -> bool {
        // page-locked from the first to the last byte? (an application may have pinned only a slice of the array)
        cudaPointerAttributes at, at2;
        if (cudaPointerGetAttributes(&at, p) == cudaSuccess &&
            cudaPointerGetAttributes(&at2, static_cast<const char*>(p) + bytes - 1) == cudaSuccess) {
            if (at.type == cudaMemoryTypeHost && at2.type == cudaMemoryTypeHost) return true;
        } else {
            cudaGetLastError();
        }
        if (ctx().opt.host_register < 2) return false;
        for (auto& r : s->hostreg)
            if (r.p == p && r.bytes >= bytes) return true;
        if (s->hostreg.size() >= 8) {
            cudaHostUnregister(const_cast<void*>(s->hostreg.front().p));
            cudaGetLastError();
            s->hostreg.erase(s->hostreg.begin());
        }
        if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        s->hostreg.push_back({p, bytes});
        return true;
    };
    const bool pb = pinned(b), px = pinned(x);
    if (pb) {
        FC_CUDA(cudaMemcpyAsync(s->d_b, b, bytes, cudaMemcpyHostToDevice, c.stream));
    } else {
        par_memcpy(s->pin, b, bytes);
        FC_CUDA(cudaMemcpyAsync(s->d_b, s->pin, bytes, cudaMemcpyHostToDevice, c.stream));
    }
    if (px) {
        FC_CUDA(cudaMemcpyAsync(s->d_x, x, bytes, cudaMemcpyHostToDevice, c.stream));
    } else {
        par_memcpy(s->pin + n, x, bytes);   // overlaps the DMA of b
        FC_CUDA(cudaMemcpyAsync(s->d_x, s->pin + n, bytes, cudaMemcpyHostToDevice, c.stream));
    }
    const int ret = solver_solve_dev(s, s->d_b, s->d_x, it);
    if (px) {
        FC_CUDA(cudaMemcpyAsync(x, s->d_x, bytes, cudaMemcpyDeviceToHost, c.stream));
        FC_CUDA(cudaStreamSynchronize(c.stream));
    } else {
        FC_CUDA(cudaMemcpyAsync(s->pin + n, s->d_x, bytes, cudaMemcpyDeviceToHost, c.stream));
        FC_CUDA(cudaStreamSynchronize(c.stream));
        par_memcpy(x, s->pin + n, bytes);
    }
    // wall clock of the whole host-pointer call: staging, H2D, solve, D2H, un-staging
    s->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    return ret;
}

double solver_stat(const fasp_cuda_solver_s* s, int what)
{
    if (!s) return -1.0;
    switch (what) {
        case 0: return (double)s->stats.iters;
        case 1: return s->stats.relres;
        case 2: return s->stats.ms;
        case 3: return (double)s->stats.launches;
        case 4: return s->ms_total;
        case 5: return s->amg ? (double)s->amg->bytes : (s->bamg ? (double)s->bamg->bytes : 0.0);
        default: return -1.0;
    }
}

int solver_history(const fasp_cuda_solver_s* s, double* relres, int max_entries)
{
    if (!s || !relres) return 0;
    int n = (int)s->stats.hist_relres.size();
    if (n > max_entries) n = max_entries;
    for (int i = 0; i < n; ++i) relres[i] = s->stats.hist_relres[i];
    return n;
}

void amg_smooth_only(Amg& h, const double* b, double* u, int nsweeps);

// fasp_amg_solve (PreMGSolve.c:49-135): cycles until ||b - A x|| / ||b|| < tol
int solver_amg_solve(AMG_data* mgl, AMG_param* param)
{
    ensure_init();
    Ctx&      c      = ctx();
    const int MaxIt  = param->maxit;
    const int prtlvl = param->print_level;
    Amg*      h      = amg_upload(mgl, param);
    int       iter   = 0;
    double    relres1 = 1.0;
    double*   work   = nullptr;
    double*   scal   = nullptr;
    try {
        const size_t n = h->lv[0].n;
        work           = dalloc<double>(3 * n);
        scal           = dalloc<double>(2);
        double *b = work, *x = work + n, *r = work + 2 * n;
        FC_CUDA(cudaMemcpyAsync(b, mgl[0].b.val, sizeof(double) * n, cudaMemcpyHostToDevice, c.stream));
        FC_CUDA(cudaMemcpyAsync(x, mgl[0].x.val, sizeof(double) * n, cudaMemcpyHostToDevice, c.stream));
        const double sumb    = vec_norm2_host(b, n);
        double       absres0 = sumb, absres, factor;
        print_itinfo(prtlvl, STOP_REL_RES, iter, relres1, sumb, 0.0);
        if (sumb <= SMALLREAL) vec_set(x, 0.0, n);
        CsrOp op(&h->lv[0].A);
        while ((iter++ < MaxIt) & (sumb > SMALLREAL)) {
            amg_cycle_inplace(*h, b, x, false, Reduce(), nullptr);
            Reduce red;
            red.nrm2_out = scal;
            op.apply(CSR_RESID, 1.0, x, b, r, red, nullptr);
            double rr = 0.0;
            FC_CUDA(cudaMemcpyAsync(&rr, scal, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
            FC_CUDA(cudaStreamSynchronize(c.stream));
            absres  = sqrt(rr);
            relres1 = absres / fmax(SMALLREAL, sumb);
            factor  = absres / absres0;
            absres0 = absres;
            print_itinfo(prtlvl, STOP_REL_RES, iter, relres1, absres, factor);
            if (relres1 < param->tol) break;
        }
        FC_CUDA(cudaMemcpyAsync(mgl[0].x.val, x, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
        FC_CUDA(cudaStreamSynchronize(c.stream));
        if (prtlvl > PRINT_NONE) print_final(iter, MaxIt, relres1);
    } catch (...) {
        dfree(work);
        dfree(scal);
        amg_free(h);
        throw;
    }
    dfree(work);
    dfree(scal);
    amg_free(h);
    return (iter > MaxIt) ? ERROR_SOLVER_MAXIT : iter;
}

} // namespace fc
