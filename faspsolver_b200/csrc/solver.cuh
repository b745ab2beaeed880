// solver.cuh — "hierarchy + Krylov" solver objects behind fasp_cuda_krylov_amg_* and the
// run-time binding to the host application's FASP setup routines.
#pragma once
#include "common.cuh"
#include "amg.cuh"
#include "krylov.cuh"
#include "dist.cuh"

struct fasp_cuda_solver_s {
    fc::Amg*       amg  = nullptr;     // CSR hierarchy (level-0 A doubles as the Krylov operator)
    fc::BAmg*      bamg = nullptr;     // or a BSR hierarchy
    fc::SolveStats stats;
    fc::PcgCache   pcg_cache;      // workspace + graphs reused across solves
    fc::GmresCache gmres_cache;    // the same for GMRES / vGMRES / vFGMRES
    double         ms_total = 0.0;     // last solve incl. H2D/D2H
    double*        d_b = nullptr;      // staging vectors for host-pointer solves
    double*        d_x = nullptr;
    double*        pin = nullptr;      // pinned host staging (2n)
    size_t         n   = 0;
    bool           x_registered = false;   // d_x is peer-mapped (multi-GPU)
    // caller buffers page-locked in place (cudaHostRegister) and remembered across solves
    struct HostReg { const void* p; size_t bytes; };
    std::vector<HostReg> hostreg;
};

namespace fc {

// FASP host routines used for the setup phase (never for the solve), resolved with dlsym
struct HostFasp {
    typedef AMG_data* (*create_t)(SHORT);
    typedef void (*free_t)(AMG_data*, AMG_param*);
    typedef SHORT (*setup_t)(AMG_data*, AMG_param*);
    typedef dCSRmat (*csrcreate_t)(const INT, const INT, const INT);
    typedef void (*csrcp_t)(const dCSRmat*, dCSRmat*);
    typedef dvector (*dveccreate_t)(const INT);
    typedef AMG_data_bsr* (*bcreate_t)(SHORT);
    typedef void (*bfree_t)(AMG_data_bsr*, AMG_param*);
    typedef SHORT (*bsetup_t)(AMG_data_bsr*, AMG_param*);
    typedef dBSRmat (*bsrcreate_t)(const INT, const INT, const INT, const INT, const INT);
    typedef void (*bsrcp_t)(const dBSRmat*, dBSRmat*);
    bool         tried = false, ok = false;
    create_t     amg_data_create = nullptr;
    free_t       amg_data_free   = nullptr;
    setup_t      setup_rs = nullptr, setup_sa = nullptr, setup_ua = nullptr;
    csrcreate_t  dcsr_create = nullptr;
    csrcp_t      dcsr_cp     = nullptr;
    dveccreate_t dvec_create = nullptr;
    bcreate_t    amg_data_bsr_create = nullptr;
    bfree_t      amg_data_bsr_free   = nullptr;
    bsetup_t     setup_sa_bsr = nullptr, setup_ua_bsr = nullptr;
    bsrcreate_t  dbsr_create = nullptr;
    bsrcp_t      dbsr_cp     = nullptr;
};
HostFasp& host_fasp();
void      require_host_fasp();

fasp_cuda_solver_s* solver_create_csr(AMG_data* mgl, AMG_param* amgparam);
fasp_cuda_solver_s* solver_create_bsr(AMG_data_bsr* mgl, AMG_param* amgparam);
fasp_cuda_solver_s* solver_create_dist(AMG_data* mgl, AMG_param* amgparam, int agg_rows);
fasp_cuda_solver_s* solver_create_dist_slabs(int nlev, const fasp_cuda_slab_level* sl, const int* tail_off,
                                             AMG_data* tail, AMG_param* amgparam);
void                solver_destroy(fasp_cuda_solver_s* s);
int    solver_solve_dev(fasp_cuda_solver_s* s, const double* b_dev, double* x_dev, ITS_param* it);
int    solver_solve_host(fasp_cuda_solver_s* s, const double* b, double* x, ITS_param* it);
double solver_stat(const fasp_cuda_solver_s* s, int what);
int    solver_history(const fasp_cuda_solver_s* s, double* relres, int max_entries);
int    solver_amg_solve(AMG_data* mgl, AMG_param* param);

void amg_smooth_only(Amg& h, const double* b, double* u, int nsweeps);

} // namespace fc
