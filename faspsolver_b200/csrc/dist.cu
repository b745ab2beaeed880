// dist.cu — row-partitioned multi-GPU AMG hierarchy.
//
// One process per GPU. Every rank holds the same host hierarchy (FASP's setup is deterministic
// and runs redundantly), keeps contiguous row slabs of the fine levels on its GPU and the small
// coarse levels in full:
//   * level l with >= agg_rows global rows: rows [off_l[r], off_l[r+1]) on rank r; the columns of
//     A_l, R_l (fine vector) and P_{l-1} (this level's vector) are renumbered
//     [owned | ghosts by owner rank]; before the matrix kernel the ghosts are packed by the
//     owners and exchanged with one grouped ncclSend/ncclRecv (NVLink 5 / NVSwitch: all peers at
//     full bandwidth, the cost is latency, so there is exactly one exchange per kernel).
//   * below the threshold the levels are replicated: the restricted residual slices are
//     all-gathered once, every rank runs the small sub-hierarchy (incl. the dense coarse solve)
//     redundantly, and the prolongation reads the full coarse vector — no scatter step.
//   * fused dot products / norms are all-reduced in place on the device scalars (reduce_finish).
// The send lists are derived from the global matrix on the host (no setup communication), which
// also makes the partition logic testable without GPUs (tests/test_dist_cpu.py).
#include "dist.cuh"
#include "comm.cuh"
#include "p2p.cuh"
#include <algorithm>

namespace fc {

void reduce_finish(const Reduce& red, const int* gate)
{
    if (!red.global || !comm_active()) return;
    if (red.dot_out) comm_allreduce(red.dot_out, 1, 0, gate);
    if (red.nrm2_out) comm_allreduce(red.nrm2_out, 1, 0, gate);
}

__global__ void k_halo_pack(int n, const int* __restrict__ idx, const double* __restrict__ x,
                            double* __restrict__ buf)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = x[idx[i]];
}

void halo_exchange(const HaloPlan& h, double* x, const int* gate, bool branch)
{
    if (!comm_active()) return;
    if (p2p_halo_exchange(h, x, gate, branch)) return;   // peer-memory push + device barrier
    if (h.nsend == 0 && h.nghost == 0) return;
    ProfScope prof(400, h.nloc, h.nsend + h.nghost, 8.0 * (h.nsend + h.nghost));
    if (h.nsend > 0) {
        int g = (h.nsend + 255) / 256;
        if (g > 1184) g = 1184;
        FC_LAUNCH(k_halo_pack, g, 256, 0, h.nsend, h.send_idx, x, h.send_buf);
    }
    comm_group_start();
    for (size_t p = 0; p < h.send_peer.size(); ++p)
        comm_send(h.send_buf + h.send_off[p], (size_t)h.send_cnt[p], h.send_peer[p]);
    for (size_t p = 0; p < h.recv_peer.size(); ++p)
        comm_recv(x + h.nloc + h.recv_off[p], (size_t)h.recv_cnt[p], h.recv_peer[p]);
    comm_group_end();
}

void halo_free(HaloPlan* h)
{
    if (!h) return;
    dfree(h->send_idx);
    dfree(h->send_buf);
    delete h;
}

std::vector<int> dist_partition(int n, int nranks)
{
    std::vector<int> off(nranks + 1);
    for (int r = 0; r <= nranks; ++r) off[r] = (int)(((long long)n * r) / nranks);
    return off;
}

static int owner_of(const std::vector<int>& off, int c)
{
    int lo = 0, hi = (int)off.size() - 1;   // off[lo] <= c < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) / 2;
        if (off[mid] <= c) lo = mid;
        else hi = mid;
    }
    return lo;
}

void dist_extract(const dCSRmat& A, int r0, int r1, const std::vector<int>& coff, int rank, bool pattern_only,
                  LocalCSR& out, const std::vector<int>& extra)
{
    const int nown = r1 - r0;
    out.rows       = nown + (int)extra.size();
    out.ia.resize((size_t)out.rows + 1);
    auto grow = [&](int i) { return i < nown ? r0 + i : extra[(size_t)(i - nown)]; };   // local row -> global row
    long long nnz = 0;
    for (int i = 0; i < out.rows; ++i) {
        out.ia[i] = (int)nnz;
        const int g = grow(i);
        nnz += A.IA[g + 1] - A.IA[g];
    }
    out.ia[out.rows] = (int)nnz;
    out.ja.resize((size_t)nnz);
    out.val.clear();
    const bool vals = !pattern_only && A.val;
    if (vals) out.val.resize((size_t)nnz);
    for (int i = 0; i < out.rows; ++i) {
        const int g = grow(i), k0 = A.IA[g], len = A.IA[g + 1] - k0;
        std::copy(A.JA + k0, A.JA + k0 + len, out.ja.begin() + out.ia[i]);   // global columns for now
        if (vals) std::copy(A.val + k0, A.val + k0 + len, out.val.begin() + out.ia[i]);
    }
    dist_renumber(out, A.col, coff, rank);
}

// columns of a slab (global numbers in out.ja) -> [owned | ghosts ascending]; fills out.ghosts / out.cols
void dist_renumber(LocalCSR& out, int global_cols, const std::vector<int>& coff, int rank)
{
    out.ghosts.clear();
    if (coff.empty()) {   // replicated column space: global numbering
        out.cols = global_cols;
        return;
    }
    const int c0 = coff[rank], c1 = coff[rank + 1];
    for (int c : out.ja)
        if (c < c0 || c >= c1) out.ghosts.push_back(c);
    std::sort(out.ghosts.begin(), out.ghosts.end());
    out.ghosts.erase(std::unique(out.ghosts.begin(), out.ghosts.end()), out.ghosts.end());
    const int nloc = c1 - c0;
    for (int& c : out.ja) {
        if (c >= c0 && c < c1) c -= c0;
        else c = nloc + (int)(std::lower_bound(out.ghosts.begin(), out.ghosts.end(), c) - out.ghosts.begin());
    }
    out.cols = nloc + (int)out.ghosts.size();
}

void dist_ghost_lists(const dCSRmat& A, const std::vector<int>& off, std::vector<std::vector<int>>& ghosts)
{
    const int nr = (int)off.size() - 1;
    ghosts.assign(nr, std::vector<int>());
    for (int q = 0; q < nr; ++q) {
        std::vector<int>& g = ghosts[q];
        for (int k = A.IA[off[q]]; k < A.IA[off[q + 1]]; ++k) {
            const int c = A.JA[k];
            if (c < off[q] || c >= off[q + 1]) g.push_back(c);
        }
        std::sort(g.begin(), g.end());
        g.erase(std::unique(g.begin(), g.end()), g.end());
    }
}

void dist_send_lists(const dCSRmat& A, const std::vector<int>& roff, const std::vector<int>& coff, int rank,
                     std::vector<std::vector<int>>& send, const std::vector<std::vector<int>>* extra_by_rank)
{
    const int nr = (int)roff.size() - 1;
    send.assign(nr, std::vector<int>());
    const int c0 = coff[rank], c1 = coff[rank + 1];
    std::vector<unsigned char> mark((size_t)(c1 - c0 > 0 ? c1 - c0 : 1));
    for (int q = 0; q < nr; ++q) {
        if (q == rank) continue;
        std::fill(mark.begin(), mark.end(), 0);
        bool any = false;
        auto scan = [&](int k0, int k1) {
            for (int k = k0; k < k1; ++k) {
                const int c = A.JA[k];
                if (c >= c0 && c < c1) mark[c - c0] = 1, any = true;
            }
        };
        scan(A.IA[roff[q]], A.IA[roff[q + 1]]);
        if (extra_by_rank)
            for (int g : (*extra_by_rank)[q]) scan(A.IA[g], A.IA[g + 1]);
        if (!any) continue;
        for (int c = 0; c < c1 - c0; ++c)
            if (mark[c]) send[q].push_back(c);   // ascending local index = ascending global index
    }
}

// plan + upload of one partitioned operator
static HaloPlan* plan_from_send_lists(const std::vector<int>& coff, const LocalCSR& loc, int rank,
                                      const std::vector<std::vector<int>>& send);

static HaloPlan* make_plan(const dCSRmat& A, const std::vector<int>& roff, const std::vector<int>& coff,
                           const LocalCSR& loc, int rank, const std::vector<std::vector<int>>* extra_by_rank)
{
    std::vector<std::vector<int>> send;
    dist_send_lists(A, roff, coff, rank, send, extra_by_rank);
    return plan_from_send_lists(coff, loc, rank, send);
}

static HaloPlan* plan_from_send_lists(const std::vector<int>& coff, const LocalCSR& loc, int rank,
                                      const std::vector<std::vector<int>>& send)
{
    HaloPlan* h = new HaloPlan();
    h->nloc     = coff[rank + 1] - coff[rank];
    h->nghost   = (int)loc.ghosts.size();
    // receive side: ghosts are sorted by global column, i.e. grouped by owner
    for (size_t g = 0; g < loc.ghosts.size();) {
        const int    q = owner_of(coff, loc.ghosts[g]);
        const size_t b = g;
        while (g < loc.ghosts.size() && loc.ghosts[g] < coff[q + 1]) ++g;
        h->recv_peer.push_back(q);
        h->recv_off.push_back((int)b);
        h->recv_cnt.push_back((int)(g - b));
    }
    std::vector<int> idx;
    for (int q = 0; q < (int)send.size(); ++q) {
        if (send[q].empty()) continue;
        h->send_peer.push_back(q);
        h->send_off.push_back((int)idx.size());
        h->send_cnt.push_back((int)send[q].size());
        idx.insert(idx.end(), send[q].begin(), send[q].end());
    }
    h->nsend = (int)idx.size();
    if (h->nsend) {
        h->send_idx = dalloc<int>(idx.size());
        h->send_buf = dalloc<double>(idx.size());
        FC_CUDA(cudaMemcpyAsync(h->send_idx, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice,
                                ctx().stream));
        FC_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    if (p2p_active()) {
        // tell every owner where its entries land behind my owned part: all-gather the receive
        // offsets (collective; every rank builds its plans in the same order)
        const int        nr = comm_size();
        std::vector<int> mine(nr, -1), all;
        for (size_t p = 0; p < h->recv_peer.size(); ++p) mine[h->recv_peer[p]] = h->recv_off[p];
        p2p_allgather_ints(mine, all);
        h->peer_dst_off.clear();
        for (size_t k = 0; k < h->send_peer.size(); ++k) {
            const int q   = h->send_peer[k];
            const int off = all[(size_t)q * nr + rank];
            if (off < 0) fail(ERROR_DATA_STRUCTURE, "halo plans of ranks %d and %d disagree", rank, q);
            h->peer_dst_off.push_back((coff[q + 1] - coff[q]) + off);
        }
        h->p2p_ready = true;
    }
    return h;
}

// extra_by_rank: rows every rank computes in addition to its slab (the ghost rows of the level the operator
// writes to); the local operator gets them appended behind the owned rows.
static HaloPlan* upload_part(DevCSR& d, const dCSRmat& A, const std::vector<int>& roff,
                             const std::vector<int>& coff, int rank, bool pattern,
                             const std::vector<std::vector<int>>* extra_by_rank = nullptr)
{
    LocalCSR loc;
    static const std::vector<int> none;
    dist_extract(A, roff[rank], roff[rank + 1], coff, rank, pattern, loc, extra_by_rank ? (*extra_by_rank)[rank] : none);
    const int nloc_cols = coff.empty() ? -1 : coff[rank + 1] - coff[rank];   // columns behind it are ghosts
    csr_upload(d, loc.rows, loc.cols, (long long)loc.ja.size(), loc.ia.data(), loc.ja.data(),
               loc.val.empty() ? nullptr : loc.val.data(), pattern, nloc_cols);
    if (coff.empty()) return nullptr;
    HaloPlan* h = make_plan(A, roff, coff, loc, rank, extra_by_rank);
    d.halo      = h;
    d.nghost    = h->nghost;
    return h;
}

// vectors of one level: identical capacity on every rank, peer mapping, the smoother divisor on the ghost rows
static void finish_level(Amg& h, Level& L, int l, int lrep, bool redundant)
{
    const int nl = h.nl;
    if (comm_active() && (l <= lrep)) {
        // identical vector capacity on every rank (peer addresses = same offsets)
        double  c  = (double)(L.cap > L.n ? L.cap : L.n);
        double* dc = dalloc<double>(1);
        FC_CUDA(cudaMemcpyAsync(dc, &c, 8, cudaMemcpyHostToDevice, ctx().stream));
        comm_allreduce(dc, 1, 2);
        FC_CUDA(cudaMemcpyAsync(&c, dc, 8, cudaMemcpyDeviceToHost, ctx().stream));
        FC_CUDA(cudaStreamSynchronize(ctx().stream));
        dfree(dc);
        if ((int)c > L.cap && h.smoother == SMOOTHER_POLY)
            for (int i = 0; i < 3; ++i) {   // allocated with this rank's own ghost count: too short for the peers' pushes
                dfree(L.pv[i]);
                L.pv[i] = nullptr;
            }
        L.cap = (int)c;
        if (l == 0) L.A.vec_cap = L.cap;
        if (h.smoother == SMOOTHER_POLY && l < nl - 1) amg_level_poly_vectors(h, L);
    }
    amg_level_vectors(h, L);
    if (l <= lrep) {   // partitioned levels + the first replicated one (all-gather target)
        const size_t bytes = sizeof(double) * ((size_t)L.cap + 8);
        p2p_register(L.b, bytes);
        p2p_register(L.xa, bytes);
        p2p_register(L.xb, bytes);
        p2p_register(L.w, bytes);
        for (int i = 0; i < 3; ++i)
            if (L.pv[i]) p2p_register(L.pv[i], bytes);
        L.p2p_registered = true;
        // the zero-guess sweep x = b / d on the ghost rows needs d there: exchanged once, here
        if (L.dist && redundant && (h.smoother == SMOOTHER_JACOBI || h.smoother == SMOOTHER_L1DIAG) && l < nl - 1) {
            const double* src = (h.smoother == SMOOTHER_JACOBI) ? L.A.diag : L.A.l1;
            L.dscale_ext      = dalloc<double>((size_t)L.cap + 8);
            FC_CUDA(cudaMemsetAsync(L.dscale_ext, 0, sizeof(double) * ((size_t)L.cap + 8), ctx().stream));
            FC_CUDA(cudaMemcpyAsync(L.dscale_ext, src, sizeof(double) * (size_t)L.n, cudaMemcpyDeviceToDevice,
                                    ctx().stream));
            p2p_register(L.dscale_ext, bytes);
            halo_exchange(*L.hA, L.dscale_ext);
            FC_CUDA(cudaStreamSynchronize(ctx().stream));
            h.bytes += bytes;
        }
    }
}

Amg* dist_amg_upload(AMG_data* mgl, AMG_param* param, int agg_rows)
{
    ensure_init();
    if (!comm_active()) return amg_upload(mgl, param);
    const int rank = comm_rank(), nr = comm_size();
    const int nl = mgl[0].num_levels;
    if (nl < 1 || nl > MAX_AMG_LVL) fail(ERROR_DATA_STRUCTURE, "dist_amg_upload: num_levels = %d", nl);
    if (param->smoother != SMOOTHER_JACOBI && param->smoother != SMOOTHER_L1DIAG && param->smoother != SMOOTHER_POLY)
        fail(ERROR_AMG_SMOOTH_TYPE, "multi-GPU cycle: smoother %d not supported (Jacobi 1, poly 9, L1 10)",
             (int)param->smoother);
    if (param->cycle_type == AMLI_CYCLE || param->cycle_type == NL_AMLI_CYCLE)
        fail(ERROR_INPUT_PAR, "AMLI cycles are not on the device path");

    // partitioned levels: 0 .. lrep-1 ; the coarsest level is always replicated (dense solve)
    int lrep = 0;
    while (lrep < nl - 1 && mgl[lrep].A.row >= agg_rows && mgl[lrep].A.row >= 4 * nr) ++lrep;
    if (lrep == 0) {
        if (nl < 2 || mgl[0].A.row < 4 * nr)
            fail(ERROR_INPUT_PAR, "multi-GPU solve needs at least two levels and 4 rows per rank");
        lrep = 1;   // the finest level is always partitioned
    }
    std::vector<std::vector<int>> off(nl);
    for (int l = 0; l < nl; ++l) off[l] = dist_partition(mgl[l].A.row, nr);
    const std::vector<int> none;

    Amg* h = new Amg();
    try {
        amg_set_params(*h, param);
        h->nl   = nl;
        h->dist = lrep > 0;
        h->off0 = off[0];
        h->lv.resize(nl);
        const bool ua = (param->AMG_type == UA_AMG);
        // Redundant ghost rows: a ghost exchange is a synchronisation of all ranks (push + flag barrier, ~15 us), and
        // a level of the cycle needs four of them (x before the residual, w before R, x_{l+1} before P, x before the
        // post-smoother). Two go away when a rank ALSO computes, with the same kernels and therefore the same bits,
        // the few rows its neighbours own but its A_l gathers: P_l gets the ghost rows of x_l appended (the
        // post-smoother then finds its ghosts up to date), R_l the ghost rows of b_{l+1} (the zero-guess pre-smoothing
        // sweep x = b / d and the residual then need none). The price: slightly longer exchanges for w_l and x_{l+1}.
        const bool redundant = ctx().opt.ghost_redundant != 0;
        std::vector<std::vector<std::vector<int>>> GA(lrep);   // [level][rank] ghost columns of A_l's slab
        if (redundant)
            for (int l = 0; l < lrep; ++l) dist_ghost_lists(mgl[l].A, off[l], GA[l]);
        for (int l = 0; l < nl; ++l) {
            Level&         L = h->lv[l];
            const dCSRmat& A = mgl[l].A;
            L.nglobal        = A.row;
            if (l < lrep) {
                L.dist = true;
                L.row0 = off[l][rank];
                L.n    = off[l][rank + 1] - off[l][rank];
                L.hA   = upload_part(L.A, A, off[l], off[l], rank, false);
                int ghost = L.A.nghost;
                const bool next_dist = (l + 1 < lrep);
                // R_l: rows = my slice of level l+1 (+ the ghost rows of b_{l+1}), gathers the level-l residual
                L.r_ext = redundant && next_dist;
                L.hR    = upload_part(L.R, mgl[l].R, off[l + 1], off[l], rank, ua, L.r_ext ? &GA[l + 1] : nullptr);
                if (L.R.nghost > ghost) ghost = L.R.nghost;
                // P_l: my fine rows (+ the ghost rows of x_l), gathers x_{l+1} (partitioned or replicated)
                L.p_ext = redundant;
                L.hP    = upload_part(L.P, mgl[l].P, off[l], next_dist ? off[l + 1] : none, rank, ua,
                                      L.p_ext ? &GA[l] : nullptr);
                if (l > 0 && h->lv[l - 1].P.nghost > ghost) ghost = h->lv[l - 1].P.nghost;
                L.cap = L.n + ghost;
                if (!next_dist) {
                    L.gcounts.resize(nr);
                    L.gdispls.resize(nr);
                    for (int r = 0; r < nr; ++r) {
                        L.gdispls[r] = (size_t)off[l + 1][r];
                        L.gcounts[r] = (size_t)(off[l + 1][r + 1] - off[l + 1][r]);
                    }
                }
                // smoother data on the local slab (the diagonal of local row i is local column i)
                dCSRmat  hostA = A;
                amg_level_smoother_data(*h, L, &hostA);
            } else {
                csr_upload(L.A, A.row, A.col, A.nnz, A.IA, A.JA, A.val);
                L.n = A.row;
                if (l < nl - 1) {
                    const dCSRmat& P = mgl[l].P;
                    const dCSRmat& R = mgl[l].R;
                    csr_upload(L.P, P.row, P.col, P.nnz, P.IA, P.JA, P.val, ua);
                    csr_upload(L.R, R.row, R.col, R.nnz, R.IA, R.JA, R.val, ua);
                    amg_level_smoother_data(*h, L, &A);
                }
            }
            finish_level(*h, L, l, lrep, redundant);
            h->bytes += L.A.bytes + L.P.bytes + L.R.bytes;
        }
        h->scal = dalloc<double>(4);
        FC_CUDA(cudaMemsetAsync(h->scal, 0, 4 * sizeof(double), ctx().stream));
        amg_setup_coarse(*h);   // replicated coarsest level: every rank factors / iterates redundantly
        FC_CUDA(cudaStreamSynchronize(ctx().stream));
    } catch (...) {
        amg_free(h);
        throw;
    }
    return h;
}

// ------------------------------------------------------------------------------------
// The same hierarchy from per-rank slabs (no global matrix on any rank; include/fasp_cuda.h
// fasp_cuda_dist_krylov_amg_create_slabs, host side: faspsolver_b200/slabsetup.py)
// ------------------------------------------------------------------------------------
// every rank's list, variable length (collective)
static void allgather_lists(const std::vector<int>& mine, std::vector<std::vector<int>>& all)
{
    const int        nr = comm_size();
    std::vector<int> cnt(1, (int)mine.size()), cnts;
    p2p_allgather_ints(cnt, cnts);
    int mx = 1;
    for (int c : cnts) mx = std::max(mx, c);
    std::vector<int> pad(mine), flat;
    pad.resize((size_t)mx, 0);
    p2p_allgather_ints(pad, flat);
    all.assign(nr, std::vector<int>());
    for (int q = 0; q < nr; ++q) all[q].assign(flat.begin() + (size_t)q * mx, flat.begin() + (size_t)q * mx + cnts[q]);
}

// M: this rank's rows (owned + extra rows behind them) with GLOBAL column numbers
static HaloPlan* upload_slab(DevCSR& d, const dCSRmat& M, const std::vector<int>& coff, int rank, bool pattern)
{
    LocalCSR loc;
    loc.rows = M.row;
    loc.ia.assign(M.IA, M.IA + M.row + 1);
    if (loc.ia[0] != 0) fail(ERROR_DATA_STRUCTURE, "slab operator: IA[0] != 0");
    const size_t nnz = (size_t)loc.ia[M.row];
    loc.ja.assign(M.JA, M.JA + nnz);
    if (!pattern && M.val) loc.val.assign(M.val, M.val + nnz);
    for (int c : loc.ja)
        if (c < 0 || c >= M.col) fail(ERROR_DATA_STRUCTURE, "slab operator: column %d outside [0, %d)", c, M.col);
    dist_renumber(loc, M.col, coff, rank);
    const int nloc_cols = coff.empty() ? -1 : coff[rank + 1] - coff[rank];
    csr_upload(d, loc.rows, loc.cols, (long long)nnz, loc.ia.data(), loc.ja.data(),
               loc.val.empty() ? nullptr : loc.val.data(), pattern, nloc_cols);
    if (coff.empty()) return nullptr;
    // what my peers gather from my owned columns: their ghost lists, all-gathered (collective)
    std::vector<std::vector<int>> ghosts_of, send(comm_size());
    allgather_lists(loc.ghosts, ghosts_of);
    const int c0 = coff[rank], c1 = coff[rank + 1];
    for (int q = 0; q < comm_size(); ++q) {
        if (q == rank) continue;
        const std::vector<int>& g = ghosts_of[q];
        auto b = std::lower_bound(g.begin(), g.end(), c0), e = std::lower_bound(g.begin(), g.end(), c1);
        for (auto it = b; it != e; ++it) send[q].push_back(*it - c0);
    }
    HaloPlan* h = plan_from_send_lists(coff, loc, rank, send);
    d.halo      = h;
    d.nghost    = h->nghost;
    return h;
}

Amg* dist_amg_upload_slabs(int nlev, const fasp_cuda_slab_level* sl, const int* tail_off, AMG_data* tail,
                           AMG_param* param)
{
    ensure_init();
    if (!comm_active()) fail(ERROR_INPUT_PAR, "slab hierarchy: no communicator (fasp_cuda_comm_init first)");
    const int rank = comm_rank(), nr = comm_size();
    if (nlev < 1 || !sl || !tail || !tail_off) fail(ERROR_INPUT_PAR, "slab hierarchy: nlev = %d / null argument", nlev);
    const int nt = tail[0].num_levels;
    const int nl = nlev + nt;
    if (nt < 1 || nl > MAX_AMG_LVL) fail(ERROR_DATA_STRUCTURE, "slab hierarchy: %d + %d levels", nlev, nt);
    if (param->smoother != SMOOTHER_JACOBI && param->smoother != SMOOTHER_L1DIAG && param->smoother != SMOOTHER_POLY)
        fail(ERROR_AMG_SMOOTH_TYPE, "multi-GPU cycle: smoother %d not supported (Jacobi 1, poly 9, L1 10)",
             (int)param->smoother);
    if (param->cycle_type == AMLI_CYCLE || param->cycle_type == NL_AMLI_CYCLE)
        fail(ERROR_INPUT_PAR, "AMLI cycles are not on the device path");
    const int lrep = nlev;
    std::vector<std::vector<int>> off(nlev + 1);
    for (int l = 0; l < nlev; ++l) off[l].assign(sl[l].row_off, sl[l].row_off + nr + 1);
    off[nlev].assign(tail_off, tail_off + nr + 1);
    for (int l = 0; l <= nlev; ++l) {
        const int n_glob = (l < nlev) ? sl[l].A.col : tail[0].A.row;
        if (off[l][0] != 0 || off[l][nr] != n_glob) fail(ERROR_DATA_STRUCTURE, "slab hierarchy: partition of level %d", l);
        for (int r = 0; r < nr; ++r)
            if (off[l][r + 1] < off[l][r]) fail(ERROR_DATA_STRUCTURE, "slab hierarchy: partition of level %d", l);
    }
    const std::vector<int> none;
    Amg* h = new Amg();
    try {
        amg_set_params(*h, param);
        h->nl   = nl;
        h->dist = true;
        h->off0 = off[0];
        h->lv.resize(nl);
        const bool ua        = (param->AMG_type == UA_AMG);
        const bool redundant = ctx().opt.ghost_redundant != 0;
        for (int l = 0; l < nl; ++l) {
            Level& L = h->lv[l];
            if (l < lrep) {
                const fasp_cuda_slab_level& S = sl[l];
                const int n_own  = off[l][rank + 1] - off[l][rank];
                const int nc_own = off[l + 1][rank + 1] - off[l + 1][rank];
                const bool next_dist = (l + 1 < lrep);
                if (S.A.row != n_own) fail(ERROR_DATA_STRUCTURE, "slab level %d: A has %d rows, the partition says %d", l, S.A.row, n_own);
                L.dist    = true;
                L.nglobal = S.A.col;
                L.row0    = off[l][rank];
                L.n       = n_own;
                L.hA      = upload_slab(L.A, S.A, off[l], rank, false);
                int ghost = L.A.nghost;
                // the extra rows must be exactly the ghost rows the cycle expects (same count on this rank); they are
                // used only when the option is on AND the host supplied them
                L.r_ext = redundant && next_dist && S.n_rext > 0;
                L.p_ext = redundant && S.n_pext > 0;
                if (S.P.row != n_own + S.n_pext || S.R.row != nc_own + S.n_rext)
                    fail(ERROR_DATA_STRUCTURE, "slab level %d: P / R row counts do not match owned + extra rows", l);
                if (S.n_pext != 0 && S.n_pext != L.A.nghost)
                    fail(ERROR_DATA_STRUCTURE, "slab level %d: %d extra rows of P, A has %d ghost columns", l, S.n_pext, L.A.nghost);
                dCSRmat Rm = S.R, Pm = S.P;
                if (!L.r_ext) Rm.row = nc_own, Rm.nnz = Rm.IA[nc_own];
                if (!L.p_ext) Pm.row = n_own, Pm.nnz = Pm.IA[n_own];
                L.hR = upload_slab(L.R, Rm, off[l], rank, ua);
                if (L.R.nghost > ghost) ghost = L.R.nghost;
                L.hP = upload_slab(L.P, Pm, next_dist ? off[l + 1] : none, rank, ua);
                if (l > 0 && h->lv[l - 1].P.nghost > ghost) ghost = h->lv[l - 1].P.nghost;
                L.cap = L.n + ghost;
                if (!next_dist) {
                    L.gcounts.resize(nr);
                    L.gdispls.resize(nr);
                    for (int r = 0; r < nr; ++r) {
                        L.gdispls[r] = (size_t)off[l + 1][r];
                        L.gcounts[r] = (size_t)(off[l + 1][r + 1] - off[l + 1][r]);
                    }
                }
                amg_level_smoother_data(*h, L, nullptr);
            } else {
                const AMG_data& T = tail[l - lrep];
                const dCSRmat&  A = T.A;
                L.nglobal         = A.row;
                csr_upload(L.A, A.row, A.col, A.nnz, A.IA, A.JA, A.val);
                L.n = A.row;
                if (l < nl - 1) {
                    csr_upload(L.P, T.P.row, T.P.col, T.P.nnz, T.P.IA, T.P.JA, T.P.val, ua);
                    csr_upload(L.R, T.R.row, T.R.col, T.R.nnz, T.R.IA, T.R.JA, T.R.val, ua);
                    amg_level_smoother_data(*h, L, &A);
                }
            }
            finish_level(*h, L, l, lrep, redundant);
            h->bytes += L.A.bytes + L.P.bytes + L.R.bytes;
        }
        // a slab level whose R carries the ghost rows of the next level: their count must be that level's ghost count
        for (int l = 0; l + 1 < lrep; ++l)
            if (h->lv[l].r_ext && sl[l].n_rext != h->lv[l + 1].A.nghost)
                fail(ERROR_DATA_STRUCTURE, "slab level %d: %d extra rows of R, A_%d has %d ghost columns", l, sl[l].n_rext,
                     l + 1, h->lv[l + 1].A.nghost);
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

} // namespace fc
