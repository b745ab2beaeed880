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
                  LocalCSR& out)
{
    out.rows = r1 - r0;
    out.ia.resize((size_t)out.rows + 1);
    const int k0 = A.IA[r0], k1 = A.IA[r1];
    out.ja.resize((size_t)(k1 - k0));
    out.val.clear();
    if (!pattern_only && A.val) out.val.assign(A.val + k0, A.val + k1);
    for (int i = r0; i <= r1; ++i) out.ia[i - r0] = A.IA[i] - k0;
    out.ghosts.clear();
    if (coff.empty()) {   // replicated column space: global numbering
        for (int k = k0; k < k1; ++k) out.ja[k - k0] = A.JA[k];
        out.cols = A.col;
        return;
    }
    const int c0 = coff[rank], c1 = coff[rank + 1];
    for (int k = k0; k < k1; ++k) {
        const int c = A.JA[k];
        if (c < c0 || c >= c1) out.ghosts.push_back(c);
    }
    std::sort(out.ghosts.begin(), out.ghosts.end());
    out.ghosts.erase(std::unique(out.ghosts.begin(), out.ghosts.end()), out.ghosts.end());
    const int nloc = c1 - c0;
    for (int k = k0; k < k1; ++k) {
        const int c = A.JA[k];
        if (c >= c0 && c < c1) out.ja[k - k0] = c - c0;
        else
            out.ja[k - k0] =
                nloc + (int)(std::lower_bound(out.ghosts.begin(), out.ghosts.end(), c) - out.ghosts.begin());
    }
    out.cols = nloc + (int)out.ghosts.size();
}

void dist_send_lists(const dCSRmat& A, const std::vector<int>& roff, const std::vector<int>& coff, int rank,
                     std::vector<std::vector<int>>& send)
{
    const int nr = (int)roff.size() - 1;
    send.assign(nr, std::vector<int>());
    const int c0 = coff[rank], c1 = coff[rank + 1];
    std::vector<unsigned char> mark((size_t)(c1 - c0 > 0 ? c1 - c0 : 1));
    for (int q = 0; q < nr; ++q) {
        if (q == rank) continue;
        std::fill(mark.begin(), mark.end(), 0);
        bool any = false;
        for (int k = A.IA[roff[q]]; k < A.IA[roff[q + 1]]; ++k) {
            const int c = A.JA[k];
            if (c >= c0 && c < c1) mark[c - c0] = 1, any = true;
        }
        if (!any) continue;
        for (int c = 0; c < c1 - c0; ++c)
            if (mark[c]) send[q].push_back(c);   // ascending local index = ascending global index
    }
}

// plan + upload of one partitioned operator
static HaloPlan* make_plan(const dCSRmat& A, const std::vector<int>& roff, const std::vector<int>& coff,
                           const LocalCSR& loc, int rank)
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
    std::vector<std::vector<int>> send;
    dist_send_lists(A, roff, coff, rank, send);
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

static HaloPlan* upload_part(DevCSR& d, const dCSRmat& A, const std::vector<int>& roff,
                             const std::vector<int>& coff, int rank, bool pattern)
{
    LocalCSR loc;
    dist_extract(A, roff[rank], roff[rank + 1], coff, rank, pattern, loc);
    const int nloc_cols = coff.empty() ? -1 : coff[rank + 1] - coff[rank];   // columns behind it are ghosts
    csr_upload(d, loc.rows, loc.cols, (long long)loc.ja.size(), loc.ia.data(), loc.ja.data(),
               loc.val.empty() ? nullptr : loc.val.data(), pattern, nloc_cols);
    if (coff.empty()) return nullptr;
    HaloPlan* h = make_plan(A, roff, coff, loc, rank);
    d.halo      = h;
    d.nghost    = h->nghost;
    return h;
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
                // R_l: rows = my slice of level l+1, gathers the level-l residual
                L.hR = upload_part(L.R, mgl[l].R, off[l + 1], off[l], rank, ua);
                if (L.R.nghost > ghost) ghost = L.R.nghost;
                // P_l: my fine rows, gathers x_{l+1} (partitioned or replicated)
                const bool next_dist = (l + 1 < lrep);
                L.hP = upload_part(L.P, mgl[l].P, off[l], next_dist ? off[l + 1] : none, rank, ua);
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
            if (comm_active() && (l <= lrep)) {
                // identical vector capacity on every rank (peer addresses = same offsets)
                double  c  = (double)(L.cap > L.n ? L.cap : L.n);
                double* dc = dalloc<double>(1);
                FC_CUDA(cudaMemcpyAsync(dc, &c, 8, cudaMemcpyHostToDevice, ctx().stream));
                comm_allreduce(dc, 1, 2);
                FC_CUDA(cudaMemcpyAsync(&c, dc, 8, cudaMemcpyDeviceToHost, ctx().stream));
                FC_CUDA(cudaStreamSynchronize(ctx().stream));
                dfree(dc);
                L.cap = (int)c;
                if (l == 0) L.A.vec_cap = L.cap;
            }
            amg_level_vectors(*h, L);
            if (l <= lrep) {   // partitioned levels + the first replicated one (all-gather target)
                const size_t bytes = sizeof(double) * ((size_t)L.cap + 8);
                p2p_register(L.b, bytes);
                p2p_register(L.xa, bytes);
                p2p_register(L.xb, bytes);
                p2p_register(L.w, bytes);
                for (int i = 0; i < 3; ++i)
                    if (L.pv[i]) p2p_register(L.pv[i], bytes);
                L.p2p_registered = true;
            }
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

} // namespace fc
