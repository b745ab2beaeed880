// dist.cuh — row-partitioned multi-GPU hierarchy: one process per GPU, contiguous row slabs per
// level, ghost ("halo") exchange before every matrix kernel, replicated coarse levels.
#pragma once
#include "common.cuh"
#include "amg.cuh"

namespace fc {

// Ghost exchange plan of one local operator. The gathered vector x is laid out
// [owned entries | ghosts grouped by owner rank, ascending global index].
struct HaloPlan {
    int nloc = 0, nghost = 0;
    std::vector<int> send_peer, send_cnt, send_off;   // what the peers need from my owned part
    std::vector<int> recv_peer, recv_cnt, recv_off;   // where their entries land behind nloc
    int*    send_idx = nullptr;   // device: local indices to pack, all peers concatenated
    double* send_buf = nullptr;   // device: packed entries
    int     nsend    = 0;
    // peer-memory path (p2p.cu): where each send segment lands inside the peer's vector
    bool             p2p_ready = false;
    std::vector<int> peer_dst_off;
};
void halo_free(HaloPlan* h);

// Contiguous row partition of n rows over nranks (proportional to a parent partition when given)
std::vector<int> dist_partition(int n, int nranks);

// Host-side extraction of the local slab of a global CSR operator: rows [r0, r1), columns
// renumbered against the column partition `coff` (size nranks+1) of the gathered vector, or kept
// global when `coff` is empty (replicated column space). Fills the plan's host lists.
struct LocalCSR {
    std::vector<int>    ia, ja;
    std::vector<double> val;
    std::vector<int>    ghosts;      // global column of every ghost, ascending
    int                 rows = 0, cols = 0;
};
// `extra`: further global rows appended behind the contiguous block [r0, r1) (the ghost rows a rank computes
// redundantly, see dist.cu); may be empty.
void dist_extract(const dCSRmat& A, int r0, int r1, const std::vector<int>& coff, int rank, bool pattern_only,
                  LocalCSR& out, const std::vector<int>& extra = std::vector<int>());
// Send lists: which of my owned columns [coff[rank], coff[rank+1]) the rows of every other rank
// reference (computed from the global matrix every rank holds on the host; no communication). `extra_by_rank`
// (optional): the extra rows of every rank.
void dist_send_lists(const dCSRmat& A, const std::vector<int>& roff, const std::vector<int>& coff, int rank,
                     std::vector<std::vector<int>>& send,
                     const std::vector<std::vector<int>>* extra_by_rank = nullptr);
// ghost columns (global, ascending) of every rank's row slab of a square level operator
void dist_ghost_lists(const dCSRmat& A, const std::vector<int>& off, std::vector<std::vector<int>>& ghosts);

void dist_renumber(LocalCSR& out, int global_cols, const std::vector<int>& coff, int rank);

// Upload: levels with >= agg_rows global rows are partitioned, the rest replicated.
Amg* dist_amg_upload(AMG_data* mgl, AMG_param* param, int agg_rows);
// The same from per-rank slabs with global column numbers (nlev partitioned levels) + the replicated rest
Amg* dist_amg_upload_slabs(int nlev, const fasp_cuda_slab_level* sl, const int* tail_off, AMG_data* tail,
                           AMG_param* param);

} // namespace fc
