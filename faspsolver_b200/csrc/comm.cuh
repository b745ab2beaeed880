// comm.cuh — process-global NCCL communicator (one process per GPU) for the row-partitioned
// multi-GPU solve: halo exchange (grouped send/recv), Krylov scalars (all-reduce), coarse-level
// agglomeration (all-gather). NCCL is bound at run time with dlopen("libnccl.so.2"), so the
// library loads on machines without NCCL and single-GPU runs never touch it.
#pragma once
#include "common.cuh"

namespace fc {

int  comm_rank();
int  comm_size();
bool comm_active();   // size > 1

void comm_unique_id(void* id128);
void comm_init(const void* id128, int rank, int nranks);
void comm_finalize();

// in-place all-reduce of `count` doubles on the library stream (op: 0 sum, 2 max). `gate`: see p2p.cuh (the
// peer-memory path skips the collective on every rank when *gate != 0; the NCCL path cannot and runs it)
void comm_allreduce(double* buf, size_t count, int op = 0, const int* gate = nullptr);
// up to 4 doubles, slot s a max-reduction when bit s of maxmask is set, a sum otherwise (one collective)
void comm_allreduce_mixed(double* buf, int count, int maxmask, const int* gate = nullptr);
void comm_allgatherv(const double* send, size_t sendcount, double* recv, const std::vector<size_t>& counts,
                     const std::vector<size_t>& displs, const int* gate = nullptr);
void comm_group_start();
void comm_group_end();
void comm_send(const double* buf, size_t count, int peer);
void comm_recv(double* buf, size_t count, int peer);

} // namespace fc
