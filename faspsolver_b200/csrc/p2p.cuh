// p2p.cuh — peer-memory data path of the multi-GPU solve (one process per GPU, NVLink 5 / NVSwitch).
//
// Instead of NCCL send/recv (≈25 µs per small exchange inside a CUDA graph), the owners PUSH their
// boundary entries straight into the ghost region of the peers' vectors with plain stores over
// NVLink, then all ranks meet in a device-side barrier built on flags in peer-mapped memory
// (≈5 µs per exchange). Dot products are combined the same way (every rank adds the ranks'
// partials in rank order: deterministic and identical everywhere). Memory is shared between the
// processes with CUDA IPC; NCCL is used only once, at setup, to exchange the IPC handles.
#pragma once
#include "common.cuh"

namespace fc {

constexpr int P2P_MAX_RANKS = 16;

bool p2p_active();
// Called once after comm_init (collective): control blocks, peer mapping. Falls back to NCCL
// (returns false) when IPC / peer access is unavailable.
bool p2p_init();
void p2p_finalize();

// Collective: make a cudaMalloc'ed allocation readable/writable by all ranks. Every rank must call
// it in the same order with the same size.
void p2p_register(void* base, size_t bytes);
void p2p_unregister(void* base);
// peer addresses of a pointer inside a registered allocation (false if not registered)
bool p2p_lookup(const void* p, double* (&peer)[P2P_MAX_RANKS]);

struct HaloPlan;
// collective setup helper: all-gather `n` ints per rank (host vectors)
void p2p_allgather_ints(const std::vector<int>& mine, std::vector<int>& all);
void p2p_reset_order();   // forget the last exchanged vector (start of a captured graph)
int  p2p_error();         // nonzero if a device-side wait ran out of its spin budget
// push-based ghost exchange + barrier; returns false if x is not peer-mapped (caller uses NCCL)
// `gate` (device flag, identical on every rank): the kernels of the exchange return at once when *gate != 0,
// on all ranks alike, so a skipped exchange costs launches but no NVLink round trip and the barrier epochs stay
// in step. `branch` marks a gate that is a rarely-taken branch flag rather than the solver's `done` flag: the
// exchange may or may not run, so the next exchange must not rely on it having separated two pushes.
bool p2p_halo_exchange(const HaloPlan& h, double* x, const int* gate = nullptr, bool branch = false);
// in-place all-reduce of <= 4 doubles; bit s of maxmask makes slot s a max-reduction (others are sums)
void p2p_allreduce(double* buf, int count, int maxmask, const int* gate = nullptr);
// every rank's slice [displs[r], +counts[r]) of `full` is filled from its owner
bool p2p_allgatherv(double* full, const std::vector<size_t>& counts, const std::vector<size_t>& displs,
                    const int* gate = nullptr);
void p2p_barrier(const int* gate = nullptr);

} // namespace fc
