// p2p.cu — see p2p.cuh. Kernels that talk to peer GPUs directly: k_halo_push (stores into the
// peers' ghost regions), k_p2p_barrier / k_p2p_allreduce (flags and slots in peer-mapped control
// blocks). Every spin loop has a cycle budget: on expiry the kernel raises an error word instead of
// hanging the GPU.
#include "p2p.cuh"
#include "comm.cuh"
#include "dist.cuh"
#include <map>

namespace fc {

struct P2PCtrl {
    unsigned long long flags[P2P_MAX_RANKS];         // flags[q] = last epoch rank q announced to me
    double             slots[2][P2P_MAX_RANKS][4];   // all-reduce partials, double-buffered by epoch parity
    unsigned long long epoch;                        // my barrier counter (device-resident: graph replays advance it)
    int                error;                        // spin budget exceeded
    // fused ghost exchange (k_halo_push_signal + consumer-side wait): point-to-point, neighbours only
    unsigned long long dflags[P2P_MAX_RANKS];        // dflags[q] = last exchange number whose ghosts rank q stored here
    unsigned long long acks[P2P_MAX_RANKS];          // acks[q]   = rank q has consumed all exchanges <= this number
    unsigned long long seq;                          // my exchange counter
    unsigned int       ticket;                       // last-CTA detection of the push kernel
};

namespace {
struct Region {
    char*  base;
    size_t bytes;
    char*  peer[P2P_MAX_RANKS];
};
struct State {
    bool     on = false;
    int      rank = 0, size = 1;
    P2PCtrl* ctrl = nullptr;
    P2PCtrl* peer_ctrl[P2P_MAX_RANKS] = {nullptr};
    std::vector<Region> regions;
    const void* last_exchanged = nullptr;   // a vector must not be pushed into twice in a row
    // fused protocol: what the previous exchange looked like (decides which peers must acknowledge)
    const void*  fused_prev_x    = nullptr;
    unsigned int fused_prev_recv = 0;
    bool         fused_synced    = false;   // an all-to-all barrier ran since the previous exchange
    unsigned int prev_partners   = 0;       // neighbour-barrier mode: ranks the previous exchange synchronised with
};
State& S()
{
    static State s;
    return s;
}
struct PeerCtrls {
    P2PCtrl* p[P2P_MAX_RANKS];
};
constexpr long long SPIN_BUDGET = 4000000000LL;   // ~2 s of SM clocks
} // namespace

bool p2p_active() { return S().on; }
void p2p_reset_order_hook()
{
    State& s          = S();
    s.last_exchanged  = nullptr;
    s.fused_prev_x    = nullptr;
    s.fused_prev_recv = 0;
    s.fused_synced    = false;
    s.prev_partners   = 0;
}

// all-gather of a fixed-size byte blob per rank through NCCL (setup only)
static void allgather_bytes(const void* mine, size_t bytes, std::vector<char>& all)
{
    const int    nr   = comm_size();
    const size_t nd   = (bytes + 7) / 8;   // in doubles
    double*      dbuf = dalloc<double>(nd * nr);
    std::vector<double> tmp(nd, 0.0);
    memcpy(tmp.data(), mine, bytes);
    FC_CUDA(cudaMemcpyAsync(dbuf + nd * comm_rank(), tmp.data(), nd * 8, cudaMemcpyHostToDevice, ctx().stream));
    std::vector<size_t> counts(nr, nd), displs(nr);
    for (int r = 0; r < nr; ++r) displs[r] = nd * r;
    comm_allgatherv(dbuf + nd * comm_rank(), nd, dbuf, counts, displs);
    std::vector<double> host(nd * nr);
    FC_CUDA(cudaMemcpyAsync(host.data(), dbuf, nd * nr * 8, cudaMemcpyDeviceToHost, ctx().stream));
    FC_CUDA(cudaStreamSynchronize(ctx().stream));
    dfree(dbuf);
    all.resize(bytes * nr);
    for (int r = 0; r < nr; ++r) memcpy(all.data() + bytes * r, host.data() + nd * r, bytes);
}

static bool map_region(Region& R)
{
    State& s = S();
    cudaIpcMemHandle_t hnd;
    if (cudaIpcGetMemHandle(&hnd, R.base) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    std::vector<char> all;
    allgather_bytes(&hnd, sizeof(hnd), all);
    for (int q = 0; q < s.size; ++q) {
        if (q == s.rank) {
            R.peer[q] = R.base;
            continue;
        }
        cudaIpcMemHandle_t hq;
        memcpy(&hq, all.data() + sizeof(hq) * q, sizeof(hq));
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        R.peer[q] = static_cast<char*>(p);
    }
    return true;
}

void p2p_allgather_ints(const std::vector<int>& mine, std::vector<int>& all)
{
    std::vector<char> bytes;
    allgather_bytes(mine.data(), mine.size() * sizeof(int), bytes);
    all.resize(mine.size() * comm_size());
    memcpy(all.data(), bytes.data(), bytes.size());
}
void p2p_reset_order() { p2p_reset_order_hook(); }
int  p2p_error()
{
    State& s = S();
    if (!s.on) return 0;
    int e = 0;
    cudaMemcpy(&e, &s.ctrl->error, sizeof(int), cudaMemcpyDeviceToHost);
    return e;
}

bool p2p_init()
{
    State& s = S();
    if (s.on || !comm_active()) return s.on;
    if (getenv("FASP_CUDA_P2P") && atoi(getenv("FASP_CUDA_P2P")) == 0) return false;
    s.rank = comm_rank();
    s.size = comm_size();
    if (s.size > P2P_MAX_RANKS) return false;
    s.ctrl = static_cast<P2PCtrl*>(dmalloc(sizeof(P2PCtrl)));
    FC_CUDA(cudaMemset(s.ctrl, 0, sizeof(P2PCtrl)));
    Region R{reinterpret_cast<char*>(s.ctrl), sizeof(P2PCtrl), {nullptr}};
    // all ranks must agree on success: reduce a flag
    int     ok   = map_region(R) ? 1 : 0;
    double* flag = dalloc<double>(1);
    double  v    = ok ? 0.0 : 1.0;
    FC_CUDA(cudaMemcpy(flag, &v, 8, cudaMemcpyHostToDevice));
    comm_allreduce(flag, 1, 0);
    FC_CUDA(cudaMemcpyAsync(&v, flag, 8, cudaMemcpyDeviceToHost, ctx().stream));
    FC_CUDA(cudaStreamSynchronize(ctx().stream));
    dfree(flag);
    if (v != 0.0) {
        dfree(s.ctrl);
        s.ctrl = nullptr;
        return false;
    }
    for (int q = 0; q < s.size; ++q) s.peer_ctrl[q] = reinterpret_cast<P2PCtrl*>(R.peer[q]);
    s.on = true;
    return true;
}

void p2p_finalize()
{
    State& s = S();
    if (!s.on) return;
    cudaStreamSynchronize(ctx().stream);
    for (Region& R : s.regions)
        for (int q = 0; q < s.size; ++q)
            if (q != s.rank && R.peer[q]) cudaIpcCloseMemHandle(R.peer[q]);
    s.regions.clear();
    for (int q = 0; q < s.size; ++q)
        if (q != s.rank && s.peer_ctrl[q]) cudaIpcCloseMemHandle(s.peer_ctrl[q]);
    dfree(s.ctrl);
    s = State();
}

void p2p_register(void* base, size_t bytes)
{
    State& s = S();
    if (!s.on) return;
    Region R{static_cast<char*>(base), bytes, {nullptr}};
    if (!map_region(R)) fail(ERROR_SOLVER_MISC, "CUDA IPC mapping of a solver buffer failed");
    s.regions.push_back(R);
}

void p2p_unregister(void* base)
{
    State& s = S();
    if (!s.on) return;
    for (size_t i = 0; i < s.regions.size(); ++i)
        if (s.regions[i].base == base) {
            cudaStreamSynchronize(ctx().stream);
            for (int q = 0; q < s.size; ++q)
                if (q != s.rank && s.regions[i].peer[q]) cudaIpcCloseMemHandle(s.regions[i].peer[q]);
            s.regions.erase(s.regions.begin() + i);
            return;
        }
}

bool p2p_lookup(const void* p, double* (&peer)[P2P_MAX_RANKS])
{
    State&      s = S();
    const char* c = static_cast<const char*>(p);
    for (const Region& R : s.regions)
        if (c >= R.base && c < R.base + R.bytes) {
            for (int q = 0; q < s.size; ++q) peer[q] = reinterpret_cast<double*>(R.peer[q] + (c - R.base));
            return true;
        }
    return false;
}

// ------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// one CTA, one thread per rank: announce my next epoch to everybody, wait for everybody's
__device__ __forceinline__ unsigned long long barrier_body(P2PCtrl* me, const PeerCtrls& peers, int rank, int n,
                                                           unsigned int mask = 0xffffffffu)
{
    __shared__ unsigned long long s_e;
    const int q = threadIdx.x;
    if (q == 0) s_e = me->epoch + 1;
    __syncthreads();
    const unsigned long long e = s_e;
    __threadfence_system();   // my earlier stores to peer memory are ordered before the flag
    if (q < n && ((mask >> q) & 1u)) {
        st_sys(&peers.p[q]->flags[rank], e);
        const long long t0 = clock64();
        while (ld_sys(&me->flags[q]) < e) {
            if (clock64() - t0 > SPIN_BUDGET) {
                me->error = 1;
                break;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (q == 0) me->epoch = e;
    return e;
}

__global__ void k_p2p_barrier(P2PCtrl* me, PeerCtrls peers, int rank, int n, unsigned int mask)
{
    barrier_body(me, peers, rank, n, mask);
}

struct PushDst {
    double* dst[P2P_MAX_RANKS];   // per send segment: where it lands in the peer's vector
    int     off[P2P_MAX_RANKS + 1];
    int     nseg;
};
__global__ void __launch_bounds__(256)
k_halo_push(int n, const int* __restrict__ idx, const double* __restrict__ x, PushDst d)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int s = 0;
        while (s + 1 < d.nseg && i >= d.off[s + 1]) ++s;
        d.dst[s][i - d.off[s]] = x[idx[i]];
    }
}

// Fused ghost exchange, exchange number s = seq + 1 (every rank runs the same sequence):
//   1. tell the ranks that push to me that everything up to s-1 has been consumed here (all my
//      earlier kernels are complete: stream order);
//   2. make sure the receivers' ghost regions may be overwritten. For a receiver q that also SENT
//      to me in exchange s-1 (`throttle_mask`) it is enough to see its data flag of s-1 (a local
//      poll, normally long satisfied): q finished its push s-1, hence its consumer s-2 and all
//      earlier ones, and the host only puts q into this mask when exchange s targets a different
//      vector than s-1 -- the one q may still be reading. Every other receiver (`ack_mask`) must
//      have acknowledged s-1 (step 1 on its side); that handshake is off the critical path in the
//      common case and is skipped altogether right after an all-to-all barrier;
//   3. store my boundary entries straight into the peers' ghost regions;
//   4. the last CTA (ticket) publishes s in the receivers' dflags and advances seq.
// The consumer kernel waits for dflags[q] >= seq from its senders (p2p_halo_wait), only in the
// CTAs whose rows read ghosts. No ghost entry is overwritten before its reader has finished or
// read before it has arrived.
__global__ void __launch_bounds__(256)
k_halo_push_signal(P2PCtrl* me, PeerCtrls peers, int rank, int nranks, unsigned int send_mask,
                   unsigned int recv_mask, unsigned int ack_mask, unsigned int throttle_mask, int n,
                   const int* __restrict__ idx, const double* __restrict__ x, PushDst d)
{
    __shared__ int           s_last;
    const unsigned long long s = me->seq + 1;
    const int                q = threadIdx.x;
    if (q < nranks) {
        if (blockIdx.x == 0 && ((recv_mask >> q) & 1u)) st_sys(&peers.p[q]->acks[rank], s - 1);
        const bool wa = (ack_mask >> q) & 1u, wt = (throttle_mask >> q) & 1u;
        if (wa || wt) {
            const unsigned long long* w  = wa ? &me->acks[q] : &me->dflags[q];
            const long long           t0 = clock64();
            while (ld_sys(w) < s - 1) {
                if (clock64() - t0 > SPIN_BUDGET) {
                    me->error = 1;
                    break;
                }
            }
        }
    }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int sg = 0;
        while (sg + 1 < d.nseg && i >= d.off[sg + 1]) ++sg;
        d.dst[sg][i - d.off[sg]] = x[idx[i]];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();   // the CTA's stores (ordered before by the barrier) precede the ticket / the flag
        s_last = (atomicAdd(&me->ticket, 1u) == gridDim.x - 1) ? 1 : 0;
        if (s_last) __threadfence_system();
    }
    __syncthreads();
    if (!s_last) return;
    if (q < nranks && ((send_mask >> q) & 1u)) st_sys(&peers.p[q]->dflags[rank], s);
    if (q == 0) {
        me->ticket = 0;
        me->seq    = s;
    }
}

__global__ void k_p2p_allreduce(P2PCtrl* me, PeerCtrls peers, int rank, int n, double* buf, int count, int op)
{
    const int                q  = threadIdx.x;
    const unsigned long long e1 = me->epoch + 1;   // parity of the epoch this barrier will reach
    const int                par = (int)(e1 & 1ULL);
    if (q < n)
        for (int c = 0; c < count; ++c) peers.p[q]->slots[par][rank][c] = buf[c];
    barrier_body(me, peers, rank, n);
    if (q < count) {
        double acc = me->slots[par][0][q];
        for (int r = 1; r < n; ++r) {
            const double v = me->slots[par][r][q];
            acc            = (op == 2) ? (acc > v ? acc : v) : acc + v;
        }
        buf[q] = acc;
    }
}

struct GatherDst {
    double* dst[P2P_MAX_RANKS];
    int     n;
};
__global__ void __launch_bounds__(256) k_p2p_gather_push(const double* __restrict__ src, size_t cnt, GatherDst d)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (size_t)gridDim.x * blockDim.x) {
        const double v = src[i];
        for (int q = 0; q < d.n; ++q) d.dst[q][i] = v;
    }
}

static PeerCtrls peer_ctrls()
{
    PeerCtrls pc;
    for (int q = 0; q < P2P_MAX_RANKS; ++q) pc.p[q] = S().peer_ctrl[q];
    return pc;
}

void p2p_barrier()
{
    State& s = S();
    if (!s.on) return;
    FC_LAUNCH(k_p2p_barrier, 1, 32, 0, s.ctrl, peer_ctrls(), s.rank, s.size, 0xffffffffu);
    s.fused_synced  = true;
    s.prev_partners = 0xffffffffu;
}

// barrier with the exchange partners only (the epoch still advances on every rank)
static void p2p_barrier_masked(unsigned int mask)
{
    State& s = S();
    FC_LAUNCH(k_p2p_barrier, 1, 32, 0, s.ctrl, peer_ctrls(), s.rank, s.size, mask);
}

bool p2p_halo_exchange(const HaloPlan& h, double* x, HaloWait* wait)
{
    State& s = S();
    if (!s.on || !h.p2p_ready) return false;
    double* peer[P2P_MAX_RANKS];
    if (!p2p_lookup(x, peer)) return false;
    ProfScope prof(400, h.nloc, h.nsend + h.nghost, 8.0 * (h.nsend + h.nghost));
    if (wait && ctx().opt.p2p_fused == 1) {
        unsigned int send_mask = 0, recv_mask = 0;
        PushDst      d;
        d.nseg = (int)h.send_peer.size();
        for (int k = 0; k < d.nseg; ++k) {
            d.dst[k] = peer[h.send_peer[k]] + h.peer_dst_off[k];
            d.off[k] = h.send_off[k];
            send_mask |= 1u << h.send_peer[k];
        }
        d.off[d.nseg] = h.nsend;
        for (int q : h.recv_peer) recv_mask |= 1u << q;
        // always launched, even without neighbours: the exchange counter must advance in lockstep
        int g = (h.nsend + 255) / 256;
        if (g > 592) g = 592;
        if (g < 1) g = 1;
        unsigned int throttle = 0, ack = send_mask;
        if (s.fused_synced) {
            ack = 0;   // every rank passed a barrier after its last consumer
        } else if (s.fused_prev_x != nullptr && s.fused_prev_x != x) {
            throttle = send_mask & s.fused_prev_recv;
            ack      = send_mask & ~throttle;
        }
        s.fused_prev_x    = x;
        s.fused_prev_recv = recv_mask;
        s.fused_synced    = false;
        FC_LAUNCH(k_halo_push_signal, g, 256, 0, s.ctrl, peer_ctrls(), s.rank, s.size, send_mask, recv_mask,
                  ack, throttle, h.nsend, h.send_idx, x, d);
        wait->flags = s.ctrl->dflags;
        wait->seq   = &s.ctrl->seq;
        wait->err   = &s.ctrl->error;
        wait->mask  = recv_mask;
        wait->nint  = h.ngrow;
        for (int i = 0; i < h.ngrow; ++i) wait->lo[i] = h.grow_lo[i], wait->hi[i] = h.grow_hi[i];
        s.last_exchanged = nullptr;   // the next barrier-protocol user starts with a barrier
        return true;
    }
    // the previous exchange's barrier only guarantees that every rank finished the kernel BEFORE it;
    // pushing into the vector that kernel is still reading on a slower rank needs one more barrier
    unsigned int partners = 0;
    for (int q : h.send_peer) partners |= 1u << q;
    for (int q : h.recv_peer) partners |= 1u << q;
    const bool nb = ctx().opt.p2p_fused == 2;   // barrier with the exchange partners only
    if (s.last_exchanged == x || s.last_exchanged == nullptr || (nb && (partners & ~s.prev_partners))) p2p_barrier();
    s.last_exchanged = x;
    s.prev_partners  = nb ? partners : 0xffffffffu;
    if (h.nsend > 0) {
        PushDst d;
        d.nseg = (int)h.send_peer.size();
        for (int k = 0; k < d.nseg; ++k) {
            d.dst[k] = peer[h.send_peer[k]] + h.peer_dst_off[k];
            d.off[k] = h.send_off[k];
        }
        d.off[d.nseg] = h.nsend;
        int g = (h.nsend + 255) / 256;
        if (g > 592) g = 592;
        FC_LAUNCH(k_halo_push, g, 256, 0, h.nsend, h.send_idx, x, d);
    }
    if (nb) p2p_barrier_masked(partners);
    else p2p_barrier();
    return true;
}

void p2p_allreduce(double* buf, int count, int op)
{
    State& s = S();
    ProfScope prof(401, count, 0, 8.0 * count);
    FC_LAUNCH(k_p2p_allreduce, 1, 32, 0, s.ctrl, peer_ctrls(), s.rank, s.size, buf, count, op);
    s.fused_synced = true;
}

bool p2p_allgatherv(double* full, const std::vector<size_t>& counts, const std::vector<size_t>& displs)
{
    State& s = S();
    if (!s.on) return false;
    double* peer[P2P_MAX_RANKS];
    if (!p2p_lookup(full, peer)) return false;
    ProfScope prof(402, (int)counts[s.rank], 0, 8.0 * counts[s.rank]);
    // neighbour-only exchange barriers leave non-neighbours unsynchronised: meet everybody first
    if (s.last_exchanged == full || s.last_exchanged == nullptr || ctx().opt.p2p_fused == 2) p2p_barrier();
    s.last_exchanged = full;
    GatherDst d;
    d.n = 0;
    for (int q = 0; q < s.size; ++q)
        if (q != s.rank) d.dst[d.n++] = peer[q] + displs[s.rank];
    const size_t cnt = counts[s.rank];
    if (cnt > 0 && d.n > 0) {
        int g = (int)((cnt + 255) / 256);
        if (g > 592) g = 592;
        FC_LAUNCH(k_p2p_gather_push, g, 256, 0, full + displs[s.rank], cnt, d);
    }
    p2p_barrier();
    return true;
}

} // namespace fc
