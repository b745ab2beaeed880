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
void p2p_reset_order_hook() { S().last_exchanged = nullptr; }

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
void p2p_reset_order() { S().last_exchanged = nullptr; }
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
    // "the vector exchanged last" is compared by LOCAL address; across solver lifetimes the allocator may hand the
    // address of a freed vector to a new one on one rank and not on another, and the ranks would then disagree on
    // the extra barrier (epochs out of step: bench line followed by the slab solver, r02g). Registration is
    // collective, so forgetting the vector here keeps the ranks' decisions identical.
    s.last_exchanged = nullptr;
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
            s.last_exchanged = nullptr;
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
__device__ __forceinline__ unsigned long long barrier_body(P2PCtrl* me, const PeerCtrls& peers, int rank, int n)
{
    __shared__ unsigned long long s_e;
    const int q = threadIdx.x;
    if (q == 0) s_e = me->epoch + 1;
    __syncthreads();
    const unsigned long long e = s_e;
    __threadfence_system();   // my earlier stores to peer memory are ordered before the flag
    if (q < n) {
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

__global__ void k_p2p_barrier(P2PCtrl* me, PeerCtrls peers, int rank, int n, const int* gate)
{
    if (gate != nullptr && *gate != 0) return;   // the same decision on every rank
    barrier_body(me, peers, rank, n);
}

struct PushDst {
    double* dst[P2P_MAX_RANKS];   // per send segment: where it lands in the peer's vector
    int     off[P2P_MAX_RANKS + 1];
    int     nseg;
};
__global__ void __launch_bounds__(256)
k_halo_push(int n, const int* __restrict__ idx, const double* __restrict__ x, PushDst d, const int* gate)
{
    if (gate != nullptr && *gate != 0) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int s = 0;
        while (s + 1 < d.nseg && i >= d.off[s + 1]) ++s;
        d.dst[s][i - d.off[s]] = x[idx[i]];
    }
}

__global__ void k_p2p_allreduce(P2PCtrl* me, PeerCtrls peers, int rank, int n, double* buf, int count, int maxmask,
                                const int* gate)
{
    if (gate != nullptr && *gate != 0) return;
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
            acc            = ((maxmask >> q) & 1) ? (acc > v ? acc : v) : acc + v;
        }
        buf[q] = acc;
    }
}

struct GatherDst {
    double* dst[P2P_MAX_RANKS];
    int     n;
};
__global__ void __launch_bounds__(256)
k_p2p_gather_push(const double* __restrict__ src, size_t cnt, GatherDst d, const int* gate)
{
    if (gate != nullptr && *gate != 0) return;
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

void p2p_barrier(const int* gate)
{
    State& s = S();
    if (!s.on) return;
    FC_LAUNCH(k_p2p_barrier, 1, 32, 0, s.ctrl, peer_ctrls(), s.rank, s.size, gate);
}

bool p2p_halo_exchange(const HaloPlan& h, double* x, const int* gate, bool branch)
{
    State& s = S();
    if (!s.on || !h.p2p_ready) return false;
    double* peer[P2P_MAX_RANKS];
    if (!p2p_lookup(x, peer)) return false;
    ProfScope prof(400, h.nloc, h.nsend + h.nghost, 8.0 * (h.nsend + h.nghost));
    // the previous exchange's barrier only guarantees that every rank finished the kernel BEFORE it;
    // pushing into the vector that kernel is still reading on a slower rank needs one more barrier
    if (s.last_exchanged == x || s.last_exchanged == nullptr) p2p_barrier(gate);
    // an exchange behind a branch flag may not run: whatever follows must not count on it
    s.last_exchanged = branch ? nullptr : x;
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
        FC_LAUNCH(k_halo_push, g, 256, 0, h.nsend, h.send_idx, x, d, gate);
    }
    p2p_barrier(gate);
    return true;
}

void p2p_allreduce(double* buf, int count, int maxmask, const int* gate)
{
    State& s = S();
    ProfScope prof(401, count, 0, 8.0 * count);
    FC_LAUNCH(k_p2p_allreduce, 1, 32, 0, s.ctrl, peer_ctrls(), s.rank, s.size, buf, count, maxmask, gate);
}

bool p2p_allgatherv(double* full, const std::vector<size_t>& counts, const std::vector<size_t>& displs,
                    const int* gate)
{
    State& s = S();
    if (!s.on) return false;
    double* peer[P2P_MAX_RANKS];
    if (!p2p_lookup(full, peer)) return false;
    ProfScope prof(402, (int)counts[s.rank], 0, 8.0 * counts[s.rank]);
    if (s.last_exchanged == full || s.last_exchanged == nullptr) p2p_barrier(gate);
    s.last_exchanged = full;
    GatherDst d;
    d.n = 0;
    for (int q = 0; q < s.size; ++q)
        if (q != s.rank) d.dst[d.n++] = peer[q] + displs[s.rank];
    const size_t cnt = counts[s.rank];
    if (cnt > 0 && d.n > 0) {
        int g = (int)((cnt + 255) / 256);
        if (g > 592) g = 592;
        FC_LAUNCH(k_p2p_gather_push, g, 256, 0, full + displs[s.rank], cnt, d, gate);
    }
    p2p_barrier(gate);
    return true;
}

} // namespace fc
