// comm.cu — NCCL plumbing (see comm.cuh). NVLink 5 / NVSwitch: every peer at full bandwidth,
// so the cost of the small messages on this path (halo faces, 1-3 scalars) is latency, not
// link count; all operations are enqueued on the library stream and are graph-capturable.
#include "comm.cuh"
#include "p2p.cuh"
#include <dlfcn.h>
#include <nccl.h>

namespace fc {

namespace {
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                        = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                 = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t)                                           = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t)                                           = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t,
                              cudaStream_t)                                           = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)       = nullptr;
    ncclResult_t (*GroupStart)()                                                      = nullptr;
    ncclResult_t (*GroupEnd)()                                                        = nullptr;
    const char* (*GetErrorString)(ncclResult_t)                                       = nullptr;
    ncclComm_t comm  = nullptr;
    int        rank  = 0, size = 1;
};
Nccl& N()
{
    static Nccl n;
    return n;
}
void load()
{
    Nccl& n = N();
    if (n.lib) return;
    const char* names[] = {getenv("FASP_CUDA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm) continue;
        n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n.lib) break;
    }
    if (!n.lib) fail(ERROR_SOLVER_MISC, "NCCL not found (dlopen libnccl.so.2): %s", dlerror());
#define FC_SYM(field, name)                                                                \
    n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.lib, name));                     \
    if (!n.field) fail(ERROR_SOLVER_MISC, "NCCL symbol %s missing", name)
    FC_SYM(GetUniqueId, "ncclGetUniqueId");
    FC_SYM(CommInitRank, "ncclCommInitRank");
    FC_SYM(CommDestroy, "ncclCommDestroy");
    FC_SYM(AllReduce, "ncclAllReduce");
    FC_SYM(Broadcast, "ncclBroadcast");
    FC_SYM(Send, "ncclSend");
    FC_SYM(Recv, "ncclRecv");
    FC_SYM(GroupStart, "ncclGroupStart");
    FC_SYM(GroupEnd, "ncclGroupEnd");
    FC_SYM(GetErrorString, "ncclGetErrorString");
#undef FC_SYM
}
void nccl_check(ncclResult_t r, const char* what)
{
    if (r != ncclSuccess) fail(ERROR_SOLVER_MISC, "NCCL error in %s: %s", what, N().GetErrorString(r));
}
} // namespace

int  comm_rank() { return N().rank; }
int  comm_size() { return N().size; }
bool comm_active() { return N().comm != nullptr && N().size > 1; }

void comm_unique_id(void* id128)
{
    load();
    ncclUniqueId id;
    nccl_check(N().GetUniqueId(&id), "ncclGetUniqueId");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, sizeof(id));
}

void comm_init(const void* id128, int rank, int nranks)
{
    ensure_init();
    load();
    Nccl& n = N();
    if (n.comm) fail(ERROR_INPUT_PAR, "communicator already initialised");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    nccl_check(n.CommInitRank(&n.comm, nranks, id, rank), "ncclCommInitRank");
    n.rank = rank;
    n.size = nranks;
    p2p_init();   // peer-memory data path when CUDA IPC is available (else NCCL everywhere)
}

void comm_finalize()
{
    Nccl& n = N();
    if (n.comm) {
        cudaStreamSynchronize(ctx().stream);
        p2p_finalize();
        n.CommDestroy(n.comm);
    }
    n.comm = nullptr;
    n.rank = 0;
    n.size = 1;
}

void comm_allreduce_mixed(double* buf, int count, int maxmask, const int* gate)
{
    if (!comm_active()) return;
    if (p2p_active() && count <= 4) {
        p2p_allreduce(buf, count, maxmask, gate);
        return;
    }
    // NCCL: one call per run of equal operations
    int s = 0;
    while (s < count) {
        int       e  = s + 1;
        const int mx = (maxmask >> s) & 1;
        while (e < count && ((maxmask >> e) & 1) == mx) ++e;
        comm_allreduce(buf + s, (size_t)(e - s), mx ? 2 : 0, gate);
        s = e;
    }
}

void comm_allreduce(double* buf, size_t count, int op, const int* gate)
{
    if (!comm_active()) return;
    if (p2p_active() && count <= 4) {
        p2p_allreduce(buf, (int)count, op == 2 ? 0xF : 0, gate);
        return;
    }
    Ctx& c = ctx();
    ProfScope prof(401, (int)count, 0, 8.0 * count);
    nccl_check(N().AllReduce(buf, buf, count, ncclDouble, op == 2 ? ncclMax : ncclSum, N().comm, c.stream),
               "ncclAllReduce");
}

void comm_allgatherv(const double* send, size_t sendcount, double* recv, const std::vector<size_t>& counts,
                     const std::vector<size_t>& displs, const int* gate)
{
    Ctx&  c = ctx();
    Nccl& n = N();
    if (!comm_active()) {
        if (recv + displs[0] != send)
            FC_CUDA(cudaMemcpyAsync(recv + displs[0], send, sizeof(double) * sendcount, cudaMemcpyDeviceToDevice,
                                    c.stream));
        return;
    }
    if (send == recv + displs[n.rank] && p2p_allgatherv(recv, counts, displs, gate)) return;
    // variable counts: one broadcast per root inside a group (fused by NCCL into one launch)
    ProfScope prof(402, (int)sendcount, 0, 8.0 * sendcount);
    nccl_check(n.GroupStart(), "ncclGroupStart");
    for (int r = 0; r < n.size; ++r)
        nccl_check(n.Broadcast(r == n.rank ? (const void*)send : (const void*)(recv + displs[r]),
                               recv + displs[r], counts[r], ncclDouble, r, n.comm, c.stream),
                   "ncclBroadcast");
    nccl_check(n.GroupEnd(), "ncclGroupEnd");
}

void comm_group_start()
{
    if (comm_active()) nccl_check(N().GroupStart(), "ncclGroupStart");
}
void comm_group_end()
{
    if (comm_active()) nccl_check(N().GroupEnd(), "ncclGroupEnd");
}
void comm_send(const double* buf, size_t count, int peer)
{
    nccl_check(N().Send(buf, count, ncclDouble, peer, N().comm, ctx().stream), "ncclSend");
}
void comm_recv(double* buf, size_t count, int peer)
{
    nccl_check(N().Recv(buf, count, ncclDouble, peer, N().comm, ctx().stream), "ncclRecv");
}

} // namespace fc
