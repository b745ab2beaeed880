// context.cu — process context, error plumbing, device memory, reduction scratch.
#include "common.cuh"
#include <cstdarg>
#include <mutex>

namespace fc {

static thread_local std::string g_last_error;

void        set_last_error(const std::string& s) { g_last_error = s; }
const char* last_error() { return g_last_error.c_str(); }

void fail(int code, const char* fmt, ...)
{
    char    buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw Error{code, std::string(buf)};
}

Ctx& ctx()
{
    static Ctx c;
    return c;
}

void ensure_init()
{
    Ctx& c = ctx();
    if (c.inited) return;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        fail(ERROR_SOLVER_MISC,
             "libfasp_cuda: no CUDA device available (%s) - this library has no CPU fallback",
             e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    int dev = c.device;
    const char* lr = getenv("LOCAL_RANK");
    if (!c.inited && lr != nullptr && getenv("FASP_CUDA_DEVICE") == nullptr && dev == 0)
        dev = atoi(lr) % ndev;
    if (const char* fd = getenv("FASP_CUDA_DEVICE")) dev = atoi(fd);
    if (dev < 0 || dev >= ndev) fail(ERROR_INPUT_PAR, "device %d out of range (0..%d)", dev, ndev - 1);
    FC_CUDA(cudaSetDevice(dev));
    c.device = dev;
    cudaDeviceProp p;
    FC_CUDA(cudaGetDeviceProperties(&p, dev));
    c.sm_count = p.multiProcessorCount;
    c.l2_bytes = (size_t)p.l2CacheSize;
    // the main stream carries the ghost pushes and flag barriers of the multi-GPU solve: highest priority, so
    // that their few CTAs are placed ahead of the pending CTAs of the interior kernels on the side stream
    int prio_lo = 0, prio_hi = 0;
    FC_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    if (const char* e = getenv("FASP_CUDA_STREAM_PRIO")) {   // 0: both streams at the default priority (A/B measurements)
        if (atoi(e) == 0) prio_lo = prio_hi = 0;
    }
    FC_CUDA(cudaStreamCreateWithPriority(&c.stream, cudaStreamNonBlocking, prio_hi));
    FC_CUDA(cudaStreamCreateWithPriority(&c.side, cudaStreamNonBlocking, prio_lo));
    c.launch_stream = c.stream;
    FC_CUDA(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
    FC_CUDA(cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
    c.red_ticket = dalloc<unsigned int>((size_t)1 << 20);   // [0] groups done, [1+g] arrivals in group g
    FC_CUDA(cudaMemset(c.red_ticket, 0, sizeof(unsigned int) << 20));
    c.red_ticket_side = dalloc<unsigned int>((size_t)1 << 16);
    FC_CUDA(cudaMemset(c.red_ticket_side, 0, sizeof(unsigned int) << 16));
    c.side_tot = dalloc<double>(4);
    FC_CUDA(cudaMemset(c.side_tot, 0, 4 * sizeof(double)));
    c.inited = true;
    if (const char* s = getenv("FASP_CUDA_STRICT")) c.opt.strict = atoi(s);
    if (const char* s = getenv("FASP_CUDA_GRAPH")) c.opt.graph = atoi(s);
    if (const char* s = getenv("FASP_CUDA_COARSE_DENSE")) c.opt.coarse_dense = atoi(s);
    if (const char* s = getenv("FASP_CUDA_ZERO_GUESS")) c.opt.zero_guess = atoi(s);
}

void* dmalloc(size_t bytes)
{
    if (bytes == 0) bytes = 8;
    void*       p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(ERROR_ALLOC_MEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    return p;
}

void dfree(void* p)
{
    if (p) cudaFree(p);
}

ProfScope::ProfScope(int kind, int rows, long long nnz, double bytes)
{
    Ctx& c = ctx();
    on     = c.opt.profile && !c.capturing;
    if (!on) return;
    Ctx::ProfRec r;
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    r.kind = kind, r.rows = rows, r.nnz = nnz, r.bytes = bytes;
    cudaEventRecord(r.e0, c.stream);
    idx = c.prof.size();
    c.prof.push_back(r);
}
ProfScope::~ProfScope()
{
    if (on) cudaEventRecord(ctx().prof[idx].e1, ctx().stream);
}

// The partial buffer only ever grows, and superseded buffers stay alive: captured CUDA
// graphs may still hold their address.
double* red_partials(size_t nblocks)
{
    Ctx&   c    = ctx();
    size_t need = 4 * (nblocks + nblocks / 256 + 2) + 16;   // up to 4 sums per CTA + per group
    const bool side = (c.launch_stream == c.side) && c.side != nullptr;
    if (nblocks / 256 + 2 > ((size_t)1 << (side ? 16 : 20))) fail(ERROR_MAT_SIZE, "grid too large for the reduction tickets");
    double*& buf = side ? c.red_partials_side : c.red_partials;
    size_t&  cap = side ? c.red_cap_side : c.red_cap;
    if (need > cap) {
        if (c.capturing)
            fail(ERROR_SOLVER_MISC, "reduction scratch must be reserved before graph capture");
        size_t ncap = cap ? cap : (size_t)1 << 16;
        while (ncap < need) ncap *= 2;
        buf = dalloc<double>(ncap);   // old buffer intentionally kept
        cap = ncap;
    }
    return buf;
}

unsigned int* red_ticket()
{
    Ctx& c = ctx();
    return (c.launch_stream == c.side && c.side != nullptr) ? c.red_ticket_side : c.red_ticket;
}

__global__ void flush_kernel(char* buf, size_t n, char v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = i * 16; k + 16 <= n; k += stride * 16)
        *reinterpret_cast<int4*>(buf + k) = make_int4(v, v, v, v);
}

void flush_l2()
{
    Ctx& c = ctx();
    if (!c.flush_buf) {
        c.flush_bytes = 2 * c.l2_bytes > ((size_t)256 << 20) ? 2 * c.l2_bytes : ((size_t)256 << 20);
        c.flush_buf   = static_cast<char*>(dmalloc(c.flush_bytes));
    }
    static char v = 0;
    flush_kernel<<<c.sm_count * 8, 256, 0, c.stream>>>(c.flush_buf, c.flush_bytes, ++v);
    FC_CUDA(cudaPeekAtLastError());
}

} // namespace fc
