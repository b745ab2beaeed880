// reduce.cuh — deterministic fused grid reduction used by every kernel that produces a
// scalar (dot products, norms): per-thread values -> warp shuffle -> CTA -> per-CTA partial
// in global memory -> the last CTA to arrive (atomic ticket) adds the partials in CTA order
// and hands the totals to a finalize functor running on one thread. The scalar therefore
// stays on the device and the summation order is fixed for a given launch shape.
#pragma once
#include "common.cuh"

namespace fc {

// NS sums per CTA; slot s is a max-reduction instead of a sum when MAXMASK has bit s.
// All threads of the CTA must call this. blockDim.x must be a multiple of 32, <= 1024.
template <int NS, int MAXMASK, class Fin>
__device__ __forceinline__ void grid_reduce(double (&v)[NS], double* partials,
                                            unsigned int* ticket, Fin fin)
{
    __shared__ double s_part[NS][32];
    __shared__ bool   s_is_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;

#pragma unroll
    for (int s = 0; s < NS; ++s) {
        double x = v[s];
        for (int off = 16; off > 0; off >>= 1) {
            const double y = __shfl_xor_sync(0xffffffffu, x, off);
            x              = (MAXMASK >> s & 1) ? (x > y ? x : y) : x + y;
        }
        if (lane == 0) s_part[s][wid] = x;
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            double x = s_part[s][0];
            for (int w = 1; w < nw; ++w)
                x = (MAXMASK >> s & 1) ? (x > s_part[s][w] ? x : s_part[s][w]) : x + s_part[s][w];
            partials[(size_t)blockIdx.x * NS + s] = x;
        }
        __threadfence();
        s_is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    double t[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        double x = (MAXMASK >> s & 1) ? -1.0e300 : 0.0;
        for (unsigned int i = tid; i < gridDim.x; i += blockDim.x) {
            const double y = __ldcg(partials + (size_t)i * NS + s);
            x              = (MAXMASK >> s & 1) ? (x > y ? x : y) : x + y;
        }
        for (int off = 16; off > 0; off >>= 1) {
            const double y = __shfl_xor_sync(0xffffffffu, x, off);
            x              = (MAXMASK >> s & 1) ? (x > y ? x : y) : x + y;
        }
        t[s] = x;
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) s_part[s][wid] = t[s];
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            double x = s_part[s][0];
            for (int w = 1; w < nw; ++w)
                x = (MAXMASK >> s & 1) ? (x > s_part[s][w] ? x : s_part[s][w]) : x + s_part[s][w];
            t[s] = x;
        }
        *ticket = 0u;
        fin(t);
    }
}

} // namespace fc
