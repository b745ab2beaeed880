// reduce.cuh — deterministic fused grid reduction used by every kernel that produces a
// scalar (dot products, norms): per-thread values -> warp shuffle -> CTA -> per-CTA partial
// in global memory. CTAs are grouped by GROUP = 256 consecutive block ids: the last CTA of a
// group to arrive (atomic ticket) adds the group's partials in block order and publishes a
// group partial; the last group to finish adds the group partials in group order and hands
// the totals to a finalize functor running on one thread. The scalar therefore never leaves
// the device, the summation order is fixed for a given launch shape, and no CTA ever reads
// more than max(256, #groups) partials (a 65536-CTA launch ends with two 256-wide sums).
#pragma once
#include "common.cuh"

namespace fc {

constexpr unsigned int RED_GROUP = 256;

// value slots per CTA and the scratch layout (doubles):
//   partials[ b * NS + s ]                       per-CTA partials, b < gridDim.x
//   partials[ (gridDim.x + g) * NS + s ]         per-group partials, g < ngroups
// tickets (unsigned): tickets[0] = finished groups, tickets[1 + g] = arrivals in group g.
// red_partials()/red_ticket() reserve 4 doubles per CTA (+ groups) and 1 + grid/256 tickets.
template <int NS, int MAXMASK>
__device__ __forceinline__ double red_combine(int s, double a, double b)
{
    return (MAXMASK >> s & 1) ? (a > b ? a : b) : a + b;
}

// NS values per thread; slot s is a max-reduction instead of a sum when MAXMASK has bit s.
// All threads of the CTA must call this. blockDim.x: a multiple of 32, at most 256.
template <int NS, int MAXMASK, class Fin>
__device__ __forceinline__ void grid_reduce(double (&v)[NS], double* partials,
                                            unsigned int* tickets, Fin fin)
{
    __shared__ double s_part[NS][8];
    __shared__ int    s_role;   // 0: done, 1: reduce my group, 2: (after 1) reduce the groups
    const int          tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned int nb  = gridDim.x;
    const unsigned int ng  = (nb + RED_GROUP - 1) / RED_GROUP;
    const unsigned int g   = blockIdx.x / RED_GROUP;
    const int          nw  = blockDim.x >> 5;

    // ---- CTA-level reduction of the per-thread values
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        double x = v[s];
        for (int off = 16; off > 0; off >>= 1)
            x = red_combine<NS, MAXMASK>(s, x, __shfl_xor_sync(0xffffffffu, x, off));
        if (lane == 0) s_part[s][wid] = x;
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            double x = s_part[s][0];
            for (int w = 1; w < nw; ++w) x = red_combine<NS, MAXMASK>(s, x, s_part[s][w]);
            partials[(size_t)blockIdx.x * NS + s] = x;
        }
        __threadfence();
        const unsigned int first = g * RED_GROUP;
        const unsigned int gsize = (nb - first < RED_GROUP) ? nb - first : RED_GROUP;
        s_role = (atomicAdd(tickets + 1 + g, 1u) == gsize - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_role == 0) return;

    // ---- last CTA of its group: add the group's partials in block order
    __threadfence();
    {
        const unsigned int first = g * RED_GROUP;
        const unsigned int gsize = (nb - first < RED_GROUP) ? nb - first : RED_GROUP;
        double t[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            double x = (MAXMASK >> s & 1) ? -1.0e300 : 0.0;
            for (unsigned int i = tid; i < gsize; i += blockDim.x)
                x = red_combine<NS, MAXMASK>(s, x, __ldcg(partials + (size_t)(first + i) * NS + s));
            for (int off = 16; off > 0; off >>= 1)
                x = red_combine<NS, MAXMASK>(s, x, __shfl_xor_sync(0xffffffffu, x, off));
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
                for (int w = 1; w < nw; ++w) x = red_combine<NS, MAXMASK>(s, x, s_part[s][w]);
                partials[(size_t)(nb + g) * NS + s] = x;
            }
            tickets[1 + g] = 0u;
            __threadfence();
            s_role = (atomicAdd(tickets, 1u) == ng - 1) ? 2 : 0;
        }
        __syncthreads();
        if (s_role != 2) return;
    }

    // ---- last group: add the group partials in group order, finalize
    __threadfence();
    {
        double t[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            double x = (MAXMASK >> s & 1) ? -1.0e300 : 0.0;
            for (unsigned int i = tid; i < ng; i += blockDim.x)
                x = red_combine<NS, MAXMASK>(s, x, __ldcg(partials + (size_t)(nb + i) * NS + s));
            for (int off = 16; off > 0; off >>= 1)
                x = red_combine<NS, MAXMASK>(s, x, __shfl_xor_sync(0xffffffffu, x, off));
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
                for (int w = 1; w < nw; ++w) x = red_combine<NS, MAXMASK>(s, x, s_part[s][w]);
                t[s] = x;
            }
            tickets[0] = 0u;
            fin(t);
        }
    }
}

} // namespace fc
