// setup.cu — the two matrix-matrix pieces of FASP's AMG setup on the device:
//
//   fasp_dcsr_trans      BlaSparseCSR.c:952-1019   R = P^T   (counting transpose, rows of the result ascending)
//   fasp_blas_dcsr_rap   BlaSpmvCSR.c:999-1250     A_c = R A P  (SMMP-style Galerkin triple product)
//
// Both reproduce the reference's OUTPUT LAYOUT, not only its values, because everything downstream depends on
// it: the order of the entries inside a row of A_c is the summation order of every later SpMV / smoother, and
// it decides the next level's strength graph traversal. The CPU code builds row ic of R A P by walking
// R(ic,:) -> A(i1,:) -> P(i2,:) and appends a column the first time it is met (the diagonal ic is always
// entry 0); a value is the sum of its products in walking order, each product formed as (r*a)*p with separate
// roundings. Rows are independent, so here ONE THREAD owns a coarse row and does exactly that walk with an
// open-addressing hash table (column -> position in the row) in place of the CPU's dense marker arrays:
// the first-seen order, the accumulation order and every rounding are those of the CPU loop, hence the result
// is identical bit for bit. Two passes (count, then fill) as in the reference; rows are processed in chunks so
// that the hash tables fit a fixed scratch budget.
//
// The transpose counts the entries per column with atomics, scans, scatters with atomic cursors (arrival order
// is arbitrary) and then sorts every row of the result by column index — the column indices of a row of A^T are
// distinct row numbers of A, so the sorted row is unique and equals the reference's (which fills in ascending
// row order).
#include "common.cuh"
#include <algorithm>

namespace fc {

namespace {

struct DevBuf {
    void* p = nullptr;
    explicit DevBuf(size_t bytes) { p = dmalloc(bytes ? bytes : 8); }
    ~DevBuf() { dfree(p); }
    template <class T> T* as() { return static_cast<T*>(p); }
    DevBuf(const DevBuf&)            = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

struct HostCSRView {
    int           row, col, nnz;
    const int *   ia, *ja;
    const double* val;
};

struct DevCSRRaw {
    int     row = 0, col = 0, nnz = 0;
    int *   ia = nullptr, *ja = nullptr;
    double* val = nullptr;
    void upload(const dCSRmat* A)
    {
        Ctx& c = ctx();
        row = A->row, col = A->col, nnz = A->nnz;
        ia  = dalloc<int>((size_t)row + 1);
        ja  = dalloc<int>((size_t)(nnz ? nnz : 1));
        FC_CUDA(cudaMemcpyAsync(ia, A->IA, sizeof(int) * ((size_t)row + 1), cudaMemcpyHostToDevice, c.stream));
        if (nnz) FC_CUDA(cudaMemcpyAsync(ja, A->JA, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, c.stream));
        if (A->val && nnz) {
            val = dalloc<double>((size_t)nnz);
            FC_CUDA(cudaMemcpyAsync(val, A->val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, c.stream));
        }
    }
    ~DevCSRRaw()
    {
        dfree(ia);
        dfree(ja);
        dfree(val);
    }
};

// ---------------------------------------------------------------------------------------
// exclusive scan of n ints in place (three small kernels; setup phase, not a hot path)
// ---------------------------------------------------------------------------------------
constexpr int SC_T = 256, SC_ITEMS = 16, SC_TILE = SC_T * SC_ITEMS;

__global__ void __launch_bounds__(SC_T) k_scan_tile_sums(const int* __restrict__ d, long long n, long long* tile_sum)
{
    __shared__ long long s[SC_T / 32];
    const long long base = (long long)blockIdx.x * SC_TILE;
    long long       v    = 0;
    for (int e = 0; e < SC_ITEMS; ++e) {
        const long long i = base + (long long)threadIdx.x * SC_ITEMS + e;
        if (i < n) v += d[i];
    }
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < SC_T / 32; ++w) t += s[w];
        tile_sum[blockIdx.x] = t;
    }
}
__global__ void k_scan_tiles_serial(long long* tile_sum, int ntiles, long long* total)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long run = 0;
    for (int t = 0; t < ntiles; ++t) {
        const long long v = tile_sum[t];
        tile_sum[t]       = run;
        run += v;
    }
    *total = run;
}
__global__ void __launch_bounds__(SC_T) k_scan_apply(int* d, long long n, const long long* __restrict__ tile_off)
{
    __shared__ long long s[SC_T];
    const long long base = (long long)blockIdx.x * SC_TILE + (long long)threadIdx.x * SC_ITEMS;
    int             loc[SC_ITEMS];
    long long       sum = 0;
#pragma unroll
    for (int e = 0; e < SC_ITEMS; ++e) {
        loc[e] = (base + e < n) ? d[base + e] : 0;
        sum += loc[e];
    }
    s[threadIdx.x] = sum;
    __syncthreads();
    // exclusive prefix of the per-thread sums (Hillis-Steele on 256 values)
    for (int off = 1; off < SC_T; off <<= 1) {
        long long t = (threadIdx.x >= off) ? s[threadIdx.x - off] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    long long run = tile_off[blockIdx.x] + s[threadIdx.x] - sum;
#pragma unroll
    for (int e = 0; e < SC_ITEMS; ++e) {
        if (base + e < n) d[base + e] = (int)run;
        run += loc[e];
    }
}
// in place: d[i] <- sum_{j<i} d[j]; returns the total (host). Fails if the total exceeds INT32 (FASP's INT).
long long scan_exclusive(int* d, long long n)
{
    Ctx&      c      = ctx();
    const int ntiles = (int)((n + SC_TILE - 1) / SC_TILE);
    DevBuf    tiles(sizeof(long long) * ((size_t)ntiles + 1));
    long long* ts = tiles.as<long long>();
    FC_LAUNCH(k_scan_tile_sums, ntiles, SC_T, 0, d, n, ts);
    FC_LAUNCH(k_scan_tiles_serial, 1, 1, 0, ts, ntiles, ts + ntiles);
    FC_LAUNCH(k_scan_apply, ntiles, SC_T, 0, d, n, ts);
    long long total = 0;
    FC_CUDA(cudaMemcpyAsync(&total, ts + ntiles, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    FC_CUDA(cudaStreamSynchronize(c.stream));
    if (total > 2147483647LL) fail(ERROR_MAT_SIZE, "result has %lld entries: beyond FASP's 32-bit INT", total);
    return total;
}

// ---------------------------------------------------------------------------------------
// transpose
// ---------------------------------------------------------------------------------------
__global__ void k_tr_count(int nnz, const int* __restrict__ ja, int* cnt)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += gridDim.x * blockDim.x) atomicAdd(cnt + ja[k], 1);
}
__global__ void k_tr_scatter(int n, const int* __restrict__ ia, const int* __restrict__ ja, const double* __restrict__ val,
                             int* cursor, int* tja, double* tval)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = ia[i]; k < ia[i + 1]; ++k) {
        const int pos = atomicAdd(cursor + ja[k], 1);
        tja[pos]      = i;
        if (val) tval[pos] = val[k];
    }
}
// every row of A^T sorted by column index (= row index of A): insertion sort, rows are short
__global__ void k_tr_sort_rows(int m, const int* __restrict__ tia, int* tja, double* tval)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const int a = tia[r], b = tia[r + 1];
    for (int k = a + 1; k < b; ++k) {
        const int    key = tja[k];
        const double kv  = tval ? tval[k] : 0.0;
        int          p   = k - 1;
        while (p >= a && tja[p] > key) {
            tja[p + 1] = tja[p];
            if (tval) tval[p + 1] = tval[p];
            --p;
        }
        tja[p + 1] = key;
        if (tval) tval[p + 1] = kv;
    }
}

// ---------------------------------------------------------------------------------------
// R A P
// ---------------------------------------------------------------------------------------
struct RapIn {
    int           nc, nf;
    const int *   Ri, *Rj;
    const double* Rv;
    const int *   Ai, *Aj;
    const double* Av;
    const int *   Pi, *Pj;
    const double* Pv;
};

// upper bound of the entries of coarse row ic: every product counted, plus the diagonal
__global__ void k_rap_bound(RapIn in, int* ub)
{
    const int ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (ic >= in.nc) return;
    long long s = 1;
    for (int j1 = in.Ri[ic]; j1 < in.Ri[ic + 1]; ++j1) {
        const int i1 = in.Rj[j1];
        for (int j2 = in.Ai[i1]; j2 < in.Ai[i1 + 1]; ++j2) {
            const int i2 = in.Aj[j2];
            s += in.Pi[i2 + 1] - in.Pi[i2];
        }
    }
    if (s > in.nc) s = in.nc;
    ub[ic] = (int)s;
}

__device__ __forceinline__ unsigned int rap_hash(int key, unsigned int mask)
{
    return ((unsigned int)key * 2654435761u >> 7) & mask;
}

// table slot: key (column, -1 = empty) and its position in the row. Returns the position; *fresh tells whether
// the column was met for the first time.
__device__ __forceinline__ int rap_lookup(int2* tab, unsigned int mask, int key, int& counter, bool& fresh)
{
    unsigned int h = rap_hash(key, mask);
    while (true) {
        const int2 e = tab[h];
        if (e.x == key) {
            fresh = false;
            return e.y;
        }
        if (e.x == -1) {
            tab[h] = make_int2(key, counter);
            fresh  = true;
            return counter++;
        }
        h = (h + 1) & mask;
    }
}

// FILL == false: count the distinct columns of every row of the chunk [row0, row0 + nrows)
// FILL == true : write RAP_j / RAP_data in the reference's order (BlaSpmvCSR.c:1203-1246)
template <bool FILL>
__global__ void __launch_bounds__(128)
k_rap_rows(RapIn in, int row0, int nrows, const long long* __restrict__ tab_off, int2* tabs, int* cnt,
           const int* __restrict__ rap_i, int* rap_j, double* rap_v)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nrows) return;
    const int          ic   = row0 + t;
    int2*              tab  = tabs + tab_off[t];
    const unsigned int mask = (unsigned int)(tab_off[t + 1] - tab_off[t]) - 1u;
    const int          base = FILL ? rap_i[ic] : 0;
    int                counter = 0;
    bool               fresh;
    // the diagonal entry comes first, value 0 until products arrive (:1204-1208)
    rap_lookup(tab, mask, ic, counter, fresh);
    if (FILL) {
        rap_j[base] = ic;
        rap_v[base] = 0.0;
    }
    for (int j1 = in.Ri[ic]; j1 < in.Ri[ic + 1]; ++j1) {
        const double r  = in.Rv ? in.Rv[j1] : 1.0;
        const int    i1 = in.Rj[j1];
        for (int j2 = in.Ai[i1]; j2 < in.Ai[i1 + 1]; ++j2) {
            const double ra = FILL ? __dmul_rn(r, in.Av[j2]) : 0.0;
            const int    i2 = in.Aj[j2];
            for (int j3 = in.Pi[i2]; j3 < in.Pi[i2 + 1]; ++j3) {
                const int i3  = in.Pj[j3];
                const int pos = rap_lookup(tab, mask, i3, counter, fresh);
                if (FILL) {
                    const double rap = __dmul_rn(ra, in.Pv ? in.Pv[j3] : 1.0);
                    if (fresh) {
                        rap_j[base + pos] = i3;
                        rap_v[base + pos] = rap;
                    } else {
                        rap_v[base + pos] = __dadd_rn(rap_v[base + pos], rap);
                    }
                }
            }
        }
    }
    if (!FILL) cnt[ic] = counter;
}

int grid_for(long long n, int block) { return (int)std::max<long long>(1, (n + block - 1) / block); }

} // namespace

// A^T on the device; host arrays of AT are allocated with calloc (FASP frees them with free())
void setup_transpose(const dCSRmat* A, dCSRmat* AT)
{
    ensure_init();
    Ctx&      c = ctx();
    const int n = A->row, m = A->col, nnz = A->nnz;
    AT->row = m, AT->col = n, AT->nnz = nnz;
    AT->IA  = static_cast<INT*>(calloc((size_t)m + 1, sizeof(INT)));
    AT->JA  = static_cast<INT*>(calloc((size_t)(nnz ? nnz : 1), sizeof(INT)));
    AT->val = A->val ? static_cast<REAL*>(calloc((size_t)(nnz ? nnz : 1), sizeof(REAL))) : nullptr;
    if (!AT->IA || !AT->JA || (A->val && !AT->val)) fail(ERROR_ALLOC_MEM, "fasp_cuda_dcsr_trans: host allocation failed");
    if (nnz == 0 || m == 0) return;
    DevCSRRaw dA;
    dA.upload(A);
    DevBuf tia(sizeof(int) * ((size_t)m + 2)), cur(sizeof(int) * ((size_t)m + 1)), tja(sizeof(int) * (size_t)nnz);
    DevBuf tval(A->val ? sizeof(double) * (size_t)nnz : 8);
    int*   d_tia = tia.as<int>();
    FC_CUDA(cudaMemsetAsync(d_tia, 0, sizeof(int) * ((size_t)m + 2), c.stream));
    FC_LAUNCH(k_tr_count, std::min(grid_for(nnz, 256), c.sm_count * 16), 256, 0, nnz, dA.ja, d_tia);
    scan_exclusive(d_tia, (long long)m + 1);   // d_tia[m] = nnz
    FC_CUDA(cudaMemcpyAsync(cur.p, d_tia, sizeof(int) * ((size_t)m + 1), cudaMemcpyDeviceToDevice, c.stream));
    FC_LAUNCH(k_tr_scatter, grid_for(n, 128), 128, 0, n, dA.ia, dA.ja, dA.val, cur.as<int>(), tja.as<int>(),
              A->val ? tval.as<double>() : nullptr);
    FC_LAUNCH(k_tr_sort_rows, grid_for(m, 128), 128, 0, m, d_tia, tja.as<int>(), A->val ? tval.as<double>() : nullptr);
    FC_CUDA(cudaMemcpyAsync(AT->IA, d_tia, sizeof(int) * ((size_t)m + 1), cudaMemcpyDeviceToHost, c.stream));
    FC_CUDA(cudaMemcpyAsync(AT->JA, tja.p, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost, c.stream));
    if (A->val) FC_CUDA(cudaMemcpyAsync(AT->val, tval.p, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, c.stream));
    FC_CUDA(cudaStreamSynchronize(c.stream));
}

// RAP = R A P on the device, the reference's layout bit for bit
void setup_rap(const dCSRmat* R, const dCSRmat* A, const dCSRmat* P, dCSRmat* RAP)
{
    ensure_init();
    Ctx& c = ctx();
    if (R->col != A->row || A->col != P->row || R->row != P->col)
        fail(ERROR_MAT_SIZE, "fasp_cuda_blas_dcsr_rap: incompatible shapes");
    if (!A->val) fail(ERROR_DATA_STRUCTURE, "fasp_cuda_blas_dcsr_rap: A has no values");
    const int nc = R->row;
    DevCSRRaw dR, dA, dP;
    dR.upload(R), dA.upload(A), dP.upload(P);
    RapIn in{nc, A->row, dR.ia, dR.ja, dR.val, dA.ia, dA.ja, dA.val, dP.ia, dP.ja, dP.val};
    DevBuf ub(sizeof(int) * ((size_t)nc + 1)), cnt(sizeof(int) * ((size_t)nc + 2));
    FC_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int) * ((size_t)nc + 2), c.stream));
    if (nc > 0) FC_LAUNCH(k_rap_bound, grid_for(nc, 128), 128, 0, in, ub.as<int>());
    std::vector<int> h_ub((size_t)nc);
    if (nc > 0) FC_CUDA(cudaMemcpyAsync(h_ub.data(), ub.p, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, c.stream));
    FC_CUDA(cudaStreamSynchronize(c.stream));
    // hash-table sizes: power of two >= 1.5 x bound; chunks of rows whose tables fit the scratch budget
    const long long budget = 1LL << 27;   // slots of 8 bytes: 1 GiB
    std::vector<long long> toff;
    std::vector<std::pair<int, int>> chunks;   // [row0, row1)
    {
        int r = 0;
        while (r < nc) {
            long long used = 0;
            int       e    = r;
            while (e < nc) {
                long long sz = 8;
                while (sz < (long long)h_ub[e] + h_ub[e] / 2 + 1) sz <<= 1;
                if (e > r && used + sz > budget) break;
                used += sz;
                ++e;
            }
            chunks.push_back({r, e});
            r = e;
        }
    }
    long long max_slots = 8;
    int       max_rows  = 1;
    for (auto& ch : chunks) {
        long long used = 0;
        for (int e = ch.first; e < ch.second; ++e) {
            long long sz = 8;
            while (sz < (long long)h_ub[e] + h_ub[e] / 2 + 1) sz <<= 1;
            used += sz;
        }
        max_slots = std::max(max_slots, used);
        max_rows  = std::max(max_rows, ch.second - ch.first);
    }
    DevBuf tabs(sizeof(int2) * (size_t)max_slots), d_toff(sizeof(long long) * ((size_t)max_rows + 1));
    auto run_pass = [&](bool fill, const int* rap_i, int* rap_j, double* rap_v) {
        for (auto& ch : chunks) {
            const int nrows = ch.second - ch.first;
            toff.assign((size_t)nrows + 1, 0);
            for (int e = 0; e < nrows; ++e) {
                long long sz = 8;
                const int u  = h_ub[ch.first + e];
                while (sz < (long long)u + u / 2 + 1) sz <<= 1;
                toff[e + 1] = toff[e] + sz;
            }
            FC_CUDA(cudaMemcpyAsync(d_toff.p, toff.data(), sizeof(long long) * ((size_t)nrows + 1), cudaMemcpyHostToDevice,
                                    c.stream));
            FC_CUDA(cudaMemsetAsync(tabs.p, 0xFF, sizeof(int2) * (size_t)toff[nrows], c.stream));   // keys = -1
            if (fill)
                FC_LAUNCH(k_rap_rows<true>, grid_for(nrows, 128), 128, 0, in, ch.first, nrows, d_toff.as<long long>(),
                          tabs.as<int2>(), cnt.as<int>(), rap_i, rap_j, rap_v);
            else
                FC_LAUNCH(k_rap_rows<false>, grid_for(nrows, 128), 128, 0, in, ch.first, nrows, d_toff.as<long long>(),
                          tabs.as<int2>(), cnt.as<int>(), rap_i, rap_j, rap_v);
            FC_CUDA(cudaStreamSynchronize(c.stream));   // toff is reused by the next chunk
        }
    };
    run_pass(false, nullptr, nullptr, nullptr);
    const long long size = scan_exclusive(cnt.as<int>(), (long long)nc + 1);   // cnt becomes RAP_i
    RAP->row = nc, RAP->col = nc, RAP->nnz = (INT)size;
    RAP->IA  = static_cast<INT*>(calloc((size_t)nc + 1, sizeof(INT)));
    RAP->JA  = static_cast<INT*>(calloc((size_t)(size ? size : 1), sizeof(INT)));
    RAP->val = static_cast<REAL*>(calloc((size_t)(size ? size : 1), sizeof(REAL)));
    if (!RAP->IA || !RAP->JA || !RAP->val) fail(ERROR_ALLOC_MEM, "fasp_cuda_blas_dcsr_rap: host allocation failed");
    DevBuf rj(sizeof(int) * (size_t)(size ? size : 1)), rv(sizeof(double) * (size_t)(size ? size : 1));
    run_pass(true, cnt.as<int>(), rj.as<int>(), rv.as<double>());
    FC_CUDA(cudaMemcpyAsync(RAP->IA, cnt.p, sizeof(int) * ((size_t)nc + 1), cudaMemcpyDeviceToHost, c.stream));
    if (size) {
        FC_CUDA(cudaMemcpyAsync(RAP->JA, rj.p, sizeof(int) * (size_t)size, cudaMemcpyDeviceToHost, c.stream));
        FC_CUDA(cudaMemcpyAsync(RAP->val, rv.p, sizeof(double) * (size_t)size, cudaMemcpyDeviceToHost, c.stream));
    }
    FC_CUDA(cudaStreamSynchronize(c.stream));
}

} // namespace fc
