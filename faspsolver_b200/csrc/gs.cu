// gs.cu — multicolour Gauss-Seidel: FASP's data-parallel GS (the smoother its OpenMP build runs
// on every level, PreMGCycle.c:123-132). Colour classes come from the reference's greedy rule
// (BlaSparseCSR.c:1687-1770, restated in gs_multicolor_host): rows are visited through a circular
// queue; a row joins the current colour unless one of the rows already in it touches it, in which
// case it is re-queued; a new colour starts whenever the queue wraps to a smaller row index.
// One kernel per colour relaxes its rows in place (BlaSparseCSR.c:2123-2190):
//   u_i = (b_i - sum_{j != i} a_ij u_j) / a_ii , terms subtracted left to right.
#include "common.cuh"
#include "amg.cuh"

namespace fc {

void gs_multicolor_host(int n, const int* IA, const int* JA, std::vector<int>& IC, std::vector<int>& ICMAP)
{
    IC.clear();
    ICMAP.assign((size_t)n, 0);
    if (n == 0) {
        IC.push_back(0);
        return;
    }
    std::vector<int> queue((size_t)n + 1), touched((size_t)n + 1, -1);
    for (int k = 0; k < n; ++k) queue[k] = k;
    int front = n - 1, rear = n - 1, group = 0, count = 0, prev = 0;
    IC.push_back(0);
    do {
        if (++front == n) front = 0;
        const int i = queue[front];
        if (i <= prev) {                       // wrapped around: open a new colour with row i
            if ((int)IC.size() <= group) IC.push_back(count);
            IC[group]      = count;
            ICMAP[count++] = i;
            ++group;
            for (int j = IA[i]; j < IA[i + 1]; ++j) touched[JA[j]] = group;
        } else if (touched[i] == group) {      // coupled to the current colour: try again later
            if (++rear == n) rear = 0;
            queue[rear] = i;
        } else {                               // independent of the current colour: join it
            ICMAP[count++] = i;
            for (int j = IA[i]; j < IA[i + 1]; ++j) touched[JA[j]] = group;
        }
        prev = i;
    } while (rear != front);
    IC.resize((size_t)group + 1);
    IC[group] = count;
}

__global__ void __launch_bounds__(128)
k_gs_color(int nrows, const int* __restrict__ rows, const int* __restrict__ ia,
           const int* __restrict__ ja, const double* __restrict__ val, const double* __restrict__ b,
           double* u, const int* done)
{
    if (done && *done) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nrows) return;
    const int i = rows[t];
    double    acc = b[i], d = 0.0;
    for (int k = ia[i]; k < ia[i + 1]; ++k) {
        const int j = ja[k];
        if (j != i) acc = __dsub_rn(acc, __dmul_rn(val[k], u[j]));
        else d = val[k];
    }
    if (fabs(d) > SMALLREAL) u[i] = __ddiv_rn(acc, d);
}

void gs_multicolor_sweeps(const DevCSR& A, const int* color_rows, const std::vector<int>& color_ptr,
                          const double* b, double* u, int L, int order, const int* done)
{
    const int ncol = (int)color_ptr.size() - 1;
    while (L-- > 0) {
        for (int cc = 0; cc < ncol; ++cc) {
            const int c  = (order == -1) ? ncol - 1 - cc : cc;
            const int n  = color_ptr[c + 1] - color_ptr[c];
            if (n <= 0) continue;
            ProfScope prof(300, n, 0, 0.0);
            FC_LAUNCH(k_gs_color, (n + 127) / 128, 128, 0, n, color_rows + color_ptr[c], A.ia, A.ja, A.val, b, u,
                      done);
        }
    }
}

} // namespace fc
