// blas1.cu — vector kernels (replacing BlaArray.c:43-795, AuxArray.c:41,210, AuxVector.c:222).
// All results of reductions stay on the device; every kernel early-exits on *done.
#include "common.cuh"
#include "reduce.cuh"

namespace fc {

constexpr int VT = 256;

static inline int vec_grid(size_t n)
{
    size_t g   = (n + VT * 4 - 1) / (VT * 4);
    size_t cap = (size_t)ctx().sm_count * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

__global__ void __launch_bounds__(VT) k_set(double* x, double v, size_t n, const int* done)
{
    if (done && *done) return;
    for (size_t i = (size_t)blockIdx.x * VT + threadIdx.x; i < n; i += (size_t)gridDim.x * VT)
        x[i] = v;
}
__global__ void __launch_bounds__(VT)
k_copy(double* __restrict__ y, const double* __restrict__ x, size_t n, const int* done)
{
    if (done && *done) return;
    for (size_t i = (size_t)blockIdx.x * VT + threadIdx.x; i < n; i += (size_t)gridDim.x * VT)
        y[i] = x[i];
}
// y = a*x + b*y, FASP order of operations (BlaArray.c:620-660): y[i] = a*x[i] + b*y[i]
__global__ void __launch_bounds__(VT)
k_axpby(double a, const double* __restrict__ x, double b, double* __restrict__ y, size_t n,
        const int* done)
{
    if (done && *done) return;
    for (size_t i = (size_t)blockIdx.x * VT + threadIdx.x; i < n; i += (size_t)gridDim.x * VT)
        y[i] = __dadd_rn(__dmul_rn(a, x[i]), __dmul_rn(b, y[i]));
}
// y = (s*x)/d where |d| > SMALLREAL else 0 : the first smoother sweep from a zero guess
//   L1     (s = 1): u = 0 + (b - 0)/l1            ItrSmootherCSR.c:1560-1574
//   Jacobi (s = w): u = (1-w)*0 + w*b/d           ItrSmootherCSR.c:148-170
__global__ void __launch_bounds__(VT)
k_scale_div(double* __restrict__ y, double s, const double* __restrict__ x,
            const double* __restrict__ d, size_t n, Reduce red, double* partials,
            unsigned int* ticket, const int* done)
{
    if (done && *done) return;
    double v[2] = {0.0, 0.0};
    for (size_t i = (size_t)blockIdx.x * VT + threadIdx.x; i < n; i += (size_t)gridDim.x * VT) {
        const double di  = d[i];
        const double num = (s == 1.0) ? x[i] : __dmul_rn(s, x[i]);
        const double out = (fabs(di) > SMALLREAL) ? __ddiv_rn(num, di) : 0.0;
        y[i]             = out;
        if (red.dot_out) v[0] += out * red.dot_with[i];
        if (red.nrm2_out) v[1] += out * out;
    }
    if (red.dot_out || red.nrm2_out)
        grid_reduce<2, 0>(v, partials, ticket, [&](const double* t) {
            if (red.dot_out) *red.dot_out = t[0];
            if (red.nrm2_out) *red.nrm2_out = t[1];
        });
}
__global__ void __launch_bounds__(VT)
k_dot(const double* __restrict__ x, const double* __restrict__ y, size_t n, double* out,
      double* partials, unsigned int* ticket, const int* done)
{
    if (done && *done) return;
    double v[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * VT + threadIdx.x; i < n; i += (size_t)gridDim.x * VT)
        v[0] += x[i] * y[i];
    grid_reduce<1, 0>(v, partials, ticket, [&](const double* t) { *out = t[0]; });
}


// sums over an existing vector: dot with red.dot_with and/or squared 2-norm
__global__ void __launch_bounds__(VT)
k_reduce(const double* __restrict__ x, size_t n, Reduce red, double* partials,
         unsigned int* ticket, const int* done)
{
    if (done && *done) return;
    double v[2] = {0.0, 0.0};
    for (size_t i = (size_t)blockIdx.x * VT + threadIdx.x; i < n; i += (size_t)gridDim.x * VT) {
        const double xi = x[i];
        if (red.dot_out) v[0] += xi * red.dot_with[i];
        if (red.nrm2_out) v[1] += xi * xi;
    }
    grid_reduce<2, 0>(v, partials, ticket, [&](const double* t) {
        if (red.dot_out) *red.dot_out = t[0];
        if (red.nrm2_out) *red.nrm2_out = t[1];
    });
}
// y = d .* x
__global__ void __launch_bounds__(VT)
k_mul(double* __restrict__ y, const double* __restrict__ d, const double* __restrict__ x,
      size_t n, const int* done)
{
    if (done && *done) return;
    for (size_t i = (size_t)blockIdx.x * VT + threadIdx.x; i < n; i += (size_t)gridDim.x * VT)
        y[i] = __dmul_rn(d[i], x[i]);
}
// y += a*x with the a == 1 / a == -1 shortcuts of fasp_blas_darray_axpy (BlaArray.c:103-160)
__global__ void __launch_bounds__(VT)
k_axpy(double a, const double* __restrict__ x, double* __restrict__ y, size_t n, const int* done)
{
    if (done && *done) return;
    for (size_t i = (size_t)blockIdx.x * VT + threadIdx.x; i < n; i += (size_t)gridDim.x * VT) {
        const double xi = x[i], yi = y[i];
        y[i] = (a == 1.0) ? __dadd_rn(yi, xi) : (a == -1.0 ? __dsub_rn(yi, xi) : __dadd_rn(yi, __dmul_rn(a, xi)));
    }
}
// y += (*a) x, the scalar read from the device
__global__ void __launch_bounds__(VT)
k_axpy_dev(const double* a, const double* __restrict__ x, double* __restrict__ y, size_t n, const int* done)
{
    if (done && *done) return;
    const double al = *a;
    for (size_t i = (size_t)blockIdx.x * VT + threadIdx.x; i < n; i += (size_t)gridDim.x * VT)
        y[i] = __dadd_rn(y[i], __dmul_rn(al, x[i]));
}
// x *= a (fasp_blas_darray_ax, BlaArray.c:43: returns at once for a == 1)
__global__ void __launch_bounds__(VT) k_ax(double a, double* __restrict__ x, size_t n, const int* done)
{
    if (done && *done) return;
    for (size_t i = (size_t)blockIdx.x * VT + threadIdx.x; i < n; i += (size_t)gridDim.x * VT)
        x[i] = __dmul_rn(a, x[i]);
}
// out[0] = sum |x_i| (BlaArray.c:663), out[1] = max |x_i| (BlaArray.c:719)
__global__ void __launch_bounds__(VT)
k_norm1_inf(const double* __restrict__ x, size_t n, double* out, double* partials, unsigned int* ticket)
{
    double v[2] = {0.0, 0.0};
    for (size_t i = (size_t)blockIdx.x * VT + threadIdx.x; i < n; i += (size_t)gridDim.x * VT) {
        const double a = fabs(x[i]);
        v[0] += a;
        v[1] = fmax(v[1], a);
    }
    grid_reduce<2, 2>(v, partials, ticket, [&](const double* t) {
        out[0] = t[0];
        out[1] = t[1];
    });
}
// alpha = min(num/den, 1)  (coarse-grid scaling, PreMGCycle.c:210-216)
__global__ void k_scaling_alpha(double* s, const int* done)
{
    if (done && *done) return;
    const double a = s[1] / s[2];
    s[0]           = a < 1.0 ? a : 1.0;
}

void vec_set(double* x, double v, size_t n, const int* done)
{
    if (n == 0) return;
    FC_LAUNCH(k_set, vec_grid(n), VT, 0, x, v, n, done);
}
void vec_copy(double* y, const double* x, size_t n, const int* done)
{
    if (n == 0 || y == x) return;
    FC_LAUNCH(k_copy, vec_grid(n), VT, 0, y, x, n, done);
}
void vec_axpby(double a, const double* x, double b, double* y, size_t n, const int* done)
{
    if (n == 0) return;
    FC_LAUNCH(k_axpby, vec_grid(n), VT, 0, a, x, b, y, n, done);
}
void vec_axpy(double a, const double* x, double* y, size_t n, const int* done)
{
    if (n == 0) return;
    FC_LAUNCH(k_axpy, vec_grid(n), VT, 0, a, x, y, n, done);
}
void vec_axpy_dev(const double* a_dev, const double* x, double* y, size_t n, const int* done)
{
    if (n == 0) return;
    FC_LAUNCH(k_axpy_dev, vec_grid(n), VT, 0, a_dev, x, y, n, done);
}
void vec_ax(double a, double* x, size_t n, const int* done)
{
    if (n == 0 || a == 1.0) return;
    FC_LAUNCH(k_ax, vec_grid(n), VT, 0, a, x, n, done);
}
void vec_norm1_inf_host(const double* x, size_t n, double* norm1, double* norminf)
{
    Ctx&    c = ctx();
    double* d = dalloc<double>(2);
    const int g = vec_grid(n ? n : 1);
    FC_LAUNCH(k_norm1_inf, g, VT, 0, x, n, d, red_partials(g), red_ticket());
    double h[2] = {0.0, 0.0};
    FC_CUDA(cudaMemcpyAsync(h, d, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    FC_CUDA(cudaStreamSynchronize(c.stream));
    dfree(d);
    if (norm1) *norm1 = h[0];
    if (norminf) *norminf = h[1];
}
void vec_scale_div(double* y, double s, const double* x, const double* d, size_t n,
                   const Reduce& red, const int* done)
{
    if (n == 0) return;
    const int g = vec_grid(n);
    FC_LAUNCH(k_scale_div, g, VT, 0, y, s, x, d, n, red, red_partials(g), red_ticket(), done);
    reduce_finish(red, done);
}
void vec_dot(const double* x, const double* y, size_t n, double* out_dev, const int* done)
{
    const int g = vec_grid(n ? n : 1);
    FC_LAUNCH(k_dot, g, VT, 0, x, y, n, out_dev, red_partials(g), red_ticket(), done);
}


void vec_reduce(const double* x, size_t n, const Reduce& red, const int* done)
{
    if (!red.dot_out && !red.nrm2_out) return;
    const int g = vec_grid(n ? n : 1);
    FC_LAUNCH(k_reduce, g, VT, 0, x, n, red, red_partials(g), red_ticket(), done);
    reduce_finish(red, done);
}
void vec_mul(double* y, const double* d, const double* x, size_t n, const int* done)
{
    if (n == 0) return;
    FC_LAUNCH(k_mul, vec_grid(n), VT, 0, y, d, x, n, done);
}
void scaling_alpha(double* s, const int* done) { FC_LAUNCH(k_scaling_alpha, 1, 1, 0, s, done); }

double vec_dot_host(const double* x, const double* y, size_t n)
{
    Ctx&    c = ctx();
    double* d = dalloc<double>(1);
    vec_dot(x, y, n, d);
    double h = 0.0;
    FC_CUDA(cudaMemcpyAsync(&h, d, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    FC_CUDA(cudaStreamSynchronize(c.stream));
    dfree(d);
    return h;
}
double vec_norm2_host(const double* x, size_t n) { return sqrt(vec_dot_host(x, x, n)); }

} // namespace fc
