// bsr.cuh — device block-CSR matrix (nb x nb row-major blocks, fasp_block.h:34-66) and its kernels.
#pragma once
#include "common.cuh"

namespace fc {

struct DevBSR {
    int       ROW = 0, COL = 0, nb = 1;
    long long NNZ = 0;
    int*      ia  = nullptr;     // ROW+1 (+pad)
    int*      ja  = nullptr;     // NNZ (+pad)
    double*   val = nullptr;     // NNZ * nb*nb (+pad); nullptr when every block is the identity (`ident`)
    bool      ident = false;     // UA-AMG transfer operators: identity blocks, only ia/ja are kept (4 B per block)
    int2*     blkdesc = nullptr; // {first block row, ia[first block row]} per row-block boundary
    int       nblk    = 0;
    int       blk_cap = 0;       // blocks per stage
    int       blk_rb  = 32;      // block rows per row block = per CTA of the pipelined kernel
    size_t    bytes   = 0;
};

enum BsrMode {
    BSR_MXV    = 0,   // y = A x                                  fasp_blas_dbsr_mxv   BlaSpmvBSR.c:1055
    BSR_AXPY   = 1,   // y = alpha ((1/alpha) y + A x)            fasp_blas_dbsr_aAxpy BlaSpmvBSR.c:514
    BSR_RESID  = 2,   // y = b - A x  (copy + aAxpy(-1), PreMGCycle.c:394-395)
    BSR_JACOBI = 3,   // y_I = Dinv_I (b_I - sum_{J != I} A_IJ x_J)  fasp_smoother_dbsr_jacobi1 ItrSmootherBSR.c:263
    BSR_DINV   = 4    // y_I = Dinv_I b_I : the same sweep from x = 0 (b - A*0 == b exactly), no pass over A
};

struct BsrArgs {
    int           mode  = BSR_MXV;
    double        alpha = 1.0;
    const double* x     = nullptr;
    const double* b     = nullptr;
    double*       y     = nullptr;
    const double* diaginv = nullptr;   // ROW * nb*nb inverted diagonal blocks (Jacobi)
    Reduce        red;
    const int*    done = nullptr;
    bool          conditional = false;
};

// detect_identity: drop the values when every block is exactly the identity (UA-AMG P and R = P^T,
// PreAMGAggregationBSR.inl:185-189): the kernels then gather / sum whole block rows of x.
void bsr_upload(DevBSR& d, int ROW, int COL, long long NNZ, int nb, const int* ia, const int* ja,
                const double* val, bool detect_identity = false);
void bsr_free(DevBSR& d);
void bsr_launch(const DevBSR& A, const BsrArgs& a);

// algorithmic bytes of one pass: (8 nb^2 + 4) NNZ + 4 (ROW+1) + 8 nb (COL + ROW)
inline double bsr_spmv_bytes(const DevBSR& d, bool read_y)
{
    return ((d.ident ? 0.0 : 8.0 * d.nb * d.nb) + 4.0) * (double)d.NNZ + 4.0 * (d.ROW + 1) +
           8.0 * d.nb * ((double)d.COL + d.ROW) + (read_y ? 8.0 * d.nb * d.ROW : 0.0);
}

// expand to a dense (ROW nb) x (COL nb) row-major matrix on the host side of the device (used for
// the coarsest-level inverse)
void bsr_to_dense(const DevBSR& A, double* dense_dev);

} // namespace fc
