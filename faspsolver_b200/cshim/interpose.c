/* libfasp_cuda_setup.so — symbol interposition shim (plain C99).
 *
 * Defines FASP's own fasp_dcsr_trans (BlaSparseCSR.c:952) and fasp_blas_dcsr_rap (BlaSpmvCSR.c:999) and forwards
 * them to libfasp_cuda. Put in front of libfasp in the symbol search order (LD_PRELOAD, or linked before it), the
 * UNMODIFIED fasp_amg_setup_rs / fasp_amg_setup_sa (PreAMGSetupRS.c:212-214, PreAMGSetupSA.c:415-418) then run the
 * transpose R = P^T and the Galerkin product A_c = R A P of every level on the GPU; coarsening and interpolation
 * stay on the host. The device results are identical to the CPU's bit for bit, so the hierarchy is the same.
 * There is no CPU fallback: a failure is reported FASP's way (message + exit with the status, AuxMessage.c:213).
 */
#include <stdio.h>
#include <stdlib.h>
#include "fasp_cuda.h"

INT fasp_dcsr_trans(const dCSRmat* A, dCSRmat* AT)
{
    const INT st = fasp_cuda_dcsr_trans(A, AT);
    if (st < 0) {
        fprintf(stderr, "### ERROR: fasp_dcsr_trans on the device failed (%d): %s\n", (int)st, fasp_cuda_last_error());
        exit(st);
    }
    return st;
}

void fasp_blas_dcsr_rap(const dCSRmat* R, const dCSRmat* A, const dCSRmat* P, dCSRmat* RAP)
{
    const INT st = fasp_cuda_blas_dcsr_rap(R, A, P, RAP);
    if (st < 0) {
        fprintf(stderr, "### ERROR: fasp_blas_dcsr_rap on the device failed (%d): %s\n", (int)st, fasp_cuda_last_error());
        exit(st);
    }
}
