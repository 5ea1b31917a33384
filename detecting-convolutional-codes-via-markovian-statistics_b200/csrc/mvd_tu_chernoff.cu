// translation unit: Chernoff spectral radius (mvd_chernoff.cuh)
#include "mvd_chernoff.cuh"
#include "mvd_launch.h"

cudaError_t mvd_launch_chernoff(const ChernoffParams& P, cudaStream_t st) {
    if (P.wd) chernoff_rho_kernel<false><<<P.nu, CHERNOFF_BLOCK, 0, st>>>(P);
    else chernoff_rho_kernel<true><<<P.nu, CHERNOFF_BLOCK, 0, st>>>(P);        // no scratch: recompute the weights per product
    return cudaGetLastError();
}

cudaError_t mvd_launch_chernoff_dense(const ChernoffParams& P, cudaStream_t st) {
    chernoff_dense_kernel<<<P.nu, CHERNOFF_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}
