// translation unit: Chernoff spectral radius (mvd_chernoff.cuh)
#include "mvd_chernoff.cuh"
#include "mvd_launch.h"

cudaError_t mvd_launch_chernoff(const ChernoffParams& P, cudaStream_t st) {
    chernoff_rho_kernel<<<P.nu, CHERNOFF_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t mvd_launch_chernoff_dense(const ChernoffParams& P, cudaStream_t st) {
    chernoff_dense_kernel<<<P.nu, CHERNOFF_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}
