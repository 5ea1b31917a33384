// translation unit: parity-template baseline trials (mvd_parity.cuh)
#include "mvd_parity.cuh"
#include "mvd_launch.h"

cudaError_t mvd_launch_parity(dim3 grid, cudaStream_t st, const Params& P, const ParityBatch& B, uint32_t* satisfied) {
    parity_kernel<<<grid, PARITY_BLOCK, 0, st>>>(P, B, satisfied);
    return cudaGetLastError();
}
