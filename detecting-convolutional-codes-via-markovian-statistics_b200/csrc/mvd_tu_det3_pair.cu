// translation unit: ACS engine for m = 3, two trials per thread with a perfect-hash state lookup (mvd_detect3p.cuh)
#include "mvd_detect3p.cuh"
#include "mvd_launch.h"

cudaError_t mvd_launch_det3_pair(dim3 grid, unsigned threads, size_t smem, cudaStream_t st, const Params& P, const SegBatch& B) {
    auto kern = detect3p_kernel<0>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, threads, smem, st>>>(P, B);
    return cudaGetLastError();
}
