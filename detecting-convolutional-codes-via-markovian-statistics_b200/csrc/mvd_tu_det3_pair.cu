// translation unit: ACS engine for m = 3 / m = 4, two trials per thread with a perfect-hash state lookup (mvd_detect3p.cuh)
#include "mvd_detect3p.cuh"
#include "mvd_launch.h"

cudaError_t mvd_launch_det3_pair(int m, dim3 grid, unsigned threads, size_t smem, cudaStream_t st, const Params& P, const SegBatch& B) {
    const bool ph = P.src_mode == MVD_SRC_PHILOX;
    if (m == 3) {
        auto kern = ph ? detect3p_kernel<3, false, 1> : detect3p_kernel<3, false, 0>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, threads, smem, st>>>(P, B);
    } else if (m == 4) {
        if (ph) detect3p_kernel<4, true, 1><<<grid, threads, smem, st>>>(P, B);      // tables stay in global memory (L2)
        else detect3p_kernel<4, true, 0><<<grid, threads, smem, st>>>(P, B);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
