// translation unit: chunk-parallel learning chains (mvd_learn2.cuh)
#include "mvd_learn2.cuh"
#include "mvd_launch.h"

cudaError_t mvd_launch_learn(bool smem_tables, size_t lsmem, uint32_t nsegs, cudaStream_t st, const Params& P,
                             const LearnParams& LP) {
    const dim3 g1((LP.nchunks + LEARN_BLOCK - 1) / LEARN_BLOCK, nsegs);
    if (smem_tables) {
        auto kern = learn_spec_kernel<true>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem);
        if (e != cudaSuccess) return e;
        kern<<<g1, LEARN_BLOCK, lsmem, st>>>(P, LP);
    } else {
        learn_spec_kernel<false><<<g1, LEARN_BLOCK, 0, st>>>(P, LP);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    learn_check_kernel<<<dim3((LP.nchunks + 255) / 256, nsegs), 256, 0, st>>>(P, LP);
    learn_fix_kernel<<<nsegs, 1024, 0, st>>>(P, LP);
    return cudaGetLastError();
}
