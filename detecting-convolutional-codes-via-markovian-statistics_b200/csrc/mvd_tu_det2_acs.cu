// translation unit: fast detection kernels, ACS engine, one trial per thread (mvd_detect2.cuh)
#include "mvd_detect2.cuh"
#include "mvd_launch.h"

namespace {
template <int LK, int M>
cudaError_t launch_det2(int lls, dim3 grid, unsigned threads, size_t smem, cudaStream_t st, const Params& P, const SegBatch& B) {
#define MVD_DET2_CASE(L)                                                                              \
    case L: {                                                                                         \
        auto kern = detect2_kernel<LK, M, L>;                                                         \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                               \
        kern<<<grid, threads, smem, st>>>(P, B);                                                      \
        return cudaGetLastError();                                                                    \
    }
    switch (lls) {
        MVD_DET2_CASE(4)
        MVD_DET2_CASE(5)
        MVD_DET2_CASE(6)
        MVD_DET2_CASE(7)
        default: return cudaErrorInvalidValue;
    }
#undef MVD_DET2_CASE
}
}  // namespace

cudaError_t mvd_launch_det2_acs(int lk, int m, int lls, bool gt, dim3 grid, unsigned threads, size_t smem, cudaStream_t st,
                                const Params& P, const SegBatch& B) {
    if (gt) {                                   // tables in global memory (large S): hash lookup, no replicas
        if (lk != LK_HASH || lls != 4) return cudaErrorInvalidValue;
        if (m == 3) detect2_kernel<LK_HASH, 3, 4, true><<<grid, threads, smem, st>>>(P, B);
        else if (m == 4) detect2_kernel<LK_HASH, 4, 4, true><<<grid, threads, smem, st>>>(P, B);
        else return cudaErrorInvalidValue;
        return cudaGetLastError();
    }
    if (lk == LK_DIRECT) return m == 1 ? launch_det2<LK_DIRECT, 1>(lls, grid, threads, smem, st, P, B)
                                       : launch_det2<LK_DIRECT, 2>(lls, grid, threads, smem, st, P, B);
    if (lk == LK_HASH) return m == 2 ? launch_det2<LK_HASH, 2>(lls, grid, threads, smem, st, P, B)
                                     : launch_det2<LK_HASH, 3>(lls, grid, threads, smem, st, P, B);
    return cudaErrorInvalidValue;
}
