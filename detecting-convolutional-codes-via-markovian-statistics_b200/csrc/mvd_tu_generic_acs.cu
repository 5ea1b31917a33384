// translation unit: generic ACS kernels (mvd_kernels.cuh), all modes and memories
#include "mvd_kernels.cuh"
#include "mvd_launch.h"

namespace {
template <int MODE, int NOUT>
cudaError_t launch_acs(int m, dim3 grid, size_t smem, cudaStream_t st, const Params& P) {
#define MVD_ACS_CASE(MM)                                                                              \
    case MM: {                                                                                        \
        auto kern = acs_kernel<MODE, NOUT, MM>;                                                       \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                               \
        kern<<<grid, MVD_BLOCK, smem, st>>>(P);                                                       \
        return cudaGetLastError();                                                                    \
    }
    switch (m) {
        MVD_ACS_CASE(1)
        MVD_ACS_CASE(2)
        MVD_ACS_CASE(3)
        MVD_ACS_CASE(4)
        MVD_ACS_CASE(5)
        MVD_ACS_CASE(6)
        default: return cudaErrorInvalidValue;
    }
#undef MVD_ACS_CASE
}
}  // namespace

cudaError_t mvd_launch_generic_acs(int mode, bool n2, int m, dim3 grid, size_t smem, cudaStream_t st, const Params& P) {
    if (mode == MODE_DETECT) return n2 ? launch_acs<MODE_DETECT, 2>(m, grid, smem, st, P) : launch_acs<MODE_DETECT, 0>(m, grid, smem, st, P);
    if (mode == MODE_LEARN) return launch_acs<MODE_LEARN, 0>(m, grid, smem, st, P);
    if (mode == MODE_TRACE) return launch_acs<MODE_TRACE, 0>(m, grid, smem, st, P);
    return n2 ? launch_acs<MODE_HASH, 2>(m, grid, smem, st, P) : launch_acs<MODE_HASH, 0>(m, grid, smem, st, P);
}
