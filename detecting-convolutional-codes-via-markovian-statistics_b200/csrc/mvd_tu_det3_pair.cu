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
        // log rows by slot in global memory (L2); bucket displacements in shared memory when smem says so (> 4 KB)
        const bool ds = smem > 4096;
        auto kern = ds ? (ph ? detect3p_kernel<4, true, 1, true> : detect3p_kernel<4, true, 0, true>)
                       : (ph ? detect3p_kernel<4, true, 1, false> : detect3p_kernel<4, true, 0, false>);
        if (ds) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, threads, smem, st>>>(P, B);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t mvd_launch_slot_rows(const double2* ll, const uint32_t* pht, uint32_t slots, uint32_t SR, uint32_t ntables, double2* out,
                                 cudaStream_t st) {
    const size_t n = (size_t)slots * 4u;
    slot_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ll, pht, slots, SR, ntables, out);
    return cudaGetLastError();
}

cudaError_t mvd_launch_pack_ll(const double* lp1, const double* ltref, const uint32_t* nxt, const uint32_t* tcode, uint32_t SR,
                               uint32_t ntables, double2* ll, uint4* gfsm1, cudaStream_t st) {
    const size_t n = (size_t)SR * ntables;
    pack_ll_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(lp1, ltref, nxt, tcode, SR, ntables, ll, gfsm1);
    return cudaGetLastError();
}
