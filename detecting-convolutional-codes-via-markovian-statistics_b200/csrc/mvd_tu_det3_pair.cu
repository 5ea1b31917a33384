// translation unit: ACS engine for m = 3 / m = 4, two trials per thread with a perfect-hash state lookup (mvd_detect3p.cuh)
#include "mvd_detect3p.cuh"
#include "mvd_launch.h"

template <class K>
static cudaError_t launch_det3(K kern, dim3 grid, unsigned threads, size_t smem, cudaStream_t st, const Params& P, const SegBatch& B) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, threads, smem, st>>>(P, B);
    return cudaGetLastError();
}

// ds: (m = 4) bucket displacements in shared memory; anti: complement-label decoder (branch table of x_g, n - x_g bytes)
cudaError_t mvd_launch_det3_pair(int m, dim3 grid, unsigned threads, size_t smem, cudaStream_t st, const Params& P, const SegBatch& B,
                                 bool ds, bool anti) {
    const bool ph = P.src_mode == MVD_SRC_PHILOX;
#define DET3_PICK(M, GT, DS)                                                                                              \
    (anti ? (ph ? launch_det3(detect3p_kernel<M, GT, 1, DS, true>, grid, threads, smem, st, P, B)                          \
                : launch_det3(detect3p_kernel<M, GT, 0, DS, true>, grid, threads, smem, st, P, B))                         \
          : (ph ? launch_det3(detect3p_kernel<M, GT, 1, DS, false>, grid, threads, smem, st, P, B)                         \
                : launch_det3(detect3p_kernel<M, GT, 0, DS, false>, grid, threads, smem, st, P, B)))
    if (m == 3) return DET3_PICK(3, false, false);
    // m = 4: log rows by slot in global memory (L2)
    if (m == 4) return ds ? DET3_PICK(4, true, true) : DET3_PICK(4, true, false);
#undef DET3_PICK
    return cudaErrorInvalidValue;
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
