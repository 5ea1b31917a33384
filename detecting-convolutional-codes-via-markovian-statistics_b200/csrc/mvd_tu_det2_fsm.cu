// translation unit: fast detection kernels -- NEXT-table engines and the two-trials-per-thread ACS kernel
#include "mvd_detect2.cuh"
#include "mvd_launch.h"

namespace {
template <int LK, int M, int LLS, bool GT, int NOUT = 2, bool BIG = false>
cudaError_t launch_one(dim3 grid, unsigned threads, size_t smem, cudaStream_t st, const Params& P, const SegBatch& B) {
    auto kern = detect2_kernel<LK, M, LLS, GT, NOUT, BIG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, threads, smem, st>>>(P, B);
    return cudaGetLastError();
}
}  // namespace

cudaError_t mvd_launch_det2_fsm(int lk, int lls, bool gt, dim3 grid, unsigned threads, size_t smem, cudaStream_t st,
                                const Params& P, const SegBatch& B) {
    if (P.n == 3) {                                      // rate-1/3: the same engines behind the n = 3 driver
        if (lk == LK_FSM1) {
            if (gt) return launch_one<LK_FSM1, 1, 4, true, 3>(grid, threads, smem, st, P, B);
            switch (lls) {
                case 4: return launch_one<LK_FSM1, 1, 4, false, 3>(grid, threads, smem, st, P, B);
                case 5: return launch_one<LK_FSM1, 1, 5, false, 3>(grid, threads, smem, st, P, B);
                case 6: return launch_one<LK_FSM1, 1, 6, false, 3>(grid, threads, smem, st, P, B);
                default: return launch_one<LK_FSM1, 1, 7, false, 3>(grid, threads, smem, st, P, B);
            }
        }
        if (lk != LK_FSM || gt) return cudaErrorInvalidValue;
        switch (lls) {
            case 4: return launch_one<LK_FSM, 1, 4, false, 3>(grid, threads, smem, st, P, B);
            case 5: return launch_one<LK_FSM, 1, 5, false, 3>(grid, threads, smem, st, P, B);
            case 6: return launch_one<LK_FSM, 1, 6, false, 3>(grid, threads, smem, st, P, B);
            case 7: return launch_one<LK_FSM, 1, 7, false, 3>(grid, threads, smem, st, P, B);
            default: return cudaErrorInvalidValue;
        }
    }
    if (lk == LK_FSM1) {
        if (gt) return launch_one<LK_FSM1, 1, 4, true>(grid, threads, smem, st, P, B);
        if (threads > DET2_BLOCK)                        // two blocks of DET2_BIG_BLOCK threads per SM (plan_det2)
            switch (lls) {
                case 5: return launch_one<LK_FSM1, 1, 5, false, 2, true>(grid, threads, smem, st, P, B);
                case 6: return launch_one<LK_FSM1, 1, 6, false, 2, true>(grid, threads, smem, st, P, B);
                case 7: return launch_one<LK_FSM1, 1, 7, false, 2, true>(grid, threads, smem, st, P, B);
                default: return cudaErrorInvalidValue;
            }
        switch (lls) {
            case 4: return launch_one<LK_FSM1, 1, 4, false>(grid, threads, smem, st, P, B);
            case 5: return launch_one<LK_FSM1, 1, 5, false>(grid, threads, smem, st, P, B);
            case 6: return launch_one<LK_FSM1, 1, 6, false>(grid, threads, smem, st, P, B);
            default: return launch_one<LK_FSM1, 1, 7, false>(grid, threads, smem, st, P, B);
        }
    }
    if (lk != LK_FSM || gt) return cudaErrorInvalidValue;
    switch (lls) {
        case 4: return launch_one<LK_FSM, 1, 4, false>(grid, threads, smem, st, P, B);
        case 5: return launch_one<LK_FSM, 1, 5, false>(grid, threads, smem, st, P, B);
        case 6: return launch_one<LK_FSM, 1, 6, false>(grid, threads, smem, st, P, B);
        case 7: return launch_one<LK_FSM, 1, 7, false>(grid, threads, smem, st, P, B);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t mvd_launch_det2_pair(dim3 grid, unsigned threads, size_t smem, cudaStream_t st, const Params& P, const SegBatch& B) {
    const bool ph = P.src_mode == MVD_SRC_PHILOX;
    auto kern = P.bm_antipodal ? (ph ? detect2p_kernel<1, 1> : detect2p_kernel<0, 1>) : (ph ? detect2p_kernel<1, 0> : detect2p_kernel<0, 0>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, threads, smem, st>>>(P, B);
    return cudaGetLastError();
}

cudaError_t mvd_launch_det2(int lk, int m, int lls, bool gt, bool pair, dim3 grid, unsigned threads, size_t smem,
                            cudaStream_t st, const Params& P, const SegBatch& B) {
    if (pair) return mvd_launch_det2_pair(grid, threads, smem, st, P, B);
    if (lk == LK_FSM || lk == LK_FSM1) return mvd_launch_det2_fsm(lk, lls, gt, grid, threads, smem, st, P, B);
    return mvd_launch_det2_acs(lk, m, lls, gt, grid, threads, smem, st, P, B);
}
