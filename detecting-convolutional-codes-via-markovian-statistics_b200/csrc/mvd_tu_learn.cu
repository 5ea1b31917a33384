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

// ---- detection trials split along the time axis (mvd_split.cuh)
#include "mvd_split.cuh"

cudaError_t mvd_launch_split_tables(const double2* ll, uint32_t SR, uint32_t ntables, uint32_t* tie, float2* apx, uint32_t* flags,
                                    unsigned long long* tiek, const SplitClasses& cls, cudaStream_t st) {
    const size_t cells = (size_t)SR * ntables;
    cudaError_t e = cudaMemsetAsync(flags, 0, 4, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(tiek, 0, 16 * (size_t)ntables, st);
    if (e != cudaSuccess) return e;
    split_tables_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(ll, cells, SR, tie, apx, flags, tiek, cls);
    return cudaGetLastError();
}

namespace {
template <class K>
cudaError_t with_smem(K kern, size_t bytes) {
    return bytes ? cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) : cudaSuccess;
}

template <int EB>
cudaError_t launch_split(size_t walk_bytes, size_t isum_bytes, size_t score_bytes, cudaStream_t st, const Params& P, const SplitParams& SP) {
    const unsigned wblocks = (unsigned)((SP.nwork + SPLIT_BLOCK - 1) / SPLIT_BLOCK);
    const unsigned cb = SP.chain_block;
    cudaError_t e;
    // 1. walk: one thread per (trial, chunk)
    if (SP.fast_walk) {                                  // n = 2, warm-up a multiple of 128: the fast walk
        const unsigned long long mx = ((SP.max_chunks * ((SP.max_trials + 31ull) & ~31ull)) + SPLIT_BLOCK - 1) / SPLIT_BLOCK;
        const dim3 wgrid((unsigned)mx, P.nsegs);
        if (SP.nxt_in_smem) {
            auto kern = split_walk2_kernel<true, EB>;
            if ((e = with_smem(kern, walk_bytes)) != cudaSuccess) return e;
            kern<<<wgrid, SPLIT_BLOCK, walk_bytes, st>>>(P, SP);
        } else {
            split_walk2_kernel<false, EB><<<wgrid, SPLIT_BLOCK, 0, st>>>(P, SP);
        }
    } else if (SP.nxt_in_smem) {
        auto kern = split_walk_kernel<true, EB>;
        if ((e = with_smem(kern, walk_bytes)) != cudaSuccess) return e;
        kern<<<wblocks, SPLIT_BLOCK, walk_bytes, st>>>(P, SP);
    } else {
        split_walk_kernel<false, EB><<<wblocks, SPLIT_BLOCK, 0, st>>>(P, SP);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // 2. repair + predictions: one warp per trial
    split_plan_kernel<EB><<<(SP.nchains + (SPLIT_BLOCK / 32) - 1) / (SPLIT_BLOCK / 32), SPLIT_BLOCK, 0, st>>>(P, SP);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // 3. re-associated partial sums: one thread per (trial, chunk)
    if (!SP.sequential) {
        const dim3 igrid((unsigned)((SP.max_chunks * SP.max_trials + SPLIT_IBLOCK - 1) / SPLIT_IBLOCK), P.nsegs);
        if (SP.ll_in_smem) {
            auto kern = SP.cls.n > 0 ? split_isum_kernel<true, EB, true> : split_isum_kernel<true, EB, false>;
            if ((e = with_smem(kern, isum_bytes)) != cudaSuccess) return e;
            kern<<<igrid, SPLIT_IBLOCK, isum_bytes, st>>>(P, SP);
        } else if (SP.cls.n > 0) {
            split_isum_kernel<false, EB, true><<<igrid, SPLIT_IBLOCK, 0, st>>>(P, SP);
        } else {
            split_isum_kernel<false, EB, false><<<igrid, SPLIT_IBLOCK, 0, st>>>(P, SP);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    // 4. the sums in order, decisions: one thread per trial
    const dim3 sgrid((unsigned)((SP.max_trials + cb - 1) / cb), P.nsegs);
    if (SP.ll_in_smem) {
        auto kern = split_score_kernel<true, EB>;
        if ((e = with_smem(kern, score_bytes)) != cudaSuccess) return e;
        kern<<<sgrid, cb, score_bytes, st>>>(P, SP);
    } else {
        auto kern = split_score_kernel<false, EB>;
        if ((e = with_smem(kern, score_bytes)) != cudaSuccess) return e;
        kern<<<sgrid, cb, score_bytes, st>>>(P, SP);
    }
    return cudaGetLastError();
}
}  // namespace

cudaError_t mvd_launch_split(size_t walk_bytes, size_t isum_bytes, size_t score_bytes, cudaStream_t st, const Params& P, const SplitParams& SP) {
    if (SP.edge_bytes == 1) return launch_split<1>(walk_bytes, isum_bytes, score_bytes, st, P, SP);
    if (SP.edge_bytes == 2) return launch_split<2>(walk_bytes, isum_bytes, score_bytes, st, P, SP);
    return launch_split<4>(walk_bytes, isum_bytes, score_bytes, st, P, SP);
}
