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

namespace {
template <int EB>
cudaError_t launch_split(bool nxt_smem, size_t nxt_bytes, bool ll_smem, size_t ll_bytes, cudaStream_t st, const Params& P,
                         const SplitParams& SP) {
    const unsigned wblocks = (unsigned)((SP.nwork + SPLIT_BLOCK - 1) / SPLIT_BLOCK);
    const unsigned cb = SP.chain_block;
    const unsigned cblocks = (SP.nchains + cb - 1) / cb;
    cudaError_t e;
    if (SP.fast_walk) {                                  // n = 2, warm-up a multiple of 128: the fast walk
        unsigned long long mx = 0;
        mx = ((SP.max_chunks * ((SP.max_trials + 31ull) & ~31ull)) + SPLIT_BLOCK - 1) / SPLIT_BLOCK;
        const dim3 wgrid((unsigned)mx, P.nsegs);
        if (nxt_smem) {
            auto kern = split_walk2_kernel<true, EB>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nxt_bytes);
            if (e != cudaSuccess) return e;
            kern<<<wgrid, SPLIT_BLOCK, nxt_bytes, st>>>(P, SP);
        } else {
            split_walk2_kernel<false, EB><<<wgrid, SPLIT_BLOCK, 0, st>>>(P, SP);
        }
    } else if (nxt_smem) {
        auto kern = split_walk_kernel<true, EB>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nxt_bytes);
        if (e != cudaSuccess) return e;
        kern<<<wblocks, SPLIT_BLOCK, nxt_bytes, st>>>(P, SP);
    } else {
        split_walk_kernel<false, EB><<<wblocks, SPLIT_BLOCK, 0, st>>>(P, SP);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    split_fix_kernel<EB><<<cblocks, cb, 0, st>>>(P, SP);
    const dim3 sgrid((unsigned)((SP.max_trials + cb - 1) / cb), P.nsegs);
    const size_t ring_bytes = (size_t)cb * 16 * 32;      // SPLIT_RING groups of 16 bytes per thread
    const size_t sbytes = (size_t)SP.ring_offset + ring_bytes;
    if (ll_smem) {
        auto kern = split_score_kernel<true, EB>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbytes);
        if (e != cudaSuccess) return e;
        kern<<<sgrid, cb, sbytes, st>>>(P, SP);
    } else {
        auto kern = split_score_kernel<false, EB>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbytes);
        if (e != cudaSuccess) return e;
        kern<<<sgrid, cb, sbytes, st>>>(P, SP);
    }
    return cudaGetLastError();
}
}  // namespace

cudaError_t mvd_launch_split(bool nxt_smem, size_t nxt_bytes, bool ll_smem, size_t ll_bytes, cudaStream_t st, const Params& P,
                             const SplitParams& SP) {
    if (SP.edge_bytes == 1) return launch_split<1>(nxt_smem, nxt_bytes, ll_smem, ll_bytes, st, P, SP);
    if (SP.edge_bytes == 2) return launch_split<2>(nxt_smem, nxt_bytes, ll_smem, ll_bytes, st, P, SP);
    return launch_split<4>(nxt_smem, nxt_bytes, ll_smem, ll_bytes, st, P, SP);
}
