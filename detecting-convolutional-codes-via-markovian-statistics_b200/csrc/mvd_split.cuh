// mvd_split.cuh -- detection trials split along the time axis: few, long trials (the Pd-vs-N sweep of
// BASELINE config 3: N up to 10^5 at the reference's 10^4 iterations per point, divided over 8 GPUs).
//
// One thread per trial leaves the GPU empty when there are only a few thousand trials, and the chain of one
// trial is sequential.  But only two things in Pd_plotter.py:210-223 are order-dependent: the Markov-state
// trajectory (integers) and the two float64 sums of log_prob_sequence (:106-116).  So:
//   1. split_walk_kernel  -- one thread per (trial, chunk of SPLIT_CH steps): the position-addressed bit
//      source lets any chunk be generated on its own; the thread starts `warm` steps early from state 0 (the
//      relative-metric recursion forgets its start), records the state it has at the chunk start and
//      writes the edge index e_t = state * R + r_t of every step of its chunk;
//   2. split_fix_kernel   -- one thread per trial: a chunk whose speculated start differs from the previous
//      chunk's end is re-walked from the true state until both trajectories meet (exactly the
//      speculate / check / fix scheme of the learning chains, mvd_learn2.cuh); after it the edge list is the
//      sequential trajectory, bit for bit;
//   3. split_score_kernel -- one thread per trial adds log P1[e_t] and log Tref[e_t] in step order, so both
//      sums -- and the decision -- are the ones of the one-thread-per-trial kernels.
#pragma once
#include "mvd_detect2.cuh"
#include "mvd_learn2.cuh"

__device__ __forceinline__ uint32_t split_find(const unsigned long long* begin, uint32_t n, unsigned long long x) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (begin[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// EB = bytes per stored edge index (1: S R <= 256, 2: <= 65 536, 4 otherwise); a 16-byte group holds 16 / EB steps
template <int EB>
__device__ __forceinline__ void split_put(uint4& grp, uint32_t pos, uint32_t e) {
    const uint32_t sh = (pos * (uint32_t)EB * 8u) & 31u, w = (pos * (uint32_t)EB) >> 2;
    const uint32_t v = e << sh;
    if (w == 0) grp.x |= v; else if (w == 1) grp.y |= v; else if (w == 2) grp.z |= v; else grp.w |= v;
}

template <int EB>
__device__ __forceinline__ uint32_t split_get(const uint4& grp, uint32_t pos) {
    const uint32_t sh = (pos * (uint32_t)EB * 8u) & 31u, w = (pos * (uint32_t)EB) >> 2;
    const uint32_t word = w == 0 ? grp.x : (w == 1 ? grp.y : (w == 2 ? grp.z : grp.w));
    return EB == 4 ? word : ((word >> sh) & ((1u << (EB * 8)) - 1u));
}

template <bool SMEM, int EB>
__global__ void __launch_bounds__(SPLIT_BLOCK) split_walk_kernel(const __grid_constant__ Params P, const __grid_constant__ SplitParams SP) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t* nxt = P.nxt;
    if (SMEM) {
        uint32_t* s_nx = reinterpret_cast<uint32_t*>(smem_raw);
        for (uint32_t i = threadIdx.x; i < P.SR; i += SPLIT_BLOCK) s_nx[i] = P.nxt[i];
        __syncthreads();
        nxt = s_nx;
    }
    const unsigned long long g = (unsigned long long)blockIdx.x * SPLIT_BLOCK + threadIdx.x;
    if (g >= SP.nwork) return;
    const uint32_t seg = split_find(SP.work_begin, P.nsegs, g);
    const DevSeg sg = P.segs[seg];
    const uint32_t N = sg.N;
    const unsigned long long ntr = sg.trial_end - sg.trial_begin;
    const unsigned long long local = g - SP.work_begin[seg];
    const uint32_t c = (uint32_t)(local / ntr);                   // a warp = one chunk of 32 consecutive trials
    const unsigned long long tl = local % ntr;
    const unsigned long long trial = sg.trial_begin + tl;
    // edge of (trial tl, step t): 16-byte group (t / SPG) * ntr + tl, position t % SPG -- groups are trial-minor,
    // so that the stores here and the loads of the scoring pass are coalesced across trials
    constexpr uint32_t SPG = 16 / EB;
    uint4* E4 = reinterpret_cast<uint4*>(SP.edges + SP.edge_begin[seg]) + tl;
    const uint32_t t_begin = c * SPLIT_CH, t_end = min(N, t_begin + SPLIT_CH);
    const uint32_t w_begin = t_begin >= SP.warm ? t_begin - SP.warm : 0u;
    uint32_t sx = 0;                                              // state 0 (exact when w_begin == 0)
    uint32_t Rw[MVD_MAX_N];
    for (uint32_t b = w_begin >> 5; b * 32u < t_end; ++b) {
        const uint32_t t0 = b * 32u;
        const uint32_t valid = min(32u, N - t0);
        chain_block_words(P, sg, trial, b, valid, Rw);
        if (t0 == t_begin) SP.spec_start[g] = sx;
        const uint32_t nst = min(valid, t_end - t0);
        if (t0 >= t_begin) {
            // whole groups of the block are stored as one 16-byte word; a ragged tail group is stored as far as it goes
            for (uint32_t q = 0; q * SPG < nst; ++q) {
                uint4 grp = make_uint4(0u, 0u, 0u, 0u);
                const uint32_t cnt = min(SPG, nst - q * SPG);
#pragma unroll
                for (uint32_t u = 0; u < SPG; ++u) {
                    if (u < cnt) {
                        const uint32_t e = sx + word_of_step(Rw, P.n, q * SPG + u);
                        split_put<EB>(grp, u, e);
                        sx = SMEM ? nxt[e] : __ldg(nxt + e);
                    }
                }
                E4[(unsigned long long)(t0 / SPG + q) * ntr] = grp;
            }
        } else {
            for (uint32_t t = 0; t < nst; ++t) {
                const uint32_t e = sx + word_of_step(Rw, P.n, t);
                sx = SMEM ? nxt[e] : __ldg(nxt + e);
            }
        }
    }
    SP.end[g] = sx;
}

// The same walk for rate-1/2 codes, laid out like the one-trial-per-thread fast kernels (mvd_detect2.cuh): one info
// call per 128-step superblock, warp-voted mask-based flips (a warp = 32 trials of ONE chunk, so its control flow is
// uniform), bit-parallel encoder, received words interleaved once per 32 steps, 16 steps per straight-line stretch.
// grid (ceil(nch * ntr32 / SPLIT_BLOCK), segments), ntr32 = trials rounded up to a multiple of 32.
// Needs n = 2 and a warm-up that is a multiple of 128 steps; anything else takes split_walk_kernel.
template <bool SMEM, int EB>
__global__ void __launch_bounds__(SPLIT_BLOCK) split_walk2_kernel(const __grid_constant__ Params P, const __grid_constant__ SplitParams SP) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint4 tbm[8];
    constexpr uint32_t SPG = 16 / EB;
    const uint32_t seg = blockIdx.y;
    const DevSeg sg = P.segs[seg];
    const uint32_t N = sg.N, nch = (N + SPLIT_CH - 1u) / SPLIT_CH;
    const unsigned long long ntr = sg.trial_end - sg.trial_begin, ntr32 = (ntr + 31ull) & ~31ull;
    if ((unsigned long long)blockIdx.x * SPLIT_BLOCK >= nch * ntr32) return;      // uniform: shorter segment
    const uint32_t* nxt = P.nxt;
    if (SMEM) {
        uint32_t* s_nx = reinterpret_cast<uint32_t*>(smem_raw);
        for (uint32_t i = threadIdx.x; i < P.SR; i += SPLIT_BLOCK) s_nx[i] = P.nxt[i];
        nxt = s_nx;
    }
    if (threadIdx.x < 32u) reinterpret_cast<uint32_t*>(tbm)[threadIdx.x] = 0u - ((sg.threshold >> (31u - threadIdx.x)) & 1u);
    __syncthreads();
    const unsigned long long local = (unsigned long long)blockIdx.x * SPLIT_BLOCK + threadIdx.x;
    const uint32_t c = (uint32_t)(local / ntr32);
    if (c >= nch) return;                                          // whole warps (ntr32 is a multiple of 32)
    const unsigned long long tl = local % ntr32;
    const bool active = tl < ntr;
    const unsigned long long trial = sg.trial_begin + tl;
    const unsigned long long wid = SP.work_begin[seg] + (unsigned long long)c * ntr + tl;
    uint4* E4 = reinterpret_cast<uint4*>(SP.edges + SP.edge_begin[seg]) + tl;
    const uint32_t t_begin = c * SPLIT_CH, t_end = min(N, t_begin + SPLIT_CH);
    const uint32_t w_begin = t_begin >= SP.warm ? t_begin - SP.warm : 0u;
    const int ncalls = sg.dmin > 31u ? 0 : (int)((31u - sg.dmin) / 4u + 1u);
    const uint32_t c1 = (uint32_t)trial, c2 = (uint32_t)(trial >> 32), c3 = sg.stream;
    const uint32_t taps0 = sg.enc_taps[0], taps1 = sg.enc_taps[1];
    const int m = P.m;
    uint32_t sx = 0;                                               // state 0 (exact when w_begin == 0)
    const uint32_t sb0 = w_begin >> 7;
    uint32_t prevU = 0;
    if (sb0 > 0u && sg.random_input) prevU = philox10(((4u * (sb0 - 1u)) << 6) | 32u, c1, c2, c3, P).w;
    for (uint32_t sb = sb0; sb * 128u < t_end; ++sb) {
        uint4 Uw = philox10(((4u * sb) << 6) | 32u, c1, c2, c3, P);
        if (!sg.random_input) Uw = make_uint4(0, 0, 0, 0);
#pragma unroll 1
        for (int w = 0; w < 4; ++w) {
            const uint32_t t0 = sb * 128u + (uint32_t)w * 32u;
            if (t0 >= t_end) break;
            const uint32_t valid = min(32u, N - t0);
            const uint32_t vmask = valid == 32u ? 0xFFFFFFFFu : ((1u << valid) - 1u);
            const uint32_t U = pick(Uw, w);
            const uint32_t cb = (4u * sb + (uint32_t)w) << 6;
            const uint32_t e0 = lazy_bernoulli_s(cb, c1, c2, c3, tbm, ncalls, active ? vmask : 0u, P);
            const uint32_t e1 = lazy_bernoulli_s(cb | 8u, c1, c2, c3, tbm, ncalls, active ? vmask : 0u, P);
            uint32_t o0 = U & (0u - (taps0 & 1u)), o1 = U & (0u - (taps1 & 1u));
#pragma unroll 1
            for (int i = 1; i <= m; ++i) {
                const uint32_t sh = __funnelshift_l(prevU, U, i);
                o0 ^= sh & (0u - ((taps0 >> i) & 1u));
                o1 ^= sh & (0u - ((taps1 >> i) & 1u));
            }
            prevU = U;
            const uint32_t R0 = o0 ^ e0, R1 = o1 ^ e1;
            // received word of step t = bits (2t+1, 2t) of (whi:wlo); first output is the MSB
            const uint32_t wlo = (spread16(R0 & 0xFFFFu) << 1) | spread16(R1 & 0xFFFFu);
            const uint32_t whi = (spread16(R0 >> 16) << 1) | spread16(R1 >> 16);
            if (t0 == t_begin && active) SP.spec_start[wid] = sx;
            const bool emit = t0 >= t_begin && active;
            if (valid == 32u) {
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    const uint32_t x = h ? whi : wlo;
                    uint32_t e[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        e[j] = sx + ((x >> (2 * j)) & 3u);
                        sx = SMEM ? nxt[e[j]] : __ldg(nxt + e[j]);
                    }
                    if (emit) {
#pragma unroll
                        for (int q = 0; q < 16 / (int)SPG; ++q) {
                            uint4 grp = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                            for (int u = 0; u < (int)SPG; ++u) split_put<EB>(grp, (uint32_t)u, e[q * (int)SPG + u]);
                            E4[(unsigned long long)((t0 + 16u * (uint32_t)h) / SPG + (uint32_t)q) * ntr] = grp;
                        }
                    }
                }
            } else {
                const uint32_t nst = min(valid, t_end - t0);
                for (uint32_t q = 0; q * SPG < nst; ++q) {
                    uint4 grp = make_uint4(0u, 0u, 0u, 0u);
                    const uint32_t cnt = min(SPG, nst - q * SPG);
                    for (uint32_t u = 0; u < cnt; ++u) {
                        const uint32_t t = q * SPG + u;
                        const uint32_t r = ((t < 16u ? wlo : whi) >> (2u * (t & 15u))) & 3u;
                        const uint32_t e = sx + r;
#pragma unroll
                        for (uint32_t z = 0; z < SPG; ++z)
                            if (z == u) split_put<EB>(grp, z, e);
                        sx = SMEM ? nxt[e] : __ldg(nxt + e);
                    }
                    if (emit) E4[(unsigned long long)(t0 / SPG + q) * ntr] = grp;
                }
            }
        }
    }
    if (active) SP.end[wid] = sx;
}

// one thread per chain: repair the chunks whose speculated start state was wrong, in order
template <int EB>
__global__ void __launch_bounds__(SPLIT_BLOCK) split_fix_kernel(const __grid_constant__ Params P, const __grid_constant__ SplitParams SP) {
    constexpr uint32_t SPG = 16 / EB;
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= SP.nchains) return;
    // chain -> segment: out_offset = chains before the segment
    uint32_t lo = 0, hi = P.nsegs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (P.segs[mid].out_offset <= q) lo = mid; else hi = mid;
    }
    const uint32_t seg = lo;
    const DevSeg sg = P.segs[seg];
    const uint32_t N = sg.N, nch = (N + SPLIT_CH - 1u) / SPLIT_CH;
    const unsigned long long ntr = sg.trial_end - sg.trial_begin;
    const unsigned long long tl = q - sg.out_offset;
    const unsigned long long trial = sg.trial_begin + tl;
    const unsigned long long w0 = SP.work_begin[seg] + tl;        // work item of chunk c: w0 + c * ntr
    uint4* E4 = reinterpret_cast<uint4*>(SP.edges + SP.edge_begin[seg]) + tl;
    uint32_t fixed = 0;
    for (uint32_t c = 1; c < nch; ++c) {
        uint32_t st = SP.end[w0 + (unsigned long long)(c - 1u) * ntr], ss = SP.spec_start[w0 + (unsigned long long)c * ntr];
        if (st == ss) continue;
        ++fixed;
        const uint32_t t_begin = c * SPLIT_CH, t_end = min(N, t_begin + SPLIT_CH);
        uint32_t Rw[MVD_MAX_N];
        bool merged = false;
        for (uint32_t b = t_begin >> 5; b * 32u < t_end && !merged; ++b) {
            const uint32_t t0 = b * 32u;
            const uint32_t valid = min(32u, N - t0);
            chain_block_words(P, sg, trial, b, valid, Rw);
            const uint32_t nst = min(valid, t_end - t0);
            for (uint32_t t = 0; t < nst; ++t) {
                if (st == ss) {
                    merged = true;
                    break;
                }
                const uint32_t r = word_of_step(Rw, P.n, t);
                {
                    unsigned char* gb = reinterpret_cast<unsigned char*>(E4 + (unsigned long long)((t0 + t) / SPG) * ntr) + ((t0 + t) % SPG) * EB;
                    const uint32_t e = st + r;
                    if (EB == 1) *gb = (unsigned char)e;
                    else if (EB == 2) *reinterpret_cast<unsigned short*>(gb) = (unsigned short)e;
                    else *reinterpret_cast<uint32_t*>(gb) = e;
                }
                ss = __ldg(P.nxt + ss + r);
                st = __ldg(P.nxt + st + r);
            }
        }
        if (!merged && st != ss) SP.end[w0 + (unsigned long long)c * ntr] = st;   // the next chunk started from the wrong state too
    }
    if (fixed) atomicAdd(SP.ndirty, fixed);
}

// one thread per chain: the two sums of log_prob_sequence in step order, decision, tallies
// grid (chunks of blockDim.x chains, segments); blockDim.x <= SPLIT_BLOCK is chosen by the host so that a few
// thousand chains still spread over all SMs.
// SMEM: the {log P1, log Tref} rows of the segment's table are staged in shared memory, replicated
// 2^SP.ll_rep_shift times at 16-byte pitch (one copy per lane of a quarter warp, as in mvd_detect2.cuh) so that
// the random row reads of a warp are bank-conflict free; read with ld.shared (a generic pointer that may be
// shared or global costs a slower LD per step).
#define SPLIT_RING 32u                 // 16-byte groups of the edge stream in flight per scoring thread

template <bool SMEM>
__device__ __forceinline__ double2 split_ll(const double2* g, uint32_t sbase, uint32_t sh, uint32_t e) {
    if (SMEM) {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(sbase + (e << sh)));
        return v;
    }
    return __ldg(g + e);
}

template <bool SMEM, int EB>
__global__ void __launch_bounds__(SPLIT_BLOCK) split_score_kernel(const __grid_constant__ Params P, const __grid_constant__ SplitParams SP) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t SPG = 16 / EB;
    const uint32_t seg = blockIdx.y;
    const DevSeg sg = P.segs[seg];
    const unsigned long long ntr = sg.trial_end - sg.trial_begin;
    if ((unsigned long long)blockIdx.x * blockDim.x >= ntr) return;               // uniform: shorter segment
    const double2* ll = P.ll + (size_t)sg.table * P.SR;
    const uint32_t rs = (uint32_t)SP.ll_rep_shift, sh = rs + 4u;
    uint32_t sbase = 0;
    if (SMEM) {
        double2* s_ll = reinterpret_cast<double2*>(smem_raw);
        for (uint32_t i = threadIdx.x; i < (P.SR << rs); i += blockDim.x) s_ll[i] = ll[i >> rs];
        __syncthreads();
        sbase = (uint32_t)__cvta_generic_to_shared(smem_raw) + ((threadIdx.x & ((1u << rs) - 1u)) << 4);
    }
    const unsigned long long tl = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tl >= ntr) return;
    const uint32_t q = (uint32_t)(sg.out_offset + tl);
    const uint32_t N = sg.N;
    const uint4* E4 = reinterpret_cast<const uint4*>(SP.edges + SP.edge_begin[seg]) + tl;
    double a1 = 0.0, a0 = 0.0;
    // The adds are one dependent chain per sum (Pd_plotter.py:114-115 in step order): all that can overlap are the
    // loads.  With one warp per SM a load sees the whole DRAM latency (~2 000 cycles = 13 groups at 9.5 cycles per
    // step; measured: every group cost ~200 cycles of stall with a 12-deep register ring), so the edge stream goes
    // through an asynchronous shared-memory ring of SPLIT_RING groups per thread (cp.async), refilled as it is consumed.
    const uint32_t ngroups = N / SPG;
    const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(smem_raw) + SP.ring_offset + threadIdx.x * 16u;
    const uint32_t rstride = blockDim.x * 16u;                    // slot s of this thread: ring0 + s * rstride
#pragma unroll 1
    for (uint32_t i = 0; i < SPLIT_RING; ++i) {
        if (i < ngroups)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring0 + i * rstride), "l"(E4 + (unsigned long long)i * ntr));
        asm volatile("cp.async.commit_group;");
    }
    uint32_t slot = 0;
#pragma unroll 1
    for (uint32_t g = 0; g < ngroups; ++g) {
        asm volatile("cp.async.wait_group %0;" ::"n"(SPLIT_RING - 1));
        uint4 grp;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(grp.x), "=r"(grp.y), "=r"(grp.z), "=r"(grp.w) : "r"(ring0 + slot * rstride));
        double2 v[SPG];
#pragma unroll
        for (uint32_t u = 0; u < SPG; ++u) v[u] = split_ll<SMEM>(ll, sbase, sh, split_get<EB>(grp, u));
        if (g + SPLIT_RING < ngroups)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring0 + slot * rstride), "l"(E4 + (unsigned long long)(g + SPLIT_RING) * ntr));
        asm volatile("cp.async.commit_group;");
        slot = slot + 1u == SPLIT_RING ? 0u : slot + 1u;
#pragma unroll
        for (uint32_t u = 0; u < SPG; ++u) {
            a1 += v[u].x;
            a0 += v[u].y;
        }
    }
    if (ngroups * SPG < N) {
        const uint4 grp = __ldcs(E4 + (unsigned long long)ngroups * ntr);
        for (uint32_t u = 0; u < N - ngroups * SPG; ++u) {
            uint32_t e = 0;
#pragma unroll
            for (uint32_t w = 0; w < SPG; ++w)
                if (w == u) e = split_get<EB>(grp, w);
            const double2 v = split_ll<SMEM>(ll, sbase, sh, e);
            a1 += v.x;
            a0 += v.y;
        }
    }
    const bool win = sg.decide == 0 ? (a1 > a0) : (a1 <= a0);                    // Pd_plotter.py:215 / :222
    if (win) {
        atomicAdd(P.tallies + seg, 1ull);
        if (P.tallies2) atomicAdd(P.tallies2 + seg, 1ull);
    }
    if (P.logp) reinterpret_cast<double2*>(P.logp)[q] = make_double2(a1, a0);
}
