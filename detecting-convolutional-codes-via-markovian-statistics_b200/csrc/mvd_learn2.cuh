// mvd_learn2.cuh -- chunk-parallel learning chains (Pd_plotter.py:149-163) with exact results.
//
// The reference learns P1 from ONE sequential chain of max(5000, 200 S) steps per p (6 200 steps at
// S = 31, 87 000 at S = 435, 3 * 10^7 at S = 150 743).  Walked by one thread that is 0.1 us per step:
// at the reference's own trial counts the learning chain, not the trial loop, is the GPU time.
//
// Because the MVD-PHILOX-2 stream is addressed by position, any 32-step block of the chain can be
// generated independently.  The chain is cut into chunks of LP.chunk steps (128 .. 1024, chosen by the host):
//   1. learn_spec_kernel: thread c walks chunk c.  It starts LEARN_WARM steps early from Markov state 0
//      (the all-zero vector) -- the relative-metric recursion forgets its start as survivor paths
//      merge -- records the state it has at the chunk start (spec_start[c]), counts the transitions of
//      its chunk and records the state at the chunk end (end[c]).  Chunks whose warm-up reaches back
//      to step 0 start from the true state.
//   2. learn_check_kernel: chunk c is *clean* if spec_start[c] == end[c-1].  If every chunk is clean
//      the counts are exactly the sequential chain's (induction from chunk 0).
//   3. learn_fix_kernel (does nothing when no chunk is dirty): walks dirty chunks in order from their
//      true start state side by side with the speculated trajectory until the two meet (a
//      deterministic state machine driven by the same inputs stays merged), moving the counts of the
//      steps in between from the speculated edges to the true ones.  A fix that reaches the chunk end
//      unmerged updates end[c], which makes chunk c+1 dirty in turn.
// The result is bit-identical to the serial chain for any input (tests compare with the oracle).
#pragma once
#include "mvd_kernels.cuh"


// thread-local lazy Bernoulli word (no warp vote: chunks of a warp have different block counts)
__device__ __forceinline__ uint32_t lazy_bernoulli_t(uint32_t c0base, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t T,
                                                     int dmin, uint32_t vmask, const Params& P) {
    uint32_t und = vmask, e = 0;
    int d = 31;
    uint32_t k = 0;
    while (d >= dmin && und != 0u) {
        const uint4 w = philox10(c0base + k, c1, c2, c3, P);
        ++k;
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (d >= dmin) {
                if ((T >> d) & 1u) {
                    e |= und & ~ws[i];
                    und &= ws[i];
                } else {
                    und &= ~ws[i];
                }
                --d;
            }
        }
    }
    return e;
}


// received words of the 32-step block b of the chain of segment sg -> Rw[j] (bit t = received bit of output j)
__device__ __forceinline__ void chain_block_words(const Params& P, const DevSeg& sg, unsigned long long trial, uint32_t b,
                                                  uint32_t valid, uint32_t* Rw) {
    const int n = P.n, m = P.m;
    const uint32_t c1 = (uint32_t)trial, c2 = (uint32_t)(trial >> 32), c3 = sg.stream;
    const uint32_t vmask = valid >= 32u ? 0xFFFFFFFFu : ((1u << valid) - 1u);
    uint32_t U = 0, prevU = 0;
    if (sg.random_input) {
        const uint4 uw = philox10(((b & ~3u) << 6) | 32u, c1, c2, c3, P);
        U = pick(uw, (int)(b & 3u));
        if (b > 0u) {
            if (b & 3u) prevU = pick(uw, (int)((b - 1u) & 3u));
            else prevU = philox10((((b - 1u) & ~3u) << 6) | 32u, c1, c2, c3, P).w;
        }
    }
#pragma unroll
    for (int j = 0; j < MVD_MAX_N; ++j) {
        Rw[j] = 0;
        if (j < n) {
            const uint32_t E = lazy_bernoulli_t((b << 6) | (8u * (uint32_t)j), c1, c2, c3, sg.threshold, (int)sg.dmin, vmask, P);
            const uint32_t taps = sg.enc_taps[j];
            uint32_t o = (taps & 1u) ? U : 0u;
#pragma unroll
            for (int i = 1; i <= MVD_MAX_M; ++i)
                if (i <= m && ((taps >> i) & 1u)) o ^= __funnelshift_l(prevU, U, i);
            Rw[j] = o ^ E;
        }
    }
}

__device__ __forceinline__ void learn_block_words(const Params& P, const DevSeg& sg, uint32_t b, uint32_t valid,
                                                  uint32_t* Rw) {
    chain_block_words(P, sg, sg.trial_begin, b, valid, Rw);
}

__device__ __forceinline__ uint32_t word_of_step(const uint32_t* Rw, int n, uint32_t t) {
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < MVD_MAX_N; ++j)
        if (j < n) r = (r << 1) | ((Rw[j] >> t) & 1u);
    return r;
}

// grid (ceil(nchunks / LEARN_BLOCK), nsegs)
// (tables in global memory: the walk waits on dependent L2 reads -- long_scoreboard 17 per issued instruction at S = 25 751 --, so the
// register budget is capped for 12 blocks of 128 threads per SM instead of the 7 that 72 registers allow: 1.55 -> 1.43 ms for 7 chains
// of 5.2e6 steps.  Measured and rejected: per-thread run-length aggregation of the counts instead of the warp-level match: 1.95 ms)
template <bool SMEM>
__global__ void __launch_bounds__(LEARN_BLOCK, SMEM ? 1 : 12) learn_spec_kernel(const __grid_constant__ Params P,
                                                                 const __grid_constant__ LearnParams LP) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t seg = blockIdx.y;
    const DevSeg sg = P.segs[seg];
    const uint32_t L = sg.N;
    const uint32_t SR = P.SR;
    const uint32_t* nxt = P.nxt;
    uint32_t* hist = nullptr;
    if (SMEM) {
        uint32_t* s_nx = reinterpret_cast<uint32_t*>(smem_raw);
        hist = s_nx + SR;
        for (uint32_t i = threadIdx.x; i < SR; i += LEARN_BLOCK) {
            s_nx[i] = P.nxt[i];
            hist[i] = 0u;
        }
        __syncthreads();
        nxt = s_nx;
    }
    unsigned long long* counts = P.counts + (size_t)seg * SR;
    // counts in global memory: a per-block table of LEARN_HOT slots (edge -> count) takes the additions first and is flushed at the
    // end.  At low p a chain spends half its steps on 16 edges (S = 150 743, p = 0.001), and every block of the chain hammering
    // the same few L2 addresses made those chains 3.4 x slower than the ones at p >= 0.3, whose edges are all different
    // (one chain of 3.0e7 steps: 2.29 -> 0.45 ms at p = 0.001, 0.67 -> 0.75 ms at p = 0.5; seven chains 4.75 -> 2.2 ms).  Admitting
    // an edge only when two lanes of a warp meet on it measured no better at high p and worse at p = 0.1.
    __shared__ uint32_t hot_key[SMEM ? 1 : LEARN_HOT], hot_val[SMEM ? 1 : LEARN_HOT];
    if (!SMEM) {
        for (uint32_t i = threadIdx.x; i < LEARN_HOT; i += LEARN_BLOCK) {
            hot_key[i] = 0xFFFFFFFFu;
            hot_val[i] = 0u;
        }
        __syncthreads();
    }
    const uint32_t c = blockIdx.x * LEARN_BLOCK + threadIdx.x;
    // Every lane of a warp runs the same trip counts (warm / 32 + chunk / 32 blocks of 32 steps) with a per-lane
    // `live` predicate: chunks whose warm-up reaches back past step 0, the ragged last chunk and lanes beyond the last
    // chunk idle through the iterations they do not have, so the full-mask votes below are reached by all 32 lanes
    // together (the programming model's requirement for __ballot_sync / __match_any_sync).
    const uint32_t CH = LP.chunk;
    const bool have = c < (L + CH - 1u) / CH;
    const uint32_t t_begin = have ? c * CH : 0u;
    const uint32_t t_end = have ? min(L, t_begin + CH) : 0u;
    const uint32_t nblk = (LP.warm + CH) >> 5;              // warm is a multiple of 32 (mvd_set_option)
    const int b_first = (int)(t_begin >> 5) - (int)(LP.warm >> 5);
    uint32_t sx = 0;                                              // state 0 (exact when the walk starts at step 0)
    uint32_t Rw[MVD_MAX_N];
    for (uint32_t kb = 0; kb < nblk; ++kb) {
        const int bi = b_first + (int)kb;
        const uint32_t b = bi > 0 ? (uint32_t)bi : 0u;
        const uint32_t t0 = b * 32u;
        const bool live = have && bi >= 0 && t0 < t_end;
        uint32_t nst = 0;
        bool count = false;
        if (live) {
            const uint32_t valid = min(32u, L - t0);
            learn_block_words(P, sg, b, valid, Rw);
            if (t0 == t_begin) LP.spec_start[(size_t)seg * LP.nchunks + c] = sx;
            count = t0 >= t_begin;
            nst = min(valid, t_end - t0);
        }
        if (SMEM) {
            for (uint32_t t = 0; t < nst; ++t) {
                const uint32_t e = sx + word_of_step(Rw, P.n, t);
                if (count && t0 + t >= P.burn) atomicAdd(hist + e, 1u);
                sx = nxt[e];
            }
        } else {
            for (uint32_t t = 0; t < 32u; ++t) {
                const bool step = t < nst;
                const uint32_t e = step ? sx + word_of_step(Rw, P.n, t) : 0u;
                const bool tally = step && count && t0 + t >= P.burn;
                // the chunks of a warp belong to one chain, and at low p a chain sits on a handful of edges:
                // lanes that count the same edge elect one of them to add the group's size (one global atomic
                // per distinct edge and step instead of up to 32 serialised ones on the same address)
                const unsigned peers = __match_any_sync(0xFFFFFFFFu, tally ? e : 0xFFFFFFFFu);
                if (tally && (threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1)) {
                    const uint32_t h = (e * 0x9E3779B1u) >> (32 - LEARN_HOT_LOG2);
                    const uint32_t was = atomicCAS(hot_key + h, 0xFFFFFFFFu, e);          // claim the slot, or find it ours / taken
                    if (was == 0xFFFFFFFFu || was == e) atomicAdd(hot_val + h, (uint32_t)__popc(peers));
                    else atomicAdd(counts + e, (unsigned long long)__popc(peers));
                }
                if (step) sx = __ldg(nxt + e);
            }
        }
    }
    if (have) LP.end[(size_t)seg * LP.nchunks + c] = sx;
    if (!SMEM) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < LEARN_HOT; i += LEARN_BLOCK)
            if (hot_val[i]) atomicAdd(counts + hot_key[i], (unsigned long long)hot_val[i]);
    }
    if (SMEM) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < SR; i += LEARN_BLOCK) {
            const uint32_t v = hist[i];
            if (v) atomicAdd(counts + i, (unsigned long long)v);
        }
    }
}

// grid (ceil(nchunks / 256), nsegs): count chunks whose speculated start is not the previous chunk's end
__global__ void learn_check_kernel(const __grid_constant__ Params P, const __grid_constant__ LearnParams LP) {
    const uint32_t seg = blockIdx.y;
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nch = (P.segs[seg].N + LP.chunk - 1u) / LP.chunk;
    bool dirty = false;
    if (c >= 1u && c < nch) {
        const size_t o = (size_t)seg * LP.nchunks;
        dirty = LP.spec_start[o + c] != LP.end[o + c - 1u];
    }
    const int cnt = __syncthreads_count(dirty ? 1 : 0);
    if (threadIdx.x == 0 && cnt) atomicAdd(LP.ndirty + seg, (uint32_t)cnt);
}

// grid (nsegs), one 1024-thread block per segment; returns at once when the segment is clean
__global__ void __launch_bounds__(1024) learn_fix_kernel(const __grid_constant__ Params P,
                                                         const __grid_constant__ LearnParams LP) {
    const uint32_t seg = blockIdx.x;
    if (LP.ndirty[seg] == 0u) return;
    __shared__ uint32_t s_first;
    const DevSeg sg = P.segs[seg];
    const uint32_t L = sg.N, SR = P.SR;
    const uint32_t nch = (L + LP.chunk - 1u) / LP.chunk;
    const size_t o = (size_t)seg * LP.nchunks;
    volatile uint32_t* endv = LP.end + o;
    const uint32_t* spec = LP.spec_start + o;
    unsigned long long* counts = P.counts + (size_t)seg * SR;
    uint32_t c = 1;
    while (c < nch) {
        // next dirty chunk >= c
        if (threadIdx.x == 0) s_first = 0xFFFFFFFFu;
        __syncthreads();
        uint32_t found = 0xFFFFFFFFu;
        for (uint32_t base = c; base < nch; base += 1024u) {
            const uint32_t cc = base + threadIdx.x;
            const bool dirty = cc < nch && spec[cc] != endv[cc - 1u];
            if (dirty) atomicMin(&s_first, cc);
            __syncthreads();
            found = s_first;
            __syncthreads();
            if (found != 0xFFFFFFFFu) break;
        }
        __syncthreads();
        if (found == 0xFFFFFFFFu) break;
        if (threadIdx.x == 0) {
            // walk chunk `found` from its true start beside the speculated trajectory
            const uint32_t t_begin = found * LP.chunk, t_end = min(L, t_begin + LP.chunk);
            uint32_t st = endv[found - 1u], ss = spec[found];
            uint32_t Rw[MVD_MAX_N];
            bool merged = false;
            for (uint32_t b = t_begin >> 5; b * 32u < t_end && !merged; ++b) {
                const uint32_t t0 = b * 32u;
                const uint32_t valid = min(32u, L - t0);
                learn_block_words(P, sg, b, valid, Rw);
                const uint32_t nst = min(valid, t_end - t0);
                for (uint32_t t = 0; t < nst; ++t) {
                    if (st == ss) {
                        merged = true;
                        break;
                    }
                    const uint32_t r = word_of_step(Rw, P.n, t);
                    if (t0 + t >= P.burn) {
                        atomicAdd(counts + ss + r, ~0ull);            // -1
                        atomicAdd(counts + st + r, 1ull);
                    }
                    ss = __ldg(P.nxt + ss + r);
                    st = __ldg(P.nxt + st + r);
                }
            }
            if (!merged && st != ss) endv[found] = st;                // chunk found+1 becomes dirty
            __threadfence();
        }
        __syncthreads();
        c = found + 1u;
    }
}
