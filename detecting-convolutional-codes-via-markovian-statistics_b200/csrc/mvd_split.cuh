// mvd_split.cuh -- detection trials split along the time axis: few, long trials (the Pd-vs-N sweep of
// BASELINE config 3: N up to 10^5 at the reference's 10^4 iterations per point, divided over 8 GPUs).
//
// One thread per trial leaves the GPU empty when there are only a few thousand trials, and the chain of one
// trial is sequential.  Two things in Pd_plotter.py:210-223 are order-dependent: the Markov-state trajectory
// (integers) and the two float64 sums of log_prob_sequence (:106-116).  Both are cut into pieces here that are
// worked on in parallel and still give the reference's bits:
//   1. split_walk_kernel / split_walk2_kernel -- one thread per (trial, chunk of SP.chunk steps): the
//      position-addressed bit source lets any chunk be generated on its own; the thread starts `warm` steps early
//      from state 0 (the relative-metric recursion forgets its start), records the state it has at the chunk
//      start, writes the edge index e_t = state * R + r_t of every step of its chunk and a float32 estimate of the
//      two sums over every SPLIT_SUB steps;
//   2. split_plan_kernel -- one warp per trial: chunks whose speculated start differs from the previous chunk's
//      end are re-walked from the true state until both trajectories meet (the speculate / check / fix scheme of
//      the learning chains, mvd_learn2.cuh), after which the edge list is the sequential trajectory, bit for bit;
//      a prefix sum of the estimates predicts, for every sub-chunk, the binade [2^k, 2^(k+1)) in which each of the
//      two running sums will be while it crosses that sub-chunk;
//   3. split_isum_kernel -- one thread per (trial, chunk): for every sub-chunk with a prediction it runs the
//      float64 recurrence r <- r + v_t from r = -2^k instead of from the (unknown) running sum;
//   4. split_score_kernel -- one thread per trial goes through the sub-chunks in order: where the running sum
//      really is in the predicted binade and stays there, a + (r_end + 2^k) IS the value the step-by-step
//      recurrence reaches (see below); everywhere else it adds the sub-chunk's terms one by one.
//
// Why step 4 is exact.  All terms are log-probabilities, v_t <= 0 (checked on the device when the tables are
// installed; a table with a positive entry switches the predictions off), so |a| never decreases.  While |a| is in
// [2^k, 2^(k+1)) every sum a + v_t is rounded to a multiple of u = 2^(k-52), and because a itself is such a multiple
// the rounded result is a + rnd_u(v_t), rnd_u = nearest multiple of u -- unless v_t / u has fractional part exactly
// 1/2, where round-half-even looks at the parity of a (a "tie"; every v has exactly one binade in which that happens,
// k* = exponent(v) + 1 + trailing zeros of its mantissa, tabulated per edge).  So as long as the sum stays inside the
// binade and no tie term occurs, the recurrence adds the FIXED integers rnd_u(v_t) / u, whatever its starting point:
// started from -2^k it ends at -2^k + sum rnd_u(v_t), and that sum moves to the true starting value without any
// rounding.  The scoring thread checks exactly the two premises (exponent of a before, exponent of a + S after; a tie
// makes S a NaN), which also covers a recurrence from -2^k that left the binade: it ends at or beyond -2^(k+1), so
// |a + S| >= 2^(k+1) and the check fails.  Predictions only decide how often the cheap case applies.
#pragma once
#include "mvd_detect2.cuh"
#include "mvd_learn2.cuh"
#include <type_traits>

// SP.chunk = steps per chunk (a multiple of SPLIT_SUB, chosen by the host so that the walk fills the GPU)
#define SPLIT_PLAN_FAST (1u << 22)      // plan word of a sub-chunk: bits 0-10 / 11-21 biased exponents of the two sums

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

// ---- tables of the exact re-association, one entry per (log-likelihood table, edge); built once per mvd_set_loglik
// tie[e] = k1 | k0 << 8: the binade 2^k (k = 0 .. SPLIT_KMAX - 1, else 0xFF) in which log P1[e] / log Tref[e] is a tie term,
// apx[e] = the two terms in float32; flags bit 0: some term is positive or not a number (no predictions then)
#define SPLIT_KMAX 64u
__device__ __forceinline__ uint32_t split_tie_code(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v) & 0x7FFFFFFFFFFFFFFFull;
    if (b == 0ull || (b >> 52) == 0x7FFull) return 0xFFu;
    int ev = (int)(b >> 52) - 1023;
    unsigned long long mant = b & 0xFFFFFFFFFFFFFull;
    if ((b >> 52) == 0ull) ev = -1022; else mant |= 1ull << 52;
    const int k = ev + 1 + (__ffsll((long long)mant) - 1);
    return (k >= 0 && k < (int)SPLIT_KMAX) ? (uint32_t)k : 0xFFu;
}

// tiek[table] = {binades in which some log P1 term of the table is a tie, the same for log Tref}: a sub-chunk predicted into a
// binade without any reads no tie codes at all
// Class mode (cls.n > 0): log Tref takes at most three distinct non-zero values (rate-1/2 codes: T(1/2) entries are 1/4, 1/2, 3/4, 1),
// so the second sum of a sub-chunk is determined by how often each value occurs.  apx[e].y is then not the term but a float32
// COUNTER INCREMENT 256^j for class j (0 for a zero term): the walk's float additions count exactly (at most 128 terms per
// sub-chunk: every partial sum is an integer below 2^24), and split_plan_kernel turns the counts into the exact partial sum.
__global__ void split_tables_kernel(const double2* __restrict__ ll, size_t cells, uint32_t SR, uint32_t* __restrict__ tie,
                                    float2* __restrict__ apx, uint32_t* __restrict__ flags, unsigned long long* __restrict__ tiek,
                                    const SplitClasses cls) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    const double2 v = ll[i];
    const uint32_t t1 = split_tie_code(v.x), t0 = split_tie_code(v.y);
    tie[i] = t1 | (t0 << 8);
    if (cls.n > 0) {
        float inc = 0.f;
        if (v.y == cls.val[0]) inc = 1.f;
        else if (cls.n > 1 && v.y == cls.val[1]) inc = 256.f;
        else if (cls.n > 2 && v.y == cls.val[2]) inc = 65536.f;
        else if (v.y != 0.0) atomicOr(flags, 1u);                 // not one of the classes (cannot happen: the host built them from this table)
        apx[i] = make_float2((float)v.x, inc);
        if (t1 != 0xFFu) atomicOr(tiek + 2 * (i / SR), 1ull << t1);
        if (!(v.x <= 0.0) || !(v.y <= 0.0)) atomicOr(flags, 1u);
        return;
    }
    if (t1 != 0xFFu) atomicOr(tiek + 2 * (i / SR), 1ull << t1);
    if (t0 != 0xFFu) atomicOr(tiek + 2 * (i / SR) + 1, 1ull << t0);
    apx[i] = make_float2((float)v.x, (float)v.y);
    if (!(v.x <= 0.0) || !(v.y <= 0.0)) atomicOr(flags, 1u);
}

template <bool SMEM>
__device__ __forceinline__ float2 split_apx(const float2* g, uint32_t abase, uint32_t sh, uint32_t e) {
    if (SMEM) {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(abase + (e << sh)));
        return v;
    }
    return __ldg(g + e);
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
    // so that the stores here and the loads of the later passes are coalesced across trials
    constexpr uint32_t SPG = 16 / EB;
    uint4* E4 = reinterpret_cast<uint4*>(SP.edges + SP.edge_begin[seg]) + tl;
    float2* AX = SP.apx + SP.sub_begin[seg] + tl * ((N + SPLIT_SUB - 1u) / SPLIT_SUB);   // sub-chunk records are trial-major: AX[s]
    const float2* gapx = SP.apxtab + (size_t)sg.table * P.SR;
    const uint32_t t_begin = c * SP.chunk, t_end = min(N, t_begin + SP.chunk);
    const uint32_t w_begin = t_begin >= SP.warm ? t_begin - SP.warm : 0u;
    uint32_t sx = 0;                                              // state 0 (exact when w_begin == 0)
    float s1 = 0.f, s0 = 0.f;
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
                        const float2 f = __ldg(gapx + e);
                        s1 += f.x;
                        s0 += f.y;
                        sx = SMEM ? nxt[e] : __ldg(nxt + e);
                    }
                }
                E4[(unsigned long long)(t0 / SPG + q) * ntr] = grp;
            }
            if (((t0 + 32u) % SPLIT_SUB) == 0u || t0 + 32u >= t_end) {
                AX[t0 / SPLIT_SUB] = make_float2(s1, s0);
                s1 = s0 = 0.f;
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
// SMEM: NEXT and the float32 terms (replicated 2^SP.apx_rep_shift times at 8-byte pitch: one copy per lane of a half
// warp, LDS.64 is served per half warp) sit in shared memory.
template <bool SMEM, int EB>
__global__ void __launch_bounds__(SPLIT_BLOCK) split_walk2_kernel(const __grid_constant__ Params P, const __grid_constant__ SplitParams SP) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint4 tbm[8];
    constexpr uint32_t SPG = 16 / EB;
    const uint32_t seg = blockIdx.y, bx = blockIdx.x;
    const DevSeg sg = P.segs[seg];
    const uint32_t N = sg.N, nch = (N + SP.chunk - 1u) / SP.chunk;
    const unsigned long long ntr = sg.trial_end - sg.trial_begin, ntr32 = (ntr + 31ull) & ~31ull;
    if ((unsigned long long)bx * SPLIT_BLOCK >= nch * ntr32) return;              // uniform: shorter segment
    const uint32_t* nxt = P.nxt;
    const float2* gapx = SP.apxtab + (size_t)sg.table * P.SR;
    const uint32_t as = (uint32_t)SP.apx_rep_shift, ash = as + 3u;
    uint32_t abase = 0;
    if (SMEM) {
        uint32_t* s_nx = reinterpret_cast<uint32_t*>(smem_raw);
        for (uint32_t i = threadIdx.x; i < P.SR; i += SPLIT_BLOCK) s_nx[i] = P.nxt[i];
        nxt = s_nx;
        float2* s_apx = reinterpret_cast<float2*>(smem_raw + SP.walk_apx_offset);
        for (uint32_t i = threadIdx.x; i < (P.SR << as); i += SPLIT_BLOCK) s_apx[i] = gapx[i >> as];
        abase = (uint32_t)__cvta_generic_to_shared(s_apx) + ((threadIdx.x & ((1u << as) - 1u)) << 3);
    }
    if (threadIdx.x < 32u) reinterpret_cast<uint32_t*>(tbm)[threadIdx.x] = 0u - ((sg.threshold >> (31u - threadIdx.x)) & 1u);
    __syncthreads();
    const unsigned long long local = (unsigned long long)bx * SPLIT_BLOCK + threadIdx.x;
    const uint32_t c = (uint32_t)(local / ntr32);
    if (c >= nch) return;                                          // whole warps (ntr32 is a multiple of 32)
    const unsigned long long tl = local % ntr32;
    const bool active = tl < ntr;
    const unsigned long long trial = sg.trial_begin + tl;
    const unsigned long long wid = SP.work_begin[seg] + (unsigned long long)c * ntr + tl;
    uint4* E4 = reinterpret_cast<uint4*>(SP.edges + SP.edge_begin[seg]) + tl;
    float2* AX = SP.apx + SP.sub_begin[seg] + tl * ((N + SPLIT_SUB - 1u) / SPLIT_SUB);    // sub-chunk records are trial-major: AX[s]
    const uint32_t t_begin = c * SP.chunk, t_end = min(N, t_begin + SP.chunk);
    const uint32_t w_begin = t_begin >= SP.warm ? t_begin - SP.warm : 0u;
    const int ncalls = sg.dmin > 31u ? 0 : (int)((31u - sg.dmin) / 4u + 1u);
    const uint32_t c1 = (uint32_t)trial, c2 = (uint32_t)(trial >> 32), c3 = sg.stream;
    const uint32_t taps0 = sg.enc_taps[0], taps1 = sg.enc_taps[1];
    const int m = P.m;
    uint32_t sx = 0;                                               // state 0 (exact when w_begin == 0)
    float s1 = 0.f, s0 = 0.f;                                      // float32 estimate of the two sums over the current sub-chunk
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
            // received word of step t = (R0 bit t, R1 bit t), first output is the MSB: two bit-selects put the pairs of the
            // even steps into one word and those of the odd steps into another, at bits (t | 1, t & ~1) (see detect2p_kernel)
            const uint32_t wev = bitsel(R1, R0 << 1, 0x55555555u), wod = bitsel(R1 >> 1, R0, 0x55555555u);
            if (t0 == t_begin && active) SP.spec_start[wid] = sx;
            const bool emit = t0 >= t_begin && active;
            if (valid == 32u) {
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    const uint32_t xe = wev >> (16 * h), xo = wod >> (16 * h);
                    uint32_t e[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        e[j] = sx + ((((j & 1) ? xo : xe) >> (j & ~1)) & 3u);
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
                        float p1 = 0.f, p0 = 0.f;                   // a second pair of accumulators: two short chains instead of one
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const float2 fa = split_apx<SMEM>(gapx, abase, ash, e[j]), fb = split_apx<SMEM>(gapx, abase, ash, e[j + 1]);
                            s1 += fa.x;
                            s0 += fa.y;
                            p1 += fb.x;
                            p0 += fb.y;
                        }
                        s1 += p1;
                        s0 += p0;
                    }
                }
            } else {
                const uint32_t nst = min(valid, t_end - t0);
                for (uint32_t q = 0; q * SPG < nst; ++q) {
                    uint4 grp = make_uint4(0u, 0u, 0u, 0u);
                    const uint32_t cnt = min(SPG, nst - q * SPG);
                    for (uint32_t u = 0; u < cnt; ++u) {
                        const uint32_t t = q * SPG + u;
                        const uint32_t r = (((t & 1u) ? wod : wev) >> (t & ~1u)) & 3u;
                        const uint32_t e = sx + r;
#pragma unroll
                        for (uint32_t z = 0; z < SPG; ++z)
                            if (z == u) split_put<EB>(grp, z, e);
                        if (emit) {
                            const float2 f = split_apx<SMEM>(gapx, abase, ash, e);
                            s1 += f.x;
                            s0 += f.y;
                        }
                        sx = SMEM ? nxt[e] : __ldg(nxt + e);
                    }
                    if (emit) E4[(unsigned long long)(t0 / SPG + q) * ntr] = grp;
                }
            }
        }
        if (sb * 128u >= t_begin && active) {                       // SPLIT_SUB = 128 = one superblock
            AX[sb] = make_float2(s1, s0);
            s1 = s0 = 0.f;
        }
    }
    if (active) SP.end[wid] = sx;
}

// ---- one warp per trial: repair the chunks whose speculated start state was wrong, then predict the binades
__device__ __forceinline__ uint32_t split_bexp(double x) {          // biased exponent
    return ((uint32_t)__double2hiint(x) >> 20) & 0x7FFu;
}

template <int EB>
__global__ void __launch_bounds__(SPLIT_BLOCK) split_plan_kernel(const __grid_constant__ Params P, const __grid_constant__ SplitParams SP) {
    constexpr uint32_t SPG = 16 / EB;
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (q >= SP.nchains) return;                                   // whole warps
    // chain -> segment: out_offset = chains before the segment
    uint32_t lo = 0, hi = P.nsegs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (P.segs[mid].out_offset <= q) lo = mid; else hi = mid;
    }
    const uint32_t seg = lo;
    const DevSeg sg = P.segs[seg];
    const uint32_t N = sg.N, nch = (N + SP.chunk - 1u) / SP.chunk, nsub = (N + SPLIT_SUB - 1u) / SPLIT_SUB;
    const unsigned long long ntr = sg.trial_end - sg.trial_begin;
    const unsigned long long tl = q - sg.out_offset;
    const unsigned long long trial = sg.trial_begin + tl;
    const unsigned long long w0 = SP.work_begin[seg] + tl;        // work item of chunk c: w0 + c * ntr
    uint4* E4 = reinterpret_cast<uint4*>(SP.edges + SP.edge_begin[seg]) + tl;

    // -- 1. the lanes compare chunk starts with the ends before them; a dirty chunk is rare and repaired by lane 0
    uint32_t first_bad = 0xFFFFFFFFu;
    for (uint32_t c = 1u + lane; c < nch; c += 32u)
        if (first_bad == 0xFFFFFFFFu && __ldcg(SP.end + w0 + (unsigned long long)(c - 1u) * ntr) != __ldcg(SP.spec_start + w0 + (unsigned long long)c * ntr))
            first_bad = c;
    uint32_t c = __reduce_min_sync(0xFFFFFFFFu, first_bad);
    uint32_t fixed = 0;
    while (c != 0xFFFFFFFFu) {                                     // uniform
        uint32_t carry = 0;                                        // chunk c + 1 has to be looked at again
        if (lane == 0u) {
            uint32_t st = __ldcg(SP.end + w0 + (unsigned long long)(c - 1u) * ntr), ss = __ldcg(SP.spec_start + w0 + (unsigned long long)c * ntr);
            if (st != ss) {
                ++fixed;
                const uint32_t t_begin = c * SP.chunk, t_end = min(N, t_begin + SP.chunk);
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
                        unsigned char* gb = reinterpret_cast<unsigned char*>(E4 + (unsigned long long)((t0 + t) / SPG) * ntr) + ((t0 + t) % SPG) * EB;
                        const uint32_t e = st + r;
                        if (EB == 1) *gb = (unsigned char)e;
                        else if (EB == 2) *reinterpret_cast<unsigned short*>(gb) = (unsigned short)e;
                        else *reinterpret_cast<uint32_t*>(gb) = e;
                        if (SP.cls.n > 0) {
                            // class mode: the counts of the sub-chunk are exact data, not an estimate -- move this step's
                            const float2* gapx = SP.apxtab + (size_t)sg.table * P.SR;
                            float* cy = &(SP.apx + SP.sub_begin[seg] + tl * nsub + (t0 + t) / SPLIT_SUB)->y;
                            *cy += __ldg(gapx + e).y - __ldg(gapx + ss + r).y;
                        }
                        ss = __ldg(P.nxt + ss + r);
                        st = __ldg(P.nxt + st + r);
                    }
                }
                if (!merged && st != ss) {                         // the next chunk started from the wrong state too
                    SP.end[w0 + (unsigned long long)c * ntr] = st;
                    carry = 1u;
                }
            }
        }
        __syncwarp();                                              // lane 0's stores before the other lanes' loads
        carry = __shfl_sync(0xFFFFFFFFu, carry, 0);
        uint32_t nextc = 0xFFFFFFFFu;
        if (carry && c + 1u < nch) {
            nextc = c + 1u;
        } else {
            // next chunk after c that the first comparison found dirty (later ones of this lane are found again here)
            uint32_t mine = 0xFFFFFFFFu;
            for (uint32_t cc = 1u + lane; cc < nch; cc += 32u)
                if (cc > c && mine == 0xFFFFFFFFu && __ldcg(SP.end + w0 + (unsigned long long)(cc - 1u) * ntr) != __ldcg(SP.spec_start + w0 + (unsigned long long)cc * ntr))
                    mine = cc;
            nextc = __reduce_min_sync(0xFFFFFFFFu, mine);
        }
        c = nextc;
    }
    if (lane == 0u && fixed) atomicAdd(SP.ndirty, fixed);

    // -- 2. predicted binades: prefix sums of the float32 estimates (lanes = sub-chunks)
    // the records of a trial are consecutive in memory (trial-major): the lanes of this warp read and write whole sectors
    const float2* AX = SP.apx + SP.sub_begin[seg] + tl * nsub;
    uint32_t* PL = SP.plan + SP.sub_begin[seg] + tl * nsub;
    double2* RSP = SP.res + SP.sub_begin[seg] + tl * nsub;
    const bool predict = SP.sequential == 0 && (__ldcg(SP.flags) & 1u) == 0u;
    uint32_t ctie[3];                                              // class mode: the binade in which each value is a tie term
#pragma unroll
    for (int j = 0; j < 3; ++j) ctie[j] = j < SP.cls.n ? split_tie_code(SP.cls.val[j]) : 0xFFu;
    double c1 = 0.0, c0 = 0.0;                                     // magnitudes of the sums before the current 32 sub-chunks
    float2 fnext = make_float2(0.f, 0.f);                         // the records of the next 32 sub-chunks are read one round ahead
    if (lane < nsub) fnext = __ldcg(AX + lane);
    for (uint32_t base = 0; base < nsub; base += 32u) {
        const uint32_t s = base + lane;
        const float2 f = fnext;
        fnext = make_float2(0.f, 0.f);
        if (s + 32u < nsub) fnext = __ldcg(AX + s + 32u);
        const uint32_t cnt = SP.cls.n > 0 ? (uint32_t)f.y : 0u;      // class mode: 8-bit counts of the three values, exact
        double own1 = -(double)f.x, own0 = -(double)f.y;
        if (SP.cls.n > 0)
            own0 = -((double)(cnt & 0xFFu) * SP.cls.val[0] + (double)((cnt >> 8) & 0xFFu) * SP.cls.val[1] + (double)((cnt >> 16) & 0xFFu) * SP.cls.val[2]);
        double x1 = own1, x0 = own0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double y1 = __shfl_up_sync(0xFFFFFFFFu, x1, d), y0 = __shfl_up_sync(0xFFFFFFFFu, x0, d);
            if ((int)lane >= d) {
                x1 += y1;
                x0 += y0;
            }
        }
        const double e1 = c1 + x1, e0 = c0 + x0, b1 = e1 - own1, b0 = e0 - own0;   // sums after / before sub-chunk s
        uint32_t plan = 0u;
        if (predict && s < nsub) {
            // the estimates are good to ~1e-5: no prediction where either sum is that close to a power of two
            const uint32_t k1 = split_bexp(b1 * (1.0 - 1e-4)), k0 = split_bexp(b0 * (1.0 - 1e-4));
            const bool ok = k1 == split_bexp(e1 * (1.0 + 1e-4)) && k0 == split_bexp(e0 * (1.0 + 1e-4)) && k1 >= 1023u && k1 < 1023u + SPLIT_KMAX &&
                            k0 >= 1023u && k0 < 1023u + SPLIT_KMAX;
            if (ok) plan = SPLIT_PLAN_FAST | k1 | (k0 << 11);
            if (ok && SP.cls.n > 0) {
                // class mode: what the sub-chunk adds to the second sum inside binade 2^kk, exactly: every term equal to val_j adds
                // rnd_u(val_j) = (-2^kk + val_j) + 2^kk, a multiple of u = 2^(kk-52) below 2^52 u.  count x that in float64 is exact
                // unless it reaches 2^53 u = 2^(kk+1) -- and then the scoring thread's binade test fails whatever the rounding was.
                const uint32_t kk = k0 - 1023u;
                const double m0 = __hiloint2double((int)(0x80000000u | (k0 << 20)), 0);                       // -2^kk
                double S0 = 0.0;
                bool bad = false;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const uint32_t nj = (cnt >> (8 * j)) & 0xFFu;
                    if (j < SP.cls.n && nj) {
                        const double aq = (m0 + SP.cls.val[j]) - m0;                                          // <= 0
                        if (!(aq > m0) || ctie[j] == kk) bad = true;  // a term of 2^kk or more, or a round-half-even tie in this binade
                        S0 += (double)nj * aq;
                    }
                }
                if (bad) S0 = __longlong_as_double(0x7FF8000000000000ll);
                reinterpret_cast<double*>(RSP + s)[1] = S0;
            }
        }
        if (s < nsub) PL[s] = plan;
        c1 = __shfl_sync(0xFFFFFFFFu, e1, 31);
        c0 = __shfl_sync(0xFFFFFFFFu, e0, 31);
    }
}

// ---- log-likelihood rows in shared memory, replicated 2^rs times at 16-byte pitch (one copy per lane of a quarter warp: the
// random row reads of a warp are bank-conflict free), read with ld.shared (a generic pointer costs a slower LD per step)
template <bool SMEM>
__device__ __forceinline__ double2 split_ll(const double2* g, uint32_t sbase, uint32_t sh, uint32_t e) {
    if (SMEM) {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(sbase + (e << sh)));
        return v;
    }
    return __ldg(g + e);
}

template <bool SMEM>
__device__ __forceinline__ uint32_t split_tie(const uint32_t* g, uint32_t tbase, uint32_t sh, uint32_t e) {
    if (SMEM) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tbase + (e << sh)));
        return v;
    }
    return __ldg(g + e);
}

// ---- one thread per (trial, chunk): the recurrence of every predicted sub-chunk from -2^k
// grid (ceil(chunks * trials / SPLIT_IBLOCK), segments)
#define SPLIT_IBLOCK 256

// CLS: class mode -- the second sum comes from the walk's counts (split_plan_kernel); only log P1 is read here, 8 bytes per step
// from 16 copies at 8-byte pitch (LDS.64 is served per half warp) instead of 16 bytes from 8 copies.
template <bool SMEM, int EB, bool CLS = false>
__global__ void __launch_bounds__(SPLIT_IBLOCK) split_isum_kernel(const __grid_constant__ Params P, const __grid_constant__ SplitParams SP) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t SPG = 16 / EB;
    const uint32_t seg = blockIdx.y;
    const DevSeg sg = P.segs[seg];
    const uint32_t N = sg.N, nch = (N + SP.chunk - 1u) / SP.chunk;
    const unsigned long long ntr = sg.trial_end - sg.trial_begin;
    if ((unsigned long long)blockIdx.x * SPLIT_IBLOCK >= nch * ntr) return;       // uniform: shorter segment
    const double2* ll = P.ll + (size_t)sg.table * P.SR;
    const uint32_t* tie = SP.tietab + (size_t)sg.table * P.SR;
    // copies: log rows 2^rs (16-byte pitch; class mode 2^(rs + 1) at 8-byte pitch), tie rows one per lane
    const uint32_t rs = (uint32_t)SP.ll_rep_shift, vs = CLS ? (rs ? rs + 1u : 0u) : rs, sh = CLS ? vs + 3u : rs + 4u;
    const uint32_t ts = rs ? rs + 2u : 0u, tsh = ts + 2u;
    uint32_t sbase = 0, tbase = 0;
    if (SMEM) {
        uint32_t* s_tie = reinterpret_cast<uint32_t*>(smem_raw + SP.isum_tie_offset);
        if (CLS) {
            double* s_v1 = reinterpret_cast<double*>(smem_raw);
            for (uint32_t i = threadIdx.x; i < (P.SR << vs); i += SPLIT_IBLOCK) s_v1[i] = ll[i >> vs].x;
        } else {
            double2* s_ll = reinterpret_cast<double2*>(smem_raw);
            for (uint32_t i = threadIdx.x; i < (P.SR << rs); i += SPLIT_IBLOCK) s_ll[i] = ll[i >> rs];
        }
        for (uint32_t i = threadIdx.x; i < (P.SR << ts); i += SPLIT_IBLOCK) s_tie[i] = tie[i >> ts];
        __syncthreads();
        sbase = (uint32_t)__cvta_generic_to_shared(smem_raw) + ((threadIdx.x & ((1u << vs) - 1u)) << (CLS ? 3 : 4));
        tbase = (uint32_t)__cvta_generic_to_shared(s_tie) + ((threadIdx.x & ((1u << ts) - 1u)) << 2);
    }
    const unsigned long long local = (unsigned long long)blockIdx.x * SPLIT_IBLOCK + threadIdx.x;
    if (local >= nch * ntr) return;
    const uint32_t c = (uint32_t)(local / ntr);
    const unsigned long long tl = local % ntr;
    const uint4* E4 = reinterpret_cast<const uint4*>(SP.edges + SP.edge_begin[seg]) + tl;
    const unsigned long long rec0 = SP.sub_begin[seg] + tl * ((N + SPLIT_SUB - 1u) / SPLIT_SUB);   // trial-major records
    const uint32_t* PL = SP.plan + rec0;
    double2* RS = SP.res + rec0;
    const unsigned long long tk1 = __ldg(SP.tiek + 2 * (size_t)sg.table), tk0 = __ldg(SP.tiek + 2 * (size_t)sg.table + 1);
    const uint32_t spc = SP.chunk / SPLIT_SUB;
    // the plan words of the chunk's sub-chunks (consecutive in memory), all requested before the first is needed
    constexpr uint32_t SPC_MAX = SPLIT_CH_MAX / SPLIT_SUB;
    uint32_t plans[SPC_MAX];
#pragma unroll
    for (uint32_t j = 0; j < SPC_MAX; ++j) {
        plans[j] = 0u;
        if (j < spc && (c * spc + j) * SPLIT_SUB < N) plans[j] = __ldcg(PL + c * spc + j);
    }
#pragma unroll 1
    for (uint32_t j = 0; j < spc; ++j) {
        const uint32_t s = c * spc + j, t0 = s * SPLIT_SUB;
        if (t0 >= N) break;
        uint32_t plan = 0u;
#pragma unroll
        for (uint32_t z = 0; z < SPC_MAX; ++z)
            if (z == j) plan = plans[z];
        if (!(plan & SPLIT_PLAN_FAST)) continue;
        const uint32_t k1 = plan & 0x7FFu, k0 = (plan >> 11) & 0x7FFu;
        const double m1 = __hiloint2double((int)(0x80000000u | (k1 << 20)), 0), m0 = __hiloint2double((int)(0x80000000u | (k0 << 20)), 0);   // -2^k
        double r1 = m1, r0 = m0;
        bool tie1 = false, tie0 = false;
        const uint32_t kc1 = k1 - 1023u, kc0 = (k0 - 1023u) << 8;
        const uint32_t ns = min(SPLIT_SUB, N - t0), g0 = t0 / SPG, ng = ns / SPG, rem = ns - ng * SPG;
        const uint4* Eg = E4 + (unsigned long long)g0 * ntr;
        // one term: r <- r + v in both recurrences; CHK: is it a tie term of the sub-chunk's binade?
        auto term = [&](uint32_t e, auto chk) {
            if (CLS) {
                double v1;
                if (SMEM) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v1) : "r"(sbase + (e << sh)));
                else v1 = __ldg(&ll[e].x);
                r1 += v1;
            } else {
                const double2 v = split_ll<SMEM>(ll, sbase, sh, e);
                r1 += v.x;
                r0 += v.y;
            }
            if (decltype(chk)::value) {
                const uint32_t t = split_tie<SMEM>(tie, tbase, tsh, e);
                tie1 |= (t & 0xFFu) == kc1;
                if (!CLS) tie0 |= (t & 0xFF00u) == kc0;
            }
        };
        auto groups = [&](auto chk) {
            uint4 grp = make_uint4(0u, 0u, 0u, 0u);
            if (ng || rem) grp = __ldcg(Eg);
#pragma unroll 1
            for (uint32_t g = 0; g < ng; ++g) {
                const uint4 cur = grp;
                if (g + 1u < ng || rem) grp = __ldcg(Eg + (unsigned long long)(g + 1u) * ntr);   // (three ahead measured: no faster)
#pragma unroll
                for (uint32_t u = 0; u < SPG; ++u) term(split_get<EB>(cur, u), chk);
            }
            if (rem) {
                for (uint32_t u = 0; u < rem; ++u) {
                    uint32_t e = 0;
#pragma unroll
                    for (uint32_t w = 0; w < SPG; ++w)
                        if (w == u) e = split_get<EB>(grp, w);
                    term(e, chk);
                }
            }
        };
        if (((tk1 >> kc1) | (CLS ? 0ull : tk0 >> (k0 - 1023u))) & 1ull) groups(std::true_type{});   // the table has a tie term in one of the binades
        else groups(std::false_type{});
        // what the sub-chunk adds: r_end + 2^k (exact when the recurrence stayed in its binade); a tie term voids it
        double S1 = r1 - m1, S0 = r0 - m0;
        if (tie1) S1 = __longlong_as_double(0x7FF8000000000000ll);
        if (tie0) S0 = __longlong_as_double(0x7FF8000000000000ll);
        if (CLS) reinterpret_cast<double*>(RS + s)[0] = S1;    // the second sum is split_plan_kernel's
        else RS[s] = make_double2(S1, S0);
    }
}

// ---- one thread per trial: the two sums of log_prob_sequence (Pd_plotter.py:114-115), decision, tallies
// grid (chunks of blockDim.x chains, segments); blockDim.x <= SPLIT_BLOCK is chosen by the host so that a few
// thousand chains still spread over all SMs.
// The sub-chunk records {plan, partial sums} come through a per-thread shared-memory ring (cp.async, one batch of
// SPLIT_RB records ahead).  A thread runs through the records whose partial sums apply; at the first one that does not it
// waits for the other lanes of its warp to reach theirs, and the warp adds those sub-chunks term by term TOGETHER (one
// piece of code, one call site: lanes whose binade crossings fall into different sub-chunks of a batch still share the
// pass -- per batch the warp pays for the most crossings of any lane, not for the union over its lanes).
#define SPLIT_RB ((uint32_t)SPLIT_RB_HOST)   // records per batch

template <bool SMEM, int EB>
__global__ void __launch_bounds__(SPLIT_BLOCK) split_score_kernel(const __grid_constant__ Params P, const __grid_constant__ SplitParams SP) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t SPG = 16 / EB;
    constexpr uint32_t UR = SPG < 8u ? SPG : 8u;                   // rows per pipeline unit of the term-by-term pass
    constexpr uint32_t UPG = SPG / UR;                             // units per 16-byte group
    constexpr uint32_t GPS = SPLIT_SUB / SPG;                      // groups per full sub-chunk
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
    const bool active = tl < ntr;                                  // idle lanes of the last warp stay for the votes
    const uint32_t q = (uint32_t)(sg.out_offset + tl);
    const uint32_t N = sg.N, nsub = (N + SPLIT_SUB - 1u) / SPLIT_SUB, nbatch = (nsub + SPLIT_RB - 1u) / SPLIT_RB;
    const uint4* E4 = reinterpret_cast<const uint4*>(SP.edges + SP.edge_begin[seg]) + tl;
    const uint32_t* PL = SP.plan + SP.sub_begin[seg] + tl * nsub;     // trial-major records: a batch is 16 consecutive ones
    const double2* RS = SP.res + SP.sub_begin[seg] + tl * nsub;
    // ring: [2 halves][SPLIT_RB slots][blockDim.x threads] partial sums (16 B), then the plan words (4 B)
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem_raw) + SP.score_ring_offset;
    const uint32_t rres = ring + threadIdx.x * 16u, rplan = ring + 2u * SPLIT_RB * blockDim.x * 16u + threadIdx.x * 4u;
    auto fetch = [&](uint32_t b) {                                 // batch b -> half b & 1
        if (active && b < nbatch) {
            const uint32_t half = (b & 1u) * SPLIT_RB, n = min(SPLIT_RB, nsub - b * SPLIT_RB);
            const uint32_t* pl = PL + b * SPLIT_RB;
            const double2* rs = RS + b * SPLIT_RB;
            uint32_t dp = rplan + half * blockDim.x * 4u, dr = rres + half * blockDim.x * 16u;
#pragma unroll 4
            for (uint32_t j = 0; j < n; ++j) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dp), "l"(pl));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dr), "l"(rs));
                pl += 1;
                rs += 1;
                dp += blockDim.x * 4u;
                dr += blockDim.x * 16u;
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    double a1 = 0.0, a0 = 0.0;
    uint32_t nseq = 0;                                             // sub-chunks added term by term (performance counter)
    auto rows = [&](const uint4& g, uint32_t unit, double2 (&v)[UR]) {
#pragma unroll
        for (uint32_t u = 0; u < UR; ++u) v[u] = split_ll<SMEM>(ll, sbase, sh, split_get<EB>(g, unit * UR + u));
    };
    auto adds = [&](const double2 (&v)[UR]) {
#pragma unroll
        for (uint32_t u = 0; u < UR; ++u) {
            a1 += v[u].x;
            a0 += v[u].y;
        }
    };
    fetch(0);
#pragma unroll 1
    for (uint32_t b = 0; b < nbatch; ++b) {
        fetch(b + 1u);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        const uint32_t half = (b & 1u) * SPLIT_RB;
        const uint32_t cnt = active ? min(SPLIT_RB, nsub - b * SPLIT_RB) : 0u;
        uint32_t j = 0;
#pragma unroll 1
        while (true) {
            // the records whose partial sums apply: both sums in the predicted binade before and after.  Record j + 1 is read
            // while record j is tested (its slot always exists in the ring; what it holds past the batch is never used)
            uint32_t plan = 0u;
            double2 r = make_double2(0.0, 0.0);
            auto record = [&](uint32_t jj, uint32_t& pl, double2& rr) {
                const uint32_t slot = half + min(jj, SPLIT_RB - 1u);
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pl) : "r"(rplan + slot * blockDim.x * 4u));
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rr.x), "=d"(rr.y) : "r"(rres + slot * blockDim.x * 16u));
            };
            if (j < cnt) record(j, plan, r);
#pragma unroll 1
            while (j < cnt) {
                uint32_t plan_n;
                double2 r_n;
                record(j + 1u, plan_n, r_n);
                if (!(plan & SPLIT_PLAN_FAST)) break;
                // sign and exponent (the top 12 bits) of both sums, before and after, are those of -2^k
                const uint32_t K1 = 0x80000000u | (plan << 20), K0 = 0x80000000u | ((plan >> 11) << 20);
                const double n1 = a1 + r.x, n0 = a0 + r.y;
                const uint32_t bad = ((uint32_t)__double2hiint(a1) ^ K1) | ((uint32_t)__double2hiint(n1) ^ K1) |
                                     ((uint32_t)__double2hiint(a0) ^ K0) | ((uint32_t)__double2hiint(n0) ^ K0);
                if (bad >> 20) break;
                a1 = n1;
                a0 = n0;
                ++j;
                plan = plan_n;
                r = r_n;
            }
            const bool need = j < cnt;
            if (!__any_sync(0xFFFFFFFFu, need)) break;
            if (need) {
                // sub-chunk s term by term, in step order
                const uint32_t s = b * SPLIT_RB + j;
                ++j;
                ++nseq;
                const uint32_t t0 = s * SPLIT_SUB, ns = min(SPLIT_SUB, N - t0), g0 = t0 / SPG;
                if (ns == SPLIT_SUB) {
                    // groups in batches of 8 loads; the rows of unit x + 1 are read while the terms of unit x are added
#pragma unroll 1
                    for (uint32_t gb = 0; gb < GPS; gb += 8u) {
                        uint4 G[8];
#pragma unroll
                        for (uint32_t i = 0; i < 8u; ++i)
                            if (i < GPS) G[i] = __ldcg(E4 + (unsigned long long)(g0 + gb + i) * ntr);
                        constexpr uint32_t NU = (GPS < 8u ? GPS : 8u) * UPG;
                        double2 va[UR], vb[UR];
                        rows(G[0], 0u, va);
#pragma unroll
                        for (uint32_t x = 0; x < NU; x += 2u) {
                            if (x + 1u < NU) rows(G[(x + 1u) / UPG], (x + 1u) % UPG, vb);
                            adds(va);
                            if (x + 2u < NU) rows(G[(x + 2u) / UPG], (x + 2u) % UPG, va);
                            if (x + 1u < NU) adds(vb);
                        }
                    }
                } else {
                    // the ragged last sub-chunk of a trial
                    const uint32_t ngall = (ns + SPG - 1u) / SPG;
#pragma unroll 1
                    for (uint32_t g = 0; g < ngall; ++g) {
                        const uint4 grp = __ldcg(E4 + (unsigned long long)(g0 + g) * ntr);
                        const uint32_t c2 = min(SPG, ns - g * SPG);
#pragma unroll 1
                        for (uint32_t u = 0; u < c2; ++u) {
                            uint32_t e = 0;
#pragma unroll
                            for (uint32_t w = 0; w < SPG; ++w)
                                if (w == u) e = split_get<EB>(grp, w);
                            const double2 v = split_ll<SMEM>(ll, sbase, sh, e);
                            a1 += v.x;
                            a0 += v.y;
                        }
                    }
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (!active) return;
    const bool win = sg.decide == 0 ? (a1 > a0) : (a1 <= a0);                    // Pd_plotter.py:215 / :222
    if (win) {
        atomicAdd(P.tallies + seg, 1ull);
        if (P.tallies2) atomicAdd(P.tallies2 + seg, 1ull);
    }
    if (P.logp) reinterpret_cast<double2*>(P.logp)[q] = make_double2(a1, a0);
    if (nseq) atomicAdd(SP.nseq, (unsigned long long)nseq);
}
