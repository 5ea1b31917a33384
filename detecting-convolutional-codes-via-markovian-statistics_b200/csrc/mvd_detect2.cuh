// mvd_detect2.cuh -- the throughput kernels of the detection trial loop (Pd_plotter.py:210-223)
// for rate-1/n codes with n = 2 whose tables fit shared memory (every code of BASELINE configs 1-3
// and 5).  Same arithmetic and the same per-trial results as the generic kernels in
// mvd_kernels.cuh; what changes is how the work is laid out for the SM:
//
//   * every table access is an LDS on a *bank-conflict-free replica*: the {log P1, log Tref} pair
//     of an edge is replicated 2^(LLS-4) times (one copy per lane of a quarter warp, the unit an
//     LDS.128 is served in), the metric-vector -> Markov-state table (ACS) / the NEXT table (FSM)
//     32 times (one copy per lane = per bank).  A warp's step costs 4 + 4 + 1 shared-memory
//     wavefronts (ACS: log pair, branch metrics, state) instead of ~21 with plain tables;
//   * the received word of a step is never materialised as an index: the two received-bit words
//     of 32 steps are bit-interleaved once (r_t = bits 2t+1, 2t) and a step takes
//     ((w >> (2j - LLS)) & (3 << LLS)) = r * stride directly as the byte offset of both the
//     branch-metric row and the log-likelihood row;
//   * m <= 2: the normalised metric vector packs into 8 bits, so the state_index dict of
//     Pd_plotter.py:139 is a direct 256-entry table; two IMADs turn the two 16x2 metric registers
//     into the byte address of this lane's copy (see direct_addr);
//   * the step loop is unrolled 8-fold, not 32-fold: the whole kernel is ~1/8 of the instruction
//     footprint of the generic kernel, which was instruction-cache bound.
//
// The kernels assume a *closed* state table (every successor of every state is in the table --
// verified on the host in install_states); anything else takes the generic, checked path.
#pragma once
#include "mvd_kernels.cuh"

#include <type_traits>


__device__ __forceinline__ uint32_t spread16(uint32_t x) {      // bit i -> bit 2i  (x < 2^16)
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// r * 2^LLS for step J of a 16-bit (8-step) chunk of the interleaved received word
template <int LLS, int J>
__device__ __forceinline__ uint32_t roff(uint32_t w) {
    constexpr int sh = 2 * J - LLS;
    return (sh >= 0 ? (w >> (sh >= 0 ? sh : 0)) : (w << (sh < 0 ? -sh : 0))) & (3u << LLS);
}

// 32 Bernoulli(T / 2^32) lanes (MVD-PHILOX-2, same words as lazy_bernoulli) for a block-uniform
// threshold.  tbm[k] (shared memory, staged once per block) is all-ones if bit 31-k of T is set: the
// bits of T steer LOP3 masks instead of branches.  Levels below ctz(T) inside the last call have
// T-bit 0 and can only retire undecided lanes, never set a flip, so running all four levels of a
// call gives the same word as the level-exact loop of lazy_bernoulli.
__device__ __forceinline__ uint32_t lazy_bernoulli_s(uint32_t c0base, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     const uint4* tbm, int ncalls, uint32_t vmask, const Params& P) {
    uint32_t und = vmask, e = 0;
    for (int k = 0; k < ncalls; ++k) {
        if (!__any_sync(0xFFFFFFFFu, und != 0u)) break;
        const uint4 w = philox10(c0base + (uint32_t)k, c1, c2, c3, P);
        const uint4 tb = tbm[k];
        e |= und & ~w.x & tb.x;
        und &= ~(w.x ^ tb.x);
        e |= und & ~w.y & tb.y;
        und &= ~(w.y ^ tb.y);
        e |= und & ~w.z & tb.z;
        und &= ~(w.z ^ tb.z);
        e |= und & ~w.w & tb.w;
        und &= ~(w.w ^ tb.w);
    }
    return e;
}

// absolute shared-window loads (ld.shared: no generic-address conversion)
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds_d2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// lazy_bernoulli_s with the threshold masks addressed in the shared window (ld.shared, no generic pointer)
__device__ __forceinline__ uint32_t lazy_bernoulli_a(uint32_t c0base, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     uint32_t a_tbm, int ncalls, uint32_t vmask, const Params& P) {
    uint32_t und = vmask, e = 0;
    for (int k = 0; k < ncalls; ++k) {
        if (!__any_sync(0xFFFFFFFFu, und != 0u)) break;
        const uint4 w = philox10(c0base + (uint32_t)k, c1, c2, c3, P);
        const uint4 tb = lds_v4(a_tbm + 16u * (uint32_t)k);
        e |= und & ~w.x & tb.x;
        und &= ~(w.x ^ tb.x);
        e |= und & ~w.y & tb.y;
        und &= ~(w.y ^ tb.y);
        e |= und & ~w.z & tb.z;
        und &= ~(w.z ^ tb.z);
        e |= und & ~w.w & tb.w;
        und &= ~(w.w ^ tb.w);
    }
    return e;
}

// ------------------------------------------------------------------------------------------ flip words of a 32-step block
// The four flip words of a 32-step block (2 trials x 2 outputs), MVD-PHILOX-2 bits.  With the warp-wide vote of
// lazy_bernoulli_a a third call of a word is made by all 32 threads whenever ONE of the warp's 1 024 lanes is still
// undecided after 8 levels (98 % of the time) although it serves ~4 of them, a fourth 22 % of the time: 3.2 calls
// per word, 60 % more than any single word needs.  Here calls 0 and 1 of every word (levels 31..24) are made by
// everybody without a vote -- eight independent calls, two of them in flight at a time -- and what is still undecided
// afterwards (2^-8 of the lanes, ~15 of a warp's 128 words) is queued in shared memory as (owner lane | word << 5,
// undecided mask) items.  The lanes of the warp then pick up ONE item each, make the owner's later calls for it (the
// calls are addressed by position: any thread can make them) and write the flips back into the item, where the
// owner collects them: one Philox call per item instead of one per thread and word.  Same words as lazy_bernoulli_a,
// bit for bit.
//
// Four levels of the comparison uniform < T from one call, least significant level first: lt = "these four bits of
// the uniform are below T's" needs one LOP3 per level ((~w & t) | (~(w ^ t) & lt)), "all four equal" one per level,
// and one more merges lt into the flips: 9 instead of the 12 of the level-by-level form of lazy_bernoulli_a.
__device__ __forceinline__ uint32_t lt_step(uint32_t w, uint32_t t, uint32_t lt) {      // (~w & t) | (~(w ^ t) & lt)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x8E;" : "=r"(d) : "r"(w), "r"(t), "r"(lt));
    return d;
}
__device__ __forceinline__ uint32_t eq_step(uint32_t w, uint32_t t, uint32_t u) {        // u & ~(w ^ t)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x82;" : "=r"(d) : "r"(w), "r"(t), "r"(u));
    return d;
}
__device__ __forceinline__ void levels4(const uint4& w, const uint4& tb, uint32_t& und, uint32_t& e) {
    uint32_t lt = ~w.w & tb.w;
    lt = lt_step(w.z, tb.z, lt);
    lt = lt_step(w.y, tb.y, lt);
    lt = lt_step(w.x, tb.x, lt);
    e |= und & lt;
    und = eq_step(w.x, tb.x, und);
    und = eq_step(w.y, tb.y, und);
    und = eq_step(w.z, tb.z, und);
    und = eq_step(w.w, tb.w, und);
}

struct FlipWords {
    uint32_t e0, e1, e2, e3;          // trial A output 0, 1; trial B output 0, 1
};

//   a_q: this warp's queue in shared memory, 128 items x {owner | word << 5, undecided mask -> flips}
//   vm : the valid steps of the block (undecided lanes to start with)
__device__ __forceinline__ FlipWords flip_words4(uint32_t cb, uint32_t c1A, uint32_t c2A, uint32_t c1B, uint32_t c2B, uint32_t c3,
                                                 uint32_t vm, uint32_t a_tb, int ncalls, uint32_t a_q, uint32_t lane,
                                                 uint32_t bs, const Params& P) {
    uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0, u0 = vm, u1 = vm, u2 = vm, u3 = vm;
    if (ncalls > 0) {
        const uint4 tb = lds_v4(a_tb);
        const uint4 wa = philox10(cb, c1A, c2A, c3, P), wb = philox10(cb | 8u, c1A, c2A, c3, P);
        levels4(wa, tb, u0, e0);
        levels4(wb, tb, u1, e1);
        const uint4 wc = philox10(cb, c1B, c2B, c3, P), wd = philox10(cb | 8u, c1B, c2B, c3, P);
        levels4(wc, tb, u2, e2);
        levels4(wd, tb, u3, e3);
    }
    if (ncalls > 1) {
        const uint4 tb = lds_v4(a_tb + 16u);
        const uint4 wa = philox10(cb | 1u, c1A, c2A, c3, P), wb = philox10(cb | 9u, c1A, c2A, c3, P);
        levels4(wa, tb, u0, e0);
        levels4(wb, tb, u1, e1);
        const uint4 wc = philox10(cb | 1u, c1B, c2B, c3, P), wd = philox10(cb | 9u, c1B, c2B, c3, P);
        levels4(wc, tb, u2, e2);
        levels4(wd, tb, u3, e3);
    }
    if (ncalls > 2 && __any_sync(0xFFFFFFFFu, (u0 | u1 | u2 | u3) != 0u)) {
        const uint32_t lt = ~(0xFFFFFFFFu << lane);
        const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, u0 != 0u), b1 = __ballot_sync(0xFFFFFFFFu, u1 != 0u);
        const uint32_t b2 = __ballot_sync(0xFFFFFFFFu, u2 != 0u), b3 = __ballot_sync(0xFFFFFFFFu, u3 != 0u);
        const uint32_t n0 = __popc(b0), n1 = n0 + __popc(b1), n2 = n1 + __popc(b2), total = n2 + __popc(b3);
        const uint32_t p0 = a_q + 8u * __popc(b0 & lt), p1 = a_q + 8u * (n0 + __popc(b1 & lt));
        const uint32_t p2 = a_q + 8u * (n1 + __popc(b2 & lt)), p3 = a_q + 8u * (n2 + __popc(b3 & lt));
        if (u0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(p0), "r"(lane), "r"(u0) : "memory");
        if (u1) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(p1), "r"(lane | 32u), "r"(u1) : "memory");
        if (u2) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(p2), "r"(lane | 64u), "r"(u2) : "memory");
        if (u3) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(p3), "r"(lane | 96u), "r"(u3) : "memory");
        __syncwarp();
#pragma unroll 1
        for (uint32_t base = 0; base < total; base += 32u) {
            const uint32_t i = base + lane;
            const bool mine = i < total;
            uint2 it = make_uint2(lane, 0u);
            if (mine) it = lds_v2(a_q + 8u * i);
            uint32_t t1 = __shfl_sync(0xFFFFFFFFu, c1A, (int)(it.x & 31u));             // the owner's trial id (trial A)
            uint32_t t2 = __shfl_sync(0xFFFFFFFFu, c2A, (int)(it.x & 31u));
            if (it.x & 64u) {                                                           // words 2, 3: its trial B
                t1 += bs;
                t2 += t1 < bs ? 1u : 0u;
            }
            if (mine) {
                uint32_t und = it.y, e = 0;
                uint32_t c0 = cb | ((it.x & 32u) >> 2) | 2u;
                int k = 2;
                do {
                    const uint4 w = philox10(c0, t1, t2, c3, P);
                    const uint4 tb = lds_v4(a_tb + 16u * (uint32_t)k);
                    levels4(w, tb, und, e);
                    ++c0;
                    ++k;
                } while (und != 0u && k < ncalls);
                asm volatile("st.shared.u32 [%0], %1;" :: "r"(a_q + 8u * i + 4u), "r"(e) : "memory");
            }
        }
        __syncwarp();
        if (u0) e0 |= lds_u32(p0 + 4u);
        if (u1) e1 |= lds_u32(p1 + 4u);
        if (u2) e2 |= lds_u32(p2 + 4u);
        if (u3) e3 |= lds_u32(p3 + 4u);
        __syncwarp();                                   // the queue is reused by the next block
    }
    FlipWords f;
    f.e0 = e0; f.e1 = e1; f.e2 = e2; f.e3 = e3;
    return f;
}

// (a & m) | (b & ~m) as one LOP3
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t m) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
}

// ------------------------------------------------------------------------------------------ driver (n = 2)
// MM = encoder memory when known at compile time (ACS kernels), 0 = read it from the parameters
template <int LLS, int MM, class Eng>
__device__ __forceinline__ void run_trial_n2(const Params& P, const DevSeg& sg, bool active, unsigned long long trial,
                                             unsigned long long tl, unsigned long long ntr, const uint4* tbm, Eng& eng) {
    const int m = MM ? MM : P.m;
    const uint32_t N = sg.N;
    const int ncalls = sg.dmin > 31u ? 0 : (int)((31u - sg.dmin) / 4u + 1u);      // Philox calls per flip word, at most
    const bool philox = P.src_mode == MVD_SRC_PHILOX;
    // Philox counter words, activity mask and mask-table address pinned in registers: otherwise they are
    // re-derived from blockIdx / threadIdx / the segment record before every lazy loop (see detect2p_kernel)
    uint32_t c1 = (uint32_t)trial, c2 = (uint32_t)(trial >> 32), c3 = sg.stream, am = active ? 0xFFFFFFFFu : 0u;
    uint32_t a_tbp = (uint32_t)__cvta_generic_to_shared(tbm);
    asm volatile("" : "+r"(c1), "+r"(c2), "+r"(am));
    const uint32_t taps0 = sg.enc_taps[0], taps1 = sg.enc_taps[1];
    uint32_t prevU = 0;
    const uint32_t nsb = (N + 127u) >> 7;
    for (uint32_t sb = 0; sb < nsb; ++sb) {
        uint4 Uw = make_uint4(0, 0, 0, 0), E0 = Uw, E1 = Uw;
        if (philox) {
            Uw = philox10(((4u * sb) << 6) | 32u, c1, c2, c3, P);
        } else if (active) {
            const uint4* base = P.bits + sg.bits_offset + (unsigned long long)sb * 3ull * ntr + tl;
            Uw = __ldg(base);
            E0 = __ldg(base + ntr);
            E1 = __ldg(base + 2ull * ntr);
        }
        if (!sg.random_input) Uw = make_uint4(0, 0, 0, 0);
#pragma unroll 1
        for (int w = 0; w < 4; ++w) {
            const uint32_t t0 = sb * 128u + (uint32_t)w * 32u;
            if (t0 >= N) break;
            const uint32_t valid = min(32u, N - t0);
            const uint32_t vmask = valid == 32u ? 0xFFFFFFFFu : ((1u << valid) - 1u);
            const uint32_t U = pick(Uw, w);
            uint32_t e0, e1;
            if (philox) {
                uint32_t cb = (4u * sb + (uint32_t)w) << 6, vm = vmask & am;
                asm volatile("" : "+r"(vm));
                e0 = e1 = 0u;
#pragma unroll 1
                for (int j = 0; j < 2; ++j) {                               // one copy of the lazy loop for both outputs
                    e0 = e1;
                    e1 = lazy_bernoulli_a(cb | ((uint32_t)j << 3), c1, c2, c3, a_tbp, ncalls, vm, P);
                }
            } else {
                e0 = pick(E0, w);
                e1 = pick(E1, w);
            }
            // bit-parallel encoder (viterbi_markov.py:82-106 for k = 1): 32 steps per XOR
            uint32_t o0 = U & (0u - (taps0 & 1u)), o1 = U & (0u - (taps1 & 1u));
            if (MM) {
#pragma unroll
                for (int i = 1; i <= (MM ? MM : 1); ++i) {
                    const uint32_t sh = __funnelshift_l(prevU, U, i);
                    o0 ^= sh & (0u - ((taps0 >> i) & 1u));
                    o1 ^= sh & (0u - ((taps1 >> i) & 1u));
                }
            } else {
#pragma unroll 1
                for (int i = 1; i <= m; ++i) {
                    const uint32_t sh = __funnelshift_l(prevU, U, i);
                    o0 ^= sh & (0u - ((taps0 >> i) & 1u));
                    o1 ^= sh & (0u - ((taps1 >> i) & 1u));
                }
            }
            prevU = U;
            const uint32_t R0 = o0 ^ e0, R1 = o1 ^ e1;            // BSC
            // received word of step t = (R0 bit t, R1 bit t), first output is the MSB: two bit-selects put the pairs
            // of the even steps into one word and those of the odd steps into another, at bits (t | 1, t & ~1)
            const uint32_t wev = bitsel(R1, R0 << 1, 0x55555555u), wod = bitsel(R1 >> 1, R0, 0x55555555u);
            auto oct = [&](uint32_t ev, uint32_t od) {                     // 8 steps = bits 0..7 of ev and od
                eng.step(roff<LLS, 0>(ev));
                eng.step(roff<LLS, 0>(od));
                eng.step(roff<LLS, 1>(ev));
                eng.step(roff<LLS, 1>(od));
                eng.step(roff<LLS, 2>(ev));
                eng.step(roff<LLS, 2>(od));
                eng.step(roff<LLS, 3>(ev));
                eng.step(roff<LLS, 3>(od));
            };
            if (valid == 32u) {
                uint32_t ev = wev, od = wod;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {                                // 16 steps per iteration
                    oct(ev, od);
                    oct(ev >> 8, od >> 8);
                    ev >>= 16;
                    od >>= 16;
                }
            } else {
#pragma unroll 1
                for (uint32_t c = 0; c < valid; c += 8u) {
                    const uint32_t ev = wev >> c, od = wod >> c;
                    if (c + 8u <= valid) {
                        oct(ev, od);
                    } else {
                        for (uint32_t j = 0; j < valid - c; ++j) eng.step(((((j & 1u) ? od : ev) >> (j & ~1u)) & 3u) << LLS);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ driver (n = 3)
// Rate-1/3 codes on the NEXT-walk engines (which do not care how many received words there are: a step takes r << LLS).
// Same bit source, same encoder; the received word of step t is (R0 bit t, R1 bit t, R2 bit t), first output = MSB,
// assembled with three shifts and three LOP3 per step (8 steps straight-line, shift counts as immediates).
template <int LLS, class Eng>
__device__ __forceinline__ void run_trial_n3(const Params& P, const DevSeg& sg, bool active, unsigned long long trial,
                                             unsigned long long tl, unsigned long long ntr, const uint4* tbm, Eng& eng) {
    const int m = P.m;
    const uint32_t N = sg.N;
    const int ncalls = sg.dmin > 31u ? 0 : (int)((31u - sg.dmin) / 4u + 1u);
    const bool philox = P.src_mode == MVD_SRC_PHILOX;
    uint32_t c1 = (uint32_t)trial, c2 = (uint32_t)(trial >> 32), c3 = sg.stream, am = active ? 0xFFFFFFFFu : 0u;
    uint32_t a_tbp = (uint32_t)__cvta_generic_to_shared(tbm);
    asm volatile("" : "+r"(c1), "+r"(c2), "+r"(am));
    const uint32_t taps0 = sg.enc_taps[0], taps1 = sg.enc_taps[1], taps2 = sg.enc_taps[2];
    uint32_t prevU = 0;
    const uint32_t nsb = (N + 127u) >> 7;
    for (uint32_t sb = 0; sb < nsb; ++sb) {
        uint4 Uw = make_uint4(0, 0, 0, 0), E0 = Uw, E1 = Uw, E2 = Uw;
        if (philox) {
            Uw = philox10(((4u * sb) << 6) | 32u, c1, c2, c3, P);
        } else if (active) {
            const uint4* base = P.bits + sg.bits_offset + (unsigned long long)sb * 4ull * ntr + tl;
            Uw = __ldg(base);
            E0 = __ldg(base + ntr);
            E1 = __ldg(base + 2ull * ntr);
            E2 = __ldg(base + 3ull * ntr);
        }
        if (!sg.random_input) Uw = make_uint4(0, 0, 0, 0);
#pragma unroll 1
        for (int w = 0; w < 4; ++w) {
            const uint32_t t0 = sb * 128u + (uint32_t)w * 32u;
            if (t0 >= N) break;
            const uint32_t valid = min(32u, N - t0);
            const uint32_t vmask = valid == 32u ? 0xFFFFFFFFu : ((1u << valid) - 1u);
            const uint32_t U = pick(Uw, w);
            uint32_t e0, e1, e2;
            if (philox) {
                uint32_t cb = (4u * sb + (uint32_t)w) << 6, vm = vmask & am;
                asm volatile("" : "+r"(vm));
                e0 = e1 = e2 = 0u;
#pragma unroll 1
                for (int j = 0; j < 3; ++j) {                               // one copy of the lazy loop for the three outputs
                    e0 = e1;
                    e1 = e2;
                    e2 = lazy_bernoulli_a(cb | ((uint32_t)j << 3), c1, c2, c3, a_tbp, ncalls, vm, P);
                }
            } else {
                e0 = pick(E0, w);
                e1 = pick(E1, w);
                e2 = pick(E2, w);
            }
            uint32_t o0 = U & (0u - (taps0 & 1u)), o1 = U & (0u - (taps1 & 1u)), o2 = U & (0u - (taps2 & 1u));
#pragma unroll 1
            for (int i = 1; i <= m; ++i) {
                const uint32_t sh = __funnelshift_l(prevU, U, i);
                o0 ^= sh & (0u - ((taps0 >> i) & 1u));
                o1 ^= sh & (0u - ((taps1 >> i) & 1u));
                o2 ^= sh & (0u - ((taps2 >> i) & 1u));
            }
            prevU = U;
            const uint32_t R0 = o0 ^ e0, R1 = o1 ^ e1, R2 = o2 ^ e2;        // BSC
            if (valid == 32u) {
#pragma unroll 1
                for (int h = 0; h < 4; ++h) {                               // 8 steps per iteration
                    // bits 0..7 of the three words moved to where step J wants them: output 0 at LLS + 2, 1 at LLS + 1, 2 at LLS
                    const uint32_t x0 = (R0 >> (8 * h)) << (LLS + 2), x1 = (R1 >> (8 * h)) << (LLS + 1), x2 = (R2 >> (8 * h)) << LLS;
#pragma unroll
                    for (int J = 0; J < 8; ++J)
                        eng.step(((x0 >> J) & (4u << LLS)) | ((x1 >> J) & (2u << LLS)) | ((x2 >> J) & (1u << LLS)));
                }
            } else {
                for (uint32_t t = 0; t < valid; ++t)
                    eng.step(((((R0 >> t) & 1u) << 2) | (((R1 >> t) & 1u) << 1) | ((R2 >> t) & 1u)) << LLS);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ engines
// ACS: Eq. 4-5 in registers (AcsCore), then metric vector -> Markov state.
//   sx = shared-memory byte address of this lane's copy of the log-likelihood row of the current state.
template <int LK, int M, int LLS, bool GT = false>
struct Acs2Engine {
    static constexpr int NP = AcsCore<M>::NP;
    // branch metrics: row r = 2 * NP words, stored as NV 16-byte (8-byte for m = 1) planes; plane i of
    // row r, copy c lives at bm_base + i * (4 << LLS) + (r << LLS) + 16 c -- the layout of the log rows,
    // so lanes of a quarter warp never meet in a bank whatever their r.
    static constexpr int NV = NP >= 2 ? NP / 2 : 1;
    AcsCore<M> core;
    const unsigned char* sm;         // shared memory: branch metrics and (GT = false) every table
    const unsigned char* tbg;        // GT: log-likelihood rows in global memory (L2)
    const uint32_t* hvg;             // GT: hash values (state * R) and keys in global memory
    const uint32_t* hkg;
    uint32_t sx, bm_base;
    uint32_t key_mul, key_add;       // DIRECT
    uint32_t ll_lane, st_base, hmask;   // HASH
    double a1, a0;

    __device__ __forceinline__ void step(uint32_t r_off) {
        const double2 v = GT ? __ldg(reinterpret_cast<const double2*>(tbg + sx + r_off))
                             : *reinterpret_cast<const double2*>(sm + sx + r_off);   // edge (state, r)
        a1 += v.x;                                                                // Pd_plotter.py:114-115,
        a0 += v.y;                                                                // in step order
        uint32_t bm[2 * NP];
        const uint32_t boff = bm_base + r_off;                                    // bm_base includes the lane copy
        if (NP == 1) {
            const uint2 b = *reinterpret_cast<const uint2*>(sm + boff);
            bm[0] = b.x;
            bm[1] = b.y;
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const uint4 b = *reinterpret_cast<const uint4*>(sm + boff + i * (4 << LLS));
                bm[4 * i] = b.x;
                bm[4 * i + 1] = b.y;
                bm[4 * i + 2] = b.z;
                bm[4 * i + 3] = b.w;
            }
        }
        core.step(bm);                                                            // Eq. 4 + Eq. 5
        if (LK == LK_DIRECT) {
            // D[0] = s0 | s1 << 16, D[1] = s2 | s3 << 16, every metric < 2^b (b = 2 for m = 2, 4 for m = 1):
            //   t  = s0 | s2 << 2b  |  (s1 | s3 << 2b) << 16
            //   t * ((2^(16+b) + 1) << 7): bits 31..16 = (lo * 2^b + hi) << 7 = key << 7
            // + key_add = (table base + 4 * lane) << 16, so the upper half is the byte address of this
            // lane's copy of entry `key`.
            constexpr int B = (M == 1) ? 4 : 2;
            uint32_t t = core.D[0];
            if (NP > 1) t += core.D[1] << (2 * B);
            const uint32_t u = t * key_mul + key_add;
            sx = *reinterpret_cast<const uint32_t*>(sm + (u >> 16));
        } else {
            uint32_t kw[AcsCore<M>::KW];
            core.key(kw);
            uint32_t slot = key_hash(kw, AcsCore<M>::KW) & hmask;
            const uint32_t* hv = GT ? hvg : reinterpret_cast<const uint32_t*>(sm + st_base);
            const uint32_t* hk = GT ? hkg : reinterpret_cast<const uint32_t*>(sm + st_base) + (hmask + 1u);
            for (uint32_t probe = 0; probe <= hmask; ++probe) {
                bool same = true;
#pragma unroll
                for (int i = 0; i < AcsCore<M>::KW; ++i) same = same && (hk[(size_t)i * (hmask + 1u) + slot] == kw[i]);
                if (same) break;
                slot = (slot + 1u) & hmask;
            }
            const uint32_t val = hv[slot];                                        // state * R
            sx = ll_lane + (val << LLS);
        }
    }
};

// FSM: the same chain walked through NEXT[state][r].
//   LLS == 7: sx = byte address of this lane's copy of the current state's log-likelihood row, the
//             NEXT table is lane-replicated with the same 128-byte entry stride and stores such addresses;
//   else    : sx = state * R and NEXT is a plain table.
template <int LLS>
struct Fsm2Engine {
    const unsigned char* sm;
    uint32_t sx, ll_lane, nx_lane;
    double a1, a0;

    __device__ __forceinline__ void step(uint32_t r_off) {
        if (LLS == 7) {
            const uint32_t a = sx + r_off;
            const double2 v = *reinterpret_cast<const double2*>(sm + a);
            a1 += v.x;
            a0 += v.y;
            sx = *reinterpret_cast<const uint32_t*>(sm + a + nx_lane);      // nx_lane = NEXT copy - log copy
        } else {
            const uint32_t e = sx + (r_off >> LLS);
            const double2 v = *reinterpret_cast<const double2*>(sm + ll_lane + (e << LLS));
            a1 += v.x;
            a0 += v.y;
            sx = *reinterpret_cast<const uint32_t*>(sm + nx_lane + (e << 2));
        }
    }
};

// FSM, one LDS.128 per step: the 16-byte entry of edge (state, r) is {log P1 (f64), address of the next
// state's row, high word of the double c} with log Tref = c * unit exactly (c = 0 or a power of two: for
// rate-1/2 codes T(1/2) entries are 1/4, 1/2, 1 and their logs 2u, u, 0 with u = log(1/2)).  The
// reference's a0 += log Tref becomes a0 = fma(c, unit, a0): c * unit is exact, so the rounding -- and
// every bit of the running sum -- is the same.
template <bool GT>
struct Fsm1Engine {
    const unsigned char* sm;         // shared memory, or (GT) the packed table in global memory
    uint32_t sx;
    double unit, a1, a0;

    __device__ __forceinline__ void step(uint32_t r_off) {
        const uint4 v = GT ? __ldg(reinterpret_cast<const uint4*>(sm + sx + r_off))
                           : *reinterpret_cast<const uint4*>(sm + sx + r_off);
        a1 += __hiloint2double((int)v.y, (int)v.x);
        a0 = fma(__hiloint2double((int)v.w, 0), unit, a0);
        sx = v.z;
    }
};

// ------------------------------------------------------------------------------------------ kernel
// grid: (chunks of blockDim.x trials, segments of this launch); one thread = one trial.  The segment
// descriptors are kernel parameters, so everything derived from them (N, threshold, taps, decision
// rule) is warp-uniform and lives on the uniform datapath.
// dynamic shared memory (byte offsets in P.fp): [threshold masks][branch-metric replicas][state table][log-likelihood replicas]
// GT = true: the state / log-likelihood tables stay in global memory (L2-resident; S too large for
// shared memory, e.g. m = 4 with S = 25 751 ... 232 567): nothing but the threshold masks and the
// branch metrics is staged, LLS = 4 (no replicas).
// NOUT = 3: rate-1/3 codes (NEXT-walk engines only).
// BIG: blocks of DET2_BIG_BLOCK threads, two per SM (a shared-memory table that only fits twice; see plan_det2)
template <int LK, int M, int LLS, bool GT = false, int NOUT = 2, bool BIG = false>
__global__ void __launch_bounds__(BIG ? DET2_BIG_BLOCK : DET2_BLOCK, BIG ? 2 : ((LK == LK_FSM1 || LK == LK_FSM) ? 3 : 2))
detect2_kernel(const __grid_constant__ Params P, const __grid_constant__ SegBatch B) {
    static_assert(NOUT == 2 || (NOUT == 3 && (LK == LK_FSM1 || LK == LK_FSM)), "n = 3 runs on the NEXT-walk engines");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int REP = 1 << (LLS - 4);
    const DevSeg& sg = B.s[blockIdx.y];
    const uint32_t BS = blockDim.x;                                          // <= DET2_BLOCK, chosen by the host
    const unsigned long long ntr = sg.trial_end - sg.trial_begin;
    if ((unsigned long long)blockIdx.x * BS >= ntr) return;                  // uniform: shorter segment
    const uint32_t seg = sg.block_begin;                                     // global segment index
    const unsigned long long tl = (unsigned long long)blockIdx.x * BS + threadIdx.x;
    const bool active = tl < ntr;
    const unsigned long long trial = sg.trial_begin + tl;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t SR = P.SR;
    const uint32_t copy = (lane & (uint32_t)(REP - 1)) << 4;
    const uint32_t ll_lane = P.fp.off_ll + copy;

    // ---- stage the tables of this segment's p
    if (threadIdx.x < 32u)
        *reinterpret_cast<uint32_t*>(smem_raw + P.fp.off_tb + 4u * threadIdx.x) = 0u - ((sg.threshold >> (31u - threadIdx.x)) & 1u);
    const uint4* tbm = reinterpret_cast<const uint4*>(smem_raw + P.fp.off_tb);
    if (GT) {
    } else if (LK == LK_FSM1) {
        const double2* llg = P.ll + (size_t)sg.table * SR;
        for (uint32_t i = threadIdx.x; i < SR * REP; i += BS) {
            const uint32_t e = i >> (LLS - 4), c = i & (uint32_t)(REP - 1);
            const double lp = __ldg(llg + e).x;
            *reinterpret_cast<uint4*>(smem_raw + P.fp.off_ll + (e << LLS) + (c << 4)) =
                make_uint4((uint32_t)__double2loint(lp), (uint32_t)__double2hiint(lp),
                           P.fp.off_ll + (__ldg(P.nxt + e) << LLS) + (c << 4), __ldg(P.fp.tcode + e));
        }
    } else {
        const double2* llg = P.ll + (size_t)sg.table * SR;
        for (uint32_t i = threadIdx.x; i < SR * REP; i += BS) {
            const uint32_t e = i >> (LLS - 4), c = i & (uint32_t)(REP - 1);
            *reinterpret_cast<double2*>(smem_raw + P.fp.off_ll + (e << LLS) + (c << 4)) = __ldg(llg + e);
        }
    }
    if (LK == LK_FSM1) {
    } else if (LK == LK_FSM) {
        if (LLS == 7) {
            for (uint32_t i = threadIdx.x; i < SR * 32u; i += BS)
                *reinterpret_cast<uint32_t*>(smem_raw + P.fp.off_st + 4u * i) =
                    P.fp.off_ll + (__ldg(P.nxt + (i >> 5)) << 7) + (((i & 31u) & 7u) << 4);
        } else {
            for (uint32_t i = threadIdx.x; i < SR; i += BS)
                *reinterpret_cast<uint32_t*>(smem_raw + P.fp.off_st + 4u * i) = __ldg(P.nxt + i);
        }
    } else {
        constexpr int NP = AcsCore<M>::NP;
        constexpr int NV = Acs2Engine<LK, M, LLS, GT>::NV;
        constexpr int WPV = NP >= 2 ? 4 : 2;                                 // words per plane vector
        // word w of row r -> plane w / WPV, every copy c
        for (uint32_t i = threadIdx.x; i < 4u * 2u * NP * REP; i += BS) {
            const uint32_t c = i % REP, rw = i / REP, r = rw / (2u * NP), w = rw % (2u * NP);
            *reinterpret_cast<uint32_t*>(smem_raw + P.fp.off_bm + (w / WPV) * (4u << LLS) + (r << LLS) + (c << 4) +
                                         4u * (w % WPV)) = P.bm[rw];
        }
        (void)NV;
        if (GT) {
        } else if (LK == LK_DIRECT) {
            for (uint32_t i = threadIdx.x; i < P.fp.nkeys * 32u; i += BS) {
                const uint32_t st = P.fp.dstate[i >> 5];          // 0xFFFF: not a state (never looked up)
                const uint32_t row = st == 0xFFFFu ? 0u : st * 4u;
                *reinterpret_cast<uint32_t*>(smem_raw + P.fp.off_st + 4u * i) =
                    P.fp.off_ll + (row << LLS) + (((i & 31u) & (uint32_t)(REP - 1)) << 4);
            }
        } else {
            constexpr int KW = AcsCore<M>::KW;
            for (uint32_t i = threadIdx.x; i < P.hcap; i += BS) {
                const uint32_t v = P.hvals[i];
                *reinterpret_cast<uint32_t*>(smem_raw + P.fp.off_st + 4u * i) = v == MVD_EMPTY ? 0u : v;
#pragma unroll
                for (int w = 0; w < KW; ++w)
                    *reinterpret_cast<uint32_t*>(smem_raw + P.fp.off_st + 4u * (P.hcap * (1u + w) + i)) =
                        v == MVD_EMPTY ? 0xFFFFFFFFu : P.hkeys[(size_t)w * P.hcap + i];
            }
        }
    }
    __syncthreads();

    double a1, a0;
    if (LK == LK_FSM1) {
        Fsm1Engine<GT> eng;
        eng.sm = GT ? reinterpret_cast<const unsigned char*>(P.fp.gfsm1) + (size_t)sg.table * SR * 16u : smem_raw;
        eng.sx = GT ? 0u : ll_lane;
        eng.unit = P.fp.tref_unit;
        eng.a1 = 0.0;
        eng.a0 = 0.0;
        if (NOUT == 3) run_trial_n3<LLS>(P, sg, active, trial, tl, ntr, tbm, eng);
        else run_trial_n2<LLS, 0>(P, sg, active, trial, tl, ntr, tbm, eng);
        a1 = eng.a1;
        a0 = eng.a0;
    } else if (LK == LK_FSM) {
        Fsm2Engine<LLS> eng;
        eng.sm = smem_raw;
        eng.sx = LLS == 7 ? ll_lane : 0u;
        eng.ll_lane = ll_lane;
        eng.nx_lane = LLS == 7 ? (P.fp.off_st + lane * 4u) - ll_lane : P.fp.off_st;
        eng.a1 = 0.0;
        eng.a0 = 0.0;
        if (NOUT == 3) run_trial_n3<LLS>(P, sg, active, trial, tl, ntr, tbm, eng);
        else run_trial_n2<LLS, 0>(P, sg, active, trial, tl, ntr, tbm, eng);
        a1 = eng.a1;
        a0 = eng.a0;
    } else {
        Acs2Engine<LK, M, LLS, GT> eng;
        eng.core.reset();
        eng.sm = smem_raw;
        eng.tbg = reinterpret_cast<const unsigned char*>(P.ll + (size_t)sg.table * SR);
        eng.hvg = P.hvals;
        eng.hkg = P.hkeys;
        eng.st_base = P.fp.off_st;
        eng.sx = GT ? 0u : ll_lane;                               // state 0 = the all-zero vector (viterbi_markov.py:177)
        eng.bm_base = P.fp.off_bm + copy;
        eng.key_mul = P.fp.key_mul;
        eng.key_add = (P.fp.off_st + lane * 4u) << 16;
        eng.ll_lane = GT ? 0u : ll_lane;
        eng.hmask = P.hcap - 1u;
        eng.a1 = 0.0;
        eng.a0 = 0.0;
        run_trial_n2<LLS, M>(P, sg, active, trial, tl, ntr, tbm, eng);
        a1 = eng.a1;
        a0 = eng.a0;
    }

    const bool win = active && (sg.decide == 0 ? (a1 > a0) : (a1 <= a0));          // Pd_plotter.py:215 / :222
    const int c = __syncthreads_count(win ? 1 : 0);
    if (threadIdx.x == 0 && c) {
        atomicAdd(P.tallies + seg, (unsigned long long)c);
        if (P.tallies2) atomicAdd(P.tallies2 + seg, (unsigned long long)c);
    }
    if (P.logp && active) {
        double2* o = reinterpret_cast<double2*>(P.logp) + sg.out_offset + tl;
        *o = make_double2(a1, a0);
    }
}

// =========================================================================================== pair kernel
// m = 2, n = 2, direct state table: TWO trials per thread, the 16x2 SIMD lanes run ACROSS the two
// trials instead of across trellis states.  Q[s] = (D_A[s], D_B[s]); one VIADDMNMX.U16x2 per new
// state does Eq. 4 for both trials, so the four PRMT broadcasts and the PRMT of the global minimum of
// the state-packed layout disappear (ALU pipe: 6.5 instead of 14 instructions per trellis step; the ALU
// pipe is what bounds this kernel).  Branch metrics come as (metric for r_A, metric for r_B) pairs
// from a 16-row table indexed by (r_A, r_B).
//
// All table reads use absolute shared-memory addresses (ld.shared) whose alignment lets a single
// LOP3 build them: log rows are 512-byte aligned (address = row | r * 128), the branch-metric planes
// 2048-byte aligned (address = plane | r_A * 128 | r_B * 512 | copy * 16).

__device__ __forceinline__ uint32_t madlo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// metrics are kept times 128 with this base in both lanes: an un-normalised ANTI step can lower a lane by n * 128 = 256,
// 16 steps by 4 096
#define PAIR_BASE2 0x20002000u

struct PairEngine {
    uint32_t Q0, Q1, Q2, Q3;          // (trial A, trial B) metrics of trellis states 0..3, times 128
    uint32_t sxA, sxB;                // absolute address of this lane's copy of the current log row
    uint32_t kbm, kst;                // branch-metric plane 0 | copy * 16 ; state table (32 KB-aligned) | lane * 4
    uint32_t km1;                     // -1 from the kernel parameters: n - d as an IMAD on the FMA pipe (a visible constant
                                      // would make it an IADD3 on the ALU pipe, the busiest one).  Measured and rejected: the
                                      // state-table addresses and the >> 8 / >> 16 as IMAD.HI (7.68e11 -> 7.24e11 steps/s)
    double a1A, a0A, a1B, a0B;

    // sAB: r_A at bits 7..8 and r_B at bits 9..10; sB7: r_B at bits 7..8 (other bits arbitrary).
    // The state table is indexed by an OFFSET-INVARIANT key: key(D) = sum_s c_s D[s] + bias with sum_s c_s = 0
    // (coefficients found by the host: injective on this decoder's metric vectors, keys in [0, 256)), so
    // key(D') = key(D' - min D') and a step needs the minimum only when it normalises (Eq. 5): four IMADs on the
    // FMA pipe on the packed pair -- exact modulo 2^32 whatever the low lane carries into the high one, because
    // the true result of either lane is in [0, 2^15) -- and no VIMNMX3 / VIMNMX on the ALU pipe.
    // NORM = false leaves the metrics un-normalised (Eq. 5 deferred); the last step of a 16-step stretch runs with
    // NORM = true and re-bases every lane to PAIR_BASE + (D' - min D'), so every loop boundary sees the reference's
    // normalised vector (plus a constant the key does not see).
    // ANTI: every generator has its first and its last tap set (e.g. (7,5)), so the four branches of a butterfly
    // carry the labels X, ~X, ~X, X and d(~X, r) = n - d(X, r).  With x = d(X, r) the butterfly is
    //   D'[0] = x + min(D[0], D[2] + (n - 2x)),   D'[1] = x + min(D[0] + (n - 2x), D[2]),
    // i.e. ONE VIADDMNMX per new state with no separate addition, both results carrying the same offset x; the
    // second butterfly carries y, so only its two results need an addition (of y - x) and the common offset x is
    // dropped (the key does not see it; a lane falls by at most n per step, hence PAIR_BASE).  One 16-byte read
    // {n - 2x, n - 2y, y - x} per step; the general table takes two and four additions.
    template <bool NORM, bool ANTI>
    __device__ __forceinline__ void step(uint32_t sAB, uint32_t sB7, const Params& P) {
        const double2 vA = lds_d2(sxA | (sAB & 0x180u));
        const double2 vB = lds_d2(sxB | (sB7 & 0x180u));
        a1A += vA.x;
        a0A += vA.y;
        a1B += vB.x;
        a0B += vB.y;
        const uint32_t boff = kbm | (sAB & 0x780u);
        uint32_t n0, n1, n2, n3;
        if (ANTI) {
            const uint4 t = lds_v4(boff);                  // n - 2 d(0 -> 0), n - 2 d(1 -> 2), d(1 -> 2) - d(0 -> 0) for (r_A, r_B)
            n0 = __viaddmin_u16x2(Q2, t.x, Q0);            // Eq. 4, both trials
            n1 = __viaddmin_u16x2(Q0, t.x, Q2);
            n2 = __viaddmin_u16x2(Q3, t.y, Q1) + t.z;
            n3 = __viaddmin_u16x2(Q1, t.y, Q3) + t.z;
        } else {
            const uint4 b0 = lds_v4(boff);                 // ns 0: (pred 0, pred 2), ns 1: (pred 0, pred 2)
            const uint4 b1 = lds_v4(boff + 2048u);         // ns 2: (pred 1, pred 3), ns 3: (pred 1, pred 3)
            n0 = __viaddmin_u16x2(Q0, b0.x, Q2 + b0.y);    // Eq. 4, both trials
            n1 = __viaddmin_u16x2(Q0, b0.z, Q2 + b0.w);
            n2 = __viaddmin_u16x2(Q1, b1.x, Q3 + b1.y);
            n3 = __viaddmin_u16x2(Q1, b1.z, Q3 + b1.w);
        }
        const uint32_t t7 = madlo(n3, P.fp.kc[3], madlo(n2, P.fp.kc[2], madlo(n1, P.fp.kc[1], madlo(n0, P.fp.kc[0], P.fp.kcb))));
        if (NORM) {
            const uint32_t mn = __vminu2(__vimin3_u16x2(n0, n1, n2), n3) - PAIR_BASE2;   // per-trial minimum, less the base
            Q0 = n0 - mn;                                                   // Eq. 5 (no borrow: every lane >= its minimum)
            Q1 = n1 - mn;
            Q2 = n2 - mn;
            Q3 = n3 - mn;
        } else {
            Q0 = n0;
            Q1 = n1;
            Q2 = n2;
            Q3 = n3;
        }
        uint32_t aA;
        asm("lop3.b32 %0, %1, 0xFFFF, %2, 0xEA;" : "=r"(aA) : "r"(t7), "r"(kst));       // (t7 & 0xFFFF) | kst
        sxA = lds_u32(aA);
        sxB = lds_u32((t7 >> 16) + kst);
    }

    // Eq. 5 on its own (the ragged last block of a trial)
    __device__ __forceinline__ void normalise() {
        const uint32_t mn = __vminu2(__vimin3_u16x2(Q0, Q1, Q2), Q3) - PAIR_BASE2;
        Q0 -= mn;
        Q1 -= mn;
        Q2 -= mn;
        Q3 -= mn;
    }
};

template <int PHILOX, int ANTI>
__global__ void __launch_bounds__(DET2P_BLOCK, 3) detect2p_kernel(const __grid_constant__ Params P,
                                                                  const __grid_constant__ SegBatch B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevSeg& sg = B.s[blockIdx.y];
    const unsigned long long ntr = sg.trial_end - sg.trial_begin;
    const uint32_t BS = blockDim.x;
    const unsigned long long blk0 = (unsigned long long)blockIdx.x * (2u * BS);
    if (blk0 >= ntr) return;
    const uint32_t seg = sg.block_begin;
    const unsigned long long tlA = blk0 + threadIdx.x, tlB = tlA + BS;
    const bool actA = tlA < ntr, actB = tlB < ntr;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t SR = P.SR;
    // absolute shared addresses, laid out downwards from the state table, which sits at a 32 KB-aligned address so
    // that `key * 128 | table | lane * 4` is ONE LOP3:
    //   [straggler queues: 1 KB per warp][...][masks 128][bm plane 0 (2 KB)][bm plane 1 (2 KB)][log rows SR x 128][state 256 x 128]
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t a_st = (sbase + (uint32_t)DET2P_QUEUES + 128u + 2048u + 4096u + (SR << 7) + 32767u) & ~32767u;
    const uint32_t a_ll = a_st - (SR << 7);
    const uint32_t a_bm = (a_ll - 4096u) & ~2047u;
    const uint32_t a_tb = a_bm - 128u;
    uint32_t a_sq = sbase + (threadIdx.x >> 5) * 1024u;      // this warp's straggler queue (flip_words4)
    unsigned char* g = smem_raw - sbase;                      // generic pointer of shared address 0
    {
        uint32_t dyn;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if (a_st + 32768u - sbase > dyn) {                    // the host sized the window for another base address
            if (threadIdx.x == 0) atomicOr(P.error_flag, 4);
            return;
        }
    }

    if (threadIdx.x < 32u)
        *reinterpret_cast<uint32_t*>(g + a_tb + 4u * threadIdx.x) = 0u - ((sg.threshold >> (31u - threadIdx.x)) & 1u);
    {
        const double2* llg = P.ll + (size_t)sg.table * SR;
        for (uint32_t i = threadIdx.x; i < SR * 8u; i += BS)
            *reinterpret_cast<double2*>(g + a_ll + ((i >> 3) << 7) + ((i & 7u) << 4)) = __ldg(llg + (i >> 3));
    }
    // pair branch metrics: row (rA | rB << 2), word (ns, b): lo = d(pred_b -> ns | rA), hi = ... | rB
    // P.bm[r][2 g + b] = (d(pred -> 2g), d(pred -> 2g+1)) for pred = g + 2 b
    if (ANTI) {
        // row (rA | rB << 2), 8 copies of 16 bytes: {n - 2x, n - 2y, y - x, 0} with x = d(0 -> 0), y = d(1 -> 2), times 128;
        // the first two are added lane by lane (VIADDMNMX), the third as one 32-bit word
        for (uint32_t i = threadIdx.x; i < 16u * 8u; i += BS) {
            const uint32_t c = i & 7u, row = i >> 3, rA = row & 3u, rB = row >> 2;
            const int xA = (int)(P.bm[rA * 4u] & 0xFFFFu), xB = (int)(P.bm[rB * 4u] & 0xFFFFu);
            const int yA = (int)(P.bm[rA * 4u + 2u] & 0xFFFFu), yB = (int)(P.bm[rB * 4u + 2u] & 0xFFFFu);
            const uint32_t d0 = ((uint32_t)((2 - 2 * xA) * 128) & 0xFFFFu) | ((uint32_t)((2 - 2 * xB) * 128) << 16);
            const uint32_t d1 = ((uint32_t)((2 - 2 * yA) * 128) & 0xFFFFu) | ((uint32_t)((2 - 2 * yB) * 128) << 16);
            const uint32_t z = (uint32_t)((yA - xA) * 128 + (yB - xB) * 128 * 65536);
            *reinterpret_cast<uint4*>(g + a_bm + (rA << 7) + (rB << 9) + (c << 4)) = make_uint4(d0, d1, z, 0u);
        }
    }
    for (uint32_t i = ANTI ? 16u * 8u * 8u : threadIdx.x; i < 16u * 8u * 8u; i += BS) {
        const uint32_t c = i & 7u, wd = (i >> 3) & 7u, row = i >> 6;
        const uint32_t rA = row & 3u, rB = row >> 2, ns = wd >> 1, b = wd & 1u, gg = ns >> 1, h = ns & 1u;
        const uint32_t wa = P.bm[rA * 4u + 2u * gg + b], wb = P.bm[rB * 4u + 2u * gg + b];
        const uint32_t da = h ? (wa >> 16) : (wa & 0xFFFFu), db = h ? (wb >> 16) : (wb & 0xFFFFu);
        *reinterpret_cast<uint32_t*>(g + a_bm + (wd >> 2) * 2048u + (rA << 7) + (rB << 9) + (c << 4) + 4u * (wd & 3u)) =
            (da << 7) | (db << 23);                              // times 128: the metric pairs then form the table offset by Horner
    }
    for (uint32_t i = threadIdx.x; i < 256u * 32u; i += BS) {
        const uint32_t st = P.fp.dstate2[i >> 5];                      // by the offset-invariant key (PairEngine::step)
        const uint32_t row = st == 0xFFFFu ? 0u : st * 4u;
        *reinterpret_cast<uint32_t*>(g + a_st + 4u * i) = a_ll + (row << 7) + (((i & 31u) & 7u) << 4);
    }
    __syncthreads();

    PairEngine eng;
    eng.Q0 = eng.Q1 = eng.Q2 = eng.Q3 = PAIR_BASE2;       // D_0 = 0 (plus the base)
    eng.sxA = eng.sxB = a_ll + ((lane & 7u) << 4);            // state 0 = the all-zero vector
    eng.kbm = a_bm + ((lane & 7u) << 4);
    eng.kst = a_st + lane * 4u;
    eng.a1A = eng.a0A = eng.a1B = eng.a0B = 0.0;
    eng.km1 = P.fma_km1;

    const uint32_t N = sg.N;
    const int ncalls = sg.dmin > 31u ? 0 : (int)((31u - sg.dmin) / 4u + 1u);
    constexpr bool philox = PHILOX != 0;               // the bit source is a template parameter: the bit-stream words are not live in the Philox instance
    const unsigned long long trA = sg.trial_begin + tlA, trB = sg.trial_begin + tlB;
    const uint32_t c3 = sg.stream;
    const uint32_t taps0 = sg.enc_taps[0], taps1 = sg.enc_taps[1];
    const uint32_t tm00 = 0u - (taps0 & 1u), tm01 = 0u - ((taps0 >> 1) & 1u), tm02 = 0u - ((taps0 >> 2) & 1u);
    const uint32_t tm10 = 0u - (taps1 & 1u), tm11 = 0u - ((taps1 >> 1) & 1u), tm12 = 0u - ((taps1 >> 2) & 1u);
    uint32_t prevUA = 0, prevUB = 0;
    // Philox counter words and activity masks of the two trials, pinned in registers: left to itself the
    // compiler re-derives them from blockIdx / threadIdx / the segment record before every lazy loop
    // (~50 instructions per flip word)
    uint32_t c1A = (uint32_t)trA, c2A = (uint32_t)(trA >> 32), c1B = (uint32_t)trB, c2B = (uint32_t)(trB >> 32);
    // (an inactive trial of the last block draws bits like any other; nothing of it is counted or stored)
    uint32_t c3p = c3, a_tbp = a_tb;
    asm volatile("" : "+r"(c1A), "+r"(c2A), "+r"(c1B), "+r"(c2B), "+r"(c3p), "+r"(a_tbp), "+r"(a_sq));
    const uint32_t nsb = (N + 127u) >> 7;
    for (uint32_t sb = 0; sb < nsb; ++sb) {
        uint4 UA = make_uint4(0, 0, 0, 0), UB = UA, EA0 = UA, EA1 = UA, EB0 = UA, EB1 = UA;
        if (philox) {
            UA = philox10(((4u * sb) << 6) | 32u, c1A, c2A, c3p, P);
            UB = philox10(((4u * sb) << 6) | 32u, c1B, c2B, c3p, P);
        } else {
            const uint4* base = P.bits + sg.bits_offset + (unsigned long long)sb * 3ull * ntr;
            if (actA) {
                UA = __ldg(base + tlA);
                EA0 = __ldg(base + ntr + tlA);
                EA1 = __ldg(base + 2ull * ntr + tlA);
            }
            if (actB) {
                UB = __ldg(base + tlB);
                EB0 = __ldg(base + ntr + tlB);
                EB1 = __ldg(base + 2ull * ntr + tlB);
            }
        }
        if (!sg.random_input) UA = UB = make_uint4(0, 0, 0, 0);
#pragma unroll 1
        for (int w = 0; w < 4; ++w) {
            const uint32_t t0 = sb * 128u + (uint32_t)w * 32u;
            if (t0 >= N) break;
            const uint32_t valid = min(32u, N - t0);
            const uint32_t vmask = valid == 32u ? 0xFFFFFFFFu : ((1u << valid) - 1u);
            uint32_t wev[2], wod[2];                    // received pairs of the even / odd steps (see below)
            uint32_t eA0, eA1, eB0, eB1;
            if (philox) {
                // the four flip words of this 32-step block (2 trials x 2 outputs): two calls each, stragglers by item
                uint32_t cb = (4u * sb + (uint32_t)w) << 6, vm = vmask;
                asm volatile("" : "+r"(cb), "+r"(vm));
                const FlipWords f = flip_words4(cb, c1A, c2A, c1B, c2B, c3p, vm, a_tbp, ncalls, a_sq, lane, BS, P);
                eA0 = f.e0;
                eA1 = f.e1;
                eB0 = f.e2;
                eB1 = f.e3;
            } else {
                eA0 = pick(EA0, w);
                eA1 = pick(EA1, w);
                eB0 = pick(EB0, w);
                eB1 = pick(EB1, w);
            }
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const uint32_t U = x ? UB.x : UA.x;                      // word w: the vectors rotate below
                const uint32_t e0 = x ? eB0 : eA0, e1 = x ? eB1 : eA1;
                const uint32_t pu = x ? prevUB : prevUA;
                uint32_t o0 = U & tm00, o1 = U & tm10;                  // m = 2: three tap masks per output
                const uint32_t sh1 = __funnelshift_l(pu, U, 1), sh2 = __funnelshift_l(pu, U, 2);
                o0 ^= (sh1 & tm01) ^ (sh2 & tm02);
                o1 ^= (sh1 & tm11) ^ (sh2 & tm12);
                if (x) prevUB = U; else prevUA = U;
                const uint32_t R0 = o0 ^ e0, R1 = o1 ^ e1;
                // r_t = (R0 bit t, R1 bit t) without a Morton interleave: two bit-selects put the pairs of the
                // even steps into one word and those of the odd steps into another, both at bits (t | 1, t & ~1)
                wev[x] = bitsel(R1, R0 << 1, 0x55555555u);
                wod[x] = bitsel(R1 >> 1, R0, 0x55555555u);
            }
            UA = make_uint4(UA.y, UA.z, UA.w, 0u);
            UB = make_uint4(UB.y, UB.z, UB.w, 0u);
            // steps t = 4 i + k of both trials in word q_k, field i = bits 4 i .. 4 i + 3 = (r_A, r_B): one shift puts
            // a step's pair at bits 7..10, the row offset of the branch-metric table and (bits 7..8) of trial A's log row
            uint32_t q0 = bitsel(wev[0], wev[1] << 2, 0x33333333u), q2 = bitsel(wev[0] >> 2, wev[1], 0x33333333u);
            uint32_t q1 = bitsel(wod[0], wod[1] << 2, 0x33333333u), q3 = bitsel(wod[0] >> 2, wod[1], 0x33333333u);
            if (valid == 32u) {
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {                                  // 16 steps per iteration (measured: 8 -> 6.98e11, 16 -> 7.07e11, 32 -> 6.92e11 steps/s)
                    eng.step<false, ANTI != 0>(q0 << 7, q0 << 5, P);
                    eng.step<false, ANTI != 0>(q1 << 7, q1 << 5, P);
                    eng.step<false, ANTI != 0>(q2 << 7, q2 << 5, P);
                    eng.step<false, ANTI != 0>(q3 << 7, q3 << 5, P);
                    eng.step<false, ANTI != 0>(q0 << 3, q0 << 1, P);
                    eng.step<false, ANTI != 0>(q1 << 3, q1 << 1, P);
                    eng.step<false, ANTI != 0>(q2 << 3, q2 << 1, P);
                    eng.step<false, ANTI != 0>(q3 << 3, q3 << 1, P);
                    eng.step<false, ANTI != 0>(q0 >> 1, q0 >> 3, P);
                    eng.step<false, ANTI != 0>(q1 >> 1, q1 >> 3, P);
                    eng.step<false, ANTI != 0>(q2 >> 1, q2 >> 3, P);
                    eng.step<false, ANTI != 0>(q3 >> 1, q3 >> 3, P);
                    eng.step<false, ANTI != 0>(q0 >> 5, q0 >> 7, P);
                    eng.step<false, ANTI != 0>(q1 >> 5, q1 >> 7, P);
                    eng.step<false, ANTI != 0>(q2 >> 5, q2 >> 7, P);
                    eng.step<true, ANTI != 0>(q3 >> 5, q3 >> 7, P);
                    q0 >>= 16; q1 >>= 16; q2 >>= 16; q3 >>= 16;
                }
            } else {                                                           // the last block of a trial: four steps at a time
#pragma unroll 1
                for (uint32_t c = 0; c < valid; c += 4u) {
                    const uint32_t left = valid - c;
                    eng.step<false, ANTI != 0>(q0 << 7, q0 << 5, P);
                    if (left > 1u) eng.step<false, ANTI != 0>(q1 << 7, q1 << 5, P);
                    if (left > 2u) eng.step<false, ANTI != 0>(q2 << 7, q2 << 5, P);
                    if (left > 3u) eng.step<false, ANTI != 0>(q3 << 7, q3 << 5, P);
                    q0 >>= 4; q1 >>= 4; q2 >>= 4; q3 >>= 4;
                    if ((c & 12u) == 12u) eng.normalise();
                }
                eng.normalise();
            }
        }
    }

    const bool winA = actA && (sg.decide == 0 ? (eng.a1A > eng.a0A) : (eng.a1A <= eng.a0A));
    const bool winB = actB && (sg.decide == 0 ? (eng.a1B > eng.a0B) : (eng.a1B <= eng.a0B));
    const int cnt = __syncthreads_count(winA ? 1 : 0) + __syncthreads_count(winB ? 1 : 0);
    if (threadIdx.x == 0 && cnt) {
        atomicAdd(P.tallies + seg, (unsigned long long)cnt);
        if (P.tallies2) atomicAdd(P.tallies2 + seg, (unsigned long long)cnt);
    }
    if (P.logp) {
        double2* o = reinterpret_cast<double2*>(P.logp) + sg.out_offset;
        if (actA) o[tlA] = make_double2(eng.a1A, eng.a0A);
        if (actB) o[tlB] = make_double2(eng.a1B, eng.a0B);
    }
}
