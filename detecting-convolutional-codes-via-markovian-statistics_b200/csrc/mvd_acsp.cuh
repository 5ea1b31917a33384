// mvd_acsp.cuh -- Eq. 4-5 (viterbi_markov.py:139-159) at register-pressure scale: memories m = 4, 5, 6 (16, 32, 64
// trellis states) with TWO trials per thread and all 2^m (trial A, trial B) path-metric pairs in registers, no Markov
// state table (BASELINE config 4; the state set of m = 5 has 3.2e8 members, that of m = 6 is not enumerable).
// Output: the final metric vector D_N of every trial -- the quantity mvd_acs_hash checks against the oracle.
//
// Per step and trial pair: one VIADDMNMX.U16x2 + one add per new state (Eq. 4 for both trials), a VIMNMX3 tree and
// one subtraction per state (Eq. 5).  Branch metrics are PRMT picks from the four bytes popc(L ^ r) of each trial
// (mvd_detect3p.cuh) with the selector of every branch taken straight from the kernel parameters (constant bank), so
// no register holds a table: 342 warp-instructions per step pair at m = 6 against ~1 400 for the state-packed one-trial
// kernel, whose 64 byte-lane broadcasts and metric-key packing dominate.
#pragma once
#include "mvd_detect2.cuh"

struct AcspSel {
    uint32_t sel[128];                // PRMT selector of branch (ns, b) at [2 ns + b]
};

template <int M>
__global__ void __launch_bounds__(DET2P_BLOCK, M <= 4 ? 3 : (M == 5 ? 2 : 1)) acsp_kernel(const __grid_constant__ Params P,
                                                                                           const __grid_constant__ DevSeg sg,
                                                                                           const __grid_constant__ AcspSel S,
                                                                                           uint8_t* __restrict__ final_met) {
    constexpr int NS = 1 << M, HALF = NS / 2;
    __shared__ uint4 tbm[8];
    __shared__ uint32_t Vt[4];
    const unsigned long long ntr = sg.trial_end - sg.trial_begin;
    const uint32_t BS = blockDim.x;
    const unsigned long long blk0 = (unsigned long long)blockIdx.x * (2u * BS);
    if (blk0 >= ntr) return;
    const unsigned long long tlA = blk0 + threadIdx.x, tlB = tlA + BS;
    const bool actA = tlA < ntr, actB = tlB < ntr;
    if (threadIdx.x < 32u) reinterpret_cast<uint32_t*>(tbm)[threadIdx.x] = 0u - ((sg.threshold >> (31u - threadIdx.x)) & 1u);
    if (threadIdx.x < 4u) {                                   // V(r) = bytes popc(L ^ r), L = 0..3
        uint32_t v = 0;
        for (uint32_t L = 0; L < 4u; ++L) v |= (uint32_t)__popc(L ^ threadIdx.x) << (8u * L);
        Vt[threadIdx.x] = v;
    }
    __syncthreads();
    const uint32_t kV = (uint32_t)__cvta_generic_to_shared(Vt);

    uint32_t Q[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) Q[s] = 0u;
    // norm (warp-uniform) = false defers Eq. 5: nothing reads the metrics between steps here, so the minimum is
    // only removed every 8th step and at the end of a 32-step word (a lane grows by at most 2 per step; D_N is
    // the same vector either way)
    auto step = [&](uint32_t sA2, uint32_t sB2, bool norm) {   // r_A / r_B at bits 2..3
        const uint32_t VA = lds_u32(kV | (sA2 & 0xCu)), VB = lds_u32(kV | (sB2 & 0xCu));
        uint32_t n[NS];
#pragma unroll
        for (int ns = 0; ns < NS; ++ns)                        // Eq. 4, both trials
            n[ns] = __viaddmin_u16x2(Q[ns >> 1], __byte_perm(VA, VB, S.sel[2 * ns]),
                                     Q[(ns >> 1) + HALF] + __byte_perm(VA, VB, S.sel[2 * ns + 1]));
        uint32_t mn = 0u;
        if (norm) {
            mn = n[0];
#pragma unroll
            for (int s = 1; s + 1 < NS; s += 2) mn = __vimin3_u16x2(mn, n[s], n[s + 1]);
            mn = __vminu2(mn, n[NS - 1]);
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) Q[s] = n[s] - mn;         // Eq. 5 (mn = 0: deferred)
    };

    const uint32_t N = sg.N;
    const int ncalls = sg.dmin > 31u ? 0 : (int)((31u - sg.dmin) / 4u + 1u);
    const unsigned long long trA = sg.trial_begin + tlA, trB = sg.trial_begin + tlB;
    const uint32_t c3 = sg.stream;
    uint32_t tm0[M + 1], tm1[M + 1];
#pragma unroll
    for (int i = 0; i <= M; ++i) {
        tm0[i] = 0u - ((sg.enc_taps[0] >> i) & 1u);
        tm1[i] = 0u - ((sg.enc_taps[1] >> i) & 1u);
    }
    uint32_t prevUA = 0, prevUB = 0;
    const uint32_t nsb = (N + 127u) >> 7;
    for (uint32_t sb = 0; sb < nsb; ++sb) {
        uint4 UA = philox10(((4u * sb) << 6) | 32u, (uint32_t)trA, (uint32_t)(trA >> 32), c3, P);
        uint4 UB = philox10(((4u * sb) << 6) | 32u, (uint32_t)trB, (uint32_t)(trB >> 32), c3, P);
        if (!sg.random_input) UA = UB = make_uint4(0, 0, 0, 0);
#pragma unroll 1
        for (int w = 0; w < 4; ++w) {
            const uint32_t t0 = sb * 128u + (uint32_t)w * 32u;
            if (t0 >= N) break;
            const uint32_t valid = min(32u, N - t0);
            const uint32_t vmask = valid == 32u ? 0xFFFFFFFFu : ((1u << valid) - 1u);
            uint32_t wlo[2], whi[2];
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const uint32_t U = x ? UB.x : UA.x;
                const bool act = x ? actB : actA;
                const unsigned long long tr = x ? trB : trA;
                const uint32_t cb = (4u * sb + (uint32_t)w) << 6;
                const uint32_t e0 = lazy_bernoulli_s(cb, (uint32_t)tr, (uint32_t)(tr >> 32), c3, tbm, ncalls, act ? vmask : 0u, P);
                const uint32_t e1 = lazy_bernoulli_s(cb | 8u, (uint32_t)tr, (uint32_t)(tr >> 32), c3, tbm, ncalls, act ? vmask : 0u, P);
                const uint32_t pu = x ? prevUB : prevUA;
                uint32_t o0 = U & tm0[0], o1 = U & tm1[0];
#pragma unroll
                for (int i = 1; i <= M; ++i) {
                    const uint32_t sh = __funnelshift_l(pu, U, i);
                    o0 ^= sh & tm0[i];
                    o1 ^= sh & tm1[i];
                }
                if (x) prevUB = U; else prevUA = U;
                const uint32_t R0 = o0 ^ e0, R1 = o1 ^ e1;
                wlo[x] = (spread16(R0 & 0xFFFFu) << 1) | spread16(R1 & 0xFFFFu);
                whi[x] = (spread16(R0 >> 16) << 1) | spread16(R1 >> 16);
            }
            UA = make_uint4(UA.y, UA.z, UA.w, 0u);
            UB = make_uint4(UB.y, UB.z, UB.w, 0u);
#pragma unroll 1
            for (uint32_t t = 0; t < valid; ++t) {
                const uint32_t wa = (t & 16u) ? whi[0] : wlo[0], wb = (t & 16u) ? whi[1] : wlo[1];
                const uint32_t sh = 2u * (t & 15u);
                step((wa >> sh) << 2, (wb >> sh) << 2, M >= 6 || (t & 7u) == 7u || t + 1u == valid);   // m = 6 (one block per SM): deferring measured slower (9.75e10 -> 9.43e10)
            }
        }
    }
    // relative metrics of the last step: low halves = trial A, high halves = trial B
    if (actA) {
#pragma unroll
        for (int s = 0; s < NS; ++s) final_met[(size_t)tlA * NS + s] = (uint8_t)(Q[s] & 0xFFu);
    }
    if (actB) {
#pragma unroll
        for (int s = 0; s < NS; ++s) final_met[(size_t)tlB * NS + s] = (uint8_t)((Q[s] >> 16) & 0xFFu);
    }
}
