// mvd_parity.cuh -- Monte-Carlo trials of the parity-template baseline detector (SURVEY 8f N4,
// reference comp_parity.py:65-181; paper section IV).
//
// One trial (comp_parity.py:165-176): N info bits -> feed-forward encoder with an m-step zero tail
// (encode_convolutional, :65-86: n streams of N + m bits) -> BSC(p) on every bit (:171) -> fraction of
// positions t in [max_delay, N + m) at which the parity template XOR_{(j,s)} y_j[t - s] is 0
// (parity_satisfaction_fraction, :93-116) -> decide H1 iff fraction >= gamma (parity_detector, :123-132).
//
// One thread = one trial, 32 steps per iteration, everything bit-parallel: the encoder is the XOR of
// funnel-shifted info words (as in the hybrid-detector kernels), the template is the XOR of
// funnel-shifted *received* words, and the satisfied positions are a popcount.  The bit source is
// the hybrid detector's (MVD-PHILOX-2 on the device, or host-supplied 128-bit words).
// The decision compares the same float64 quotient the reference forms (satisfied / total >= gamma).
#pragma once
#include "mvd_detect2.cuh"

__global__ void __launch_bounds__(PARITY_BLOCK) parity_kernel(const __grid_constant__ Params P, const __grid_constant__ ParityBatch B,
                                                              uint32_t* __restrict__ satisfied_out) {
    __shared__ uint4 tbm[8];                      // threshold-bit masks of this segment's p (lazy_bernoulli_s)
    const ParitySeg& sg = B.s[blockIdx.y];
    const unsigned long long ntr = sg.trial_end - sg.trial_begin;
    if ((unsigned long long)blockIdx.x * PARITY_BLOCK >= ntr) return;
    if (threadIdx.x < 32u) reinterpret_cast<uint32_t*>(tbm)[threadIdx.x] = 0u - ((sg.threshold >> (31u - threadIdx.x)) & 1u);
    __syncthreads();
    const int ncalls = sg.dmin > 31u ? 0 : (int)((31u - sg.dmin) / 4u + 1u);
    const unsigned long long tl = (unsigned long long)blockIdx.x * PARITY_BLOCK + threadIdx.x;
    const bool active = tl < ntr;
    const unsigned long long trial = sg.trial_begin + tl;
    const uint32_t N = sg.N, T = sg.N + sg.m, n = sg.n;
    const bool philox = P.src_mode == MVD_SRC_PHILOX;
    const uint32_t c1 = (uint32_t)trial, c2 = (uint32_t)(trial >> 32), c3 = sg.stream;
    uint32_t prevU = 0, prevR[MVD_MAX_N] = {0, 0, 0, 0};
    uint32_t sat = 0;
    const uint32_t nsb = (T + 127u) >> 7;
    for (uint32_t sb = 0; sb < nsb; ++sb) {
        uint4 Uw = make_uint4(0, 0, 0, 0);
        uint4 Ew[MVD_MAX_N];
#pragma unroll
        for (int j = 0; j < MVD_MAX_N; ++j) Ew[j] = make_uint4(0, 0, 0, 0);
        if (philox) {
            Uw = philox10(((4u * sb) << 6) | 32u, c1, c2, c3, P);
        } else if (active) {
            const uint4* base = P.bits + sg.bits_offset + (unsigned long long)sb * (unsigned)(1 + n) * ntr + tl;
            Uw = __ldg(base);
#pragma unroll
            for (int j = 0; j < MVD_MAX_N; ++j)
                if (j < (int)n) Ew[j] = __ldg(base + (unsigned long long)(1 + j) * ntr);
        }
#pragma unroll 1
        for (int w = 0; w < 4; ++w) {
            const uint32_t t0 = sb * 128u + (uint32_t)w * 32u;
            if (t0 >= T) break;
            const uint32_t valid = min(32u, T - t0);
            const uint32_t vmask = valid == 32u ? 0xFFFFFFFFu : ((1u << valid) - 1u);
            // info bits exist for t < N only; the tail is the zero input (comp_parity.py:79-83)
            const uint32_t ninfo = t0 >= N ? 0u : min(32u, N - t0);
            const uint32_t U = pick(Uw, w) & (ninfo == 32u ? 0xFFFFFFFFu : ((1u << ninfo) - 1u));
            uint32_t par = 0;
#pragma unroll
            for (int j = 0; j < MVD_MAX_N; ++j) {
                if (j < (int)n) {
                    uint32_t E;
                    if (philox) E = lazy_bernoulli_s(((4u * sb + (uint32_t)w) << 6) | (8u * (uint32_t)j), c1, c2, c3, tbm, ncalls,
                                                     active ? vmask : 0u, P);
                    else E = pick(Ew[j], w) & vmask;
                    const uint32_t taps = sg.enc_taps[j];
                    uint32_t o = (taps & 1u) ? U : 0u;
                    for (uint32_t i = 1; i <= sg.m; ++i)
                        if ((taps >> i) & 1u) o ^= __funnelshift_l(prevU, U, i);
                    const uint32_t Rj = (o ^ E) & vmask;
                    uint32_t tm = sg.tmpl[j];
                    if (tm & 1u) par ^= Rj;
                    tm >>= 1;
                    for (uint32_t s = 1; tm; ++s, tm >>= 1)
                        if (tm & 1u) par ^= __funnelshift_l(prevR[j], Rj, s);
                    prevR[j] = Rj;
                }
            }
            prevU = U;
            // positions t0 .. t0 + valid - 1 with t >= max_delay count (comp_parity.py:107)
            uint32_t cmask = vmask;
            if (sg.max_delay > t0) cmask &= (sg.max_delay - t0 >= 32u) ? 0u : ~((1u << (sg.max_delay - t0)) - 1u);
            sat += (uint32_t)__popc(~par & cmask);
        }
    }
    const uint32_t total = T > sg.max_delay ? T - sg.max_delay : 0u;
    const double frac = total ? (double)sat / (double)total : 0.0;                 // :116
    const bool h1 = frac >= sg.gamma;                                               // :131-132
    const bool win = active && (sg.decide == 0 ? h1 : !h1);
    const int c = __syncthreads_count(win ? 1 : 0);
    if (threadIdx.x == 0 && c) atomicAdd(P.tallies + sg.seg_index, (unsigned long long)c);
    if (satisfied_out && active) satisfied_out[sg.out_offset + tl] = sat;
}
