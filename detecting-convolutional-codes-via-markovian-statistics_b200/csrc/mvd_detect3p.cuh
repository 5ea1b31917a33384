// mvd_detect3p.cuh -- ACS engine for memory m = 3 and m = 4 rate-1/2 codes (8 / 16 trellis states; the demo's
// second predefined pair, S = 435; m = 4: S = 25 751 ... 232 567): TWO trials per thread, the 16x2 SIMD lanes
// run across the two trials as in detect2p_kernel (mvd_detect2.cuh), and the metric-vector -> Markov-state
// lookup (the state_index dict of Pd_plotter.py:139) is a *perfect hash* built on the host for the closed set:
//
//     w0   = k0 | k1 << 16,  k_i = D[4i] + 8 D[4i+1] + 64 D[4i+2] + 512 D[4i+3]   (metrics <= 7; m = 4: w1 from k2, k3)
//     slot = (((w0 c2 + w1 c4) >> 11) + disp[(w0 C1 + w1 C3) >> bshift]) & (slots - 1)        (hash, displace;
//                                                     c2, c4: odd multipliers the host build settles on)
//
// m = 3 keeps every table in shared memory.  m = 4 (template GT; S = 25 751 ... 232 567) cannot: there the log rows are
// stored BY SLOT in global memory (FastPlan::ll_slot, built once per mvd_set_loglik by slot_rows_kernel), so the slot
// itself is the row index and the slot -> state table is never read: one displacement read (shared memory when the
// bucket table fits, template DS; else L2) and one 16-byte row read per trial-step.  The first version read
// displacement, slot and row from L2 -- three dependent gathers per trial-step, L1 wavefront-bound at 74 % with
// `long_scoreboard` the top stall (profiles/r02a_m4_pair_ncu_full.txt).
//
// two dependent shared-memory reads and no probe loop, against ~1.3 probes of an open-addressing table with
// key compares and a divergent loop in detect2_kernel<LK_HASH, 3> (66 warp-instructions per trellis step instead of
// 107 at m = 3).  Results are the same bit for bit (same Eq. 4-5 arithmetic, same sums in step order).
//
// Shared memory (absolute addresses; alignment lets one LOP3 form an address):
//   [hash slots: slots x 2 copies x 4 B, aligned to its size][displacements 256 x 4 B, 1 KB aligned]
//   [threshold masks 128 B][V(r) 16 B (+ pad)][log rows S x 4 x 2 copies x 16 B, 128 B aligned]
#pragma once
#include "mvd_detect2.cuh"

#define PH3_C1 0x9E3779B1u
#define PH3_C3 0xC2B2AE35u

// Branch metrics without a table read per branch: for a rate-1/2 code the metric of a branch with label L
// against received word r is popc(L ^ r), so the 16 branch words of a step are 16 picks from the four
// bytes V(r) = (d(0,r), d(1,r), d(2,r), d(3,r)).  One 4-byte read per trial gives V(r_A), V(r_B); branch
// (ns, b) is PRMT(V_A, V_B, sel[ns][b]) = V_A[L] | V_B[L] << 16 with a kernel-constant selector (selector
// nibbles with bit 3 set yield the replicated sign bit = 0).  16 PRMTs replace four LDS.128 (16 shared-memory
// wavefronts per step pair): this kernel is shared-memory bound, the ALU pipe has room.
template <int M, bool GT, bool DS = false>
struct PairEngineN {
    static constexpr int NS = 1 << M, HALF = NS / 2;
    uint32_t Q[NS];                   // (trial A, trial B) metrics of the trellis states
    uint32_t sxA, sxB;                // smem: absolute address of this lane's copy of the current log row; GT: state * R
    uint32_t kV, kD, kT, tmask;       // V(r) table (16 B) ; smem: displacement table ; slot table | copy * 4 ; (slots - 1) << 3
    uint32_t sel[2 * NS];             // PRMT selector of branch (ns, b) at [2 ns + b]
    const uint32_t* gD;               // GT: displacements, slots and log rows in global memory
    const uint32_t* gT;
    const double2* gll;
    uint32_t bshift, gmask, c2, c4;   // c2, c4: multipliers of the second hash, chosen by the host build
    double a1A, a0A, a1B, a0B;

    __device__ __forceinline__ uint32_t lookup(uint32_t w0, uint32_t w1) const {
        if (GT) {
            const uint32_t b = (w0 * PH3_C1 + w1 * PH3_C3) >> bshift;
            const uint32_t d = DS ? lds_u32(kD + (b << 2)) : __ldg(gD + b);
            return ((((w0 * c2 + w1 * c4) >> 11) + d) & gmask) << 2;           // slot * R: the row index of ll_slot
        }
        const uint32_t d8 = lds_u32(kD | (((w0 * PH3_C1) >> 22) & 0x3FCu));           // displacement * 8
        return lds_u32(kT | ((((w0 * c2) >> 18) + d8) & tmask));
    }

    // sA5 / sB5: r_A / r_B at bits 5..6 (log rows, 32-byte entries); sA2 / sB2: r_A / r_B at bits 2..3 (V table).
    // NORM = false defers Eq. 5 (as PairEngine::step of the m = 2 kernel): the key words are corrected by
    // 585 min (585 = 1 + 8 + 64 + 512) instead of subtracting the minimum from all 2^m metrics; the last step
    // of every 8-step stretch normalises, so a lane grows by at most 16 in between.
    template <bool NORM>
    __device__ __forceinline__ void step(uint32_t sA5, uint32_t sB5, uint32_t sA2, uint32_t sB2) {
        const double2 vA = GT ? __ldg(gll + sxA + ((sA5 >> 5) & 3u)) : lds_d2(sxA | (sA5 & 0x60u));
        const double2 vB = GT ? __ldg(gll + sxB + ((sB5 >> 5) & 3u)) : lds_d2(sxB | (sB5 & 0x60u));
        a1A += vA.x;
        a0A += vA.y;
        a1B += vB.x;
        a0B += vB.y;
        const uint32_t VA = lds_u32(kV | (sA2 & 0xCu)), VB = lds_u32(kV | (sB2 & 0xCu));
        uint32_t n[NS];
#pragma unroll
        for (int ns = 0; ns < NS; ++ns)                 // new state ns from predecessors ns >> 1 and (ns >> 1) + HALF: Eq. 4, both trials
            n[ns] = __viaddmin_u16x2(Q[ns >> 1], __byte_perm(VA, VB, sel[2 * ns]), Q[(ns >> 1) + HALF] + __byte_perm(VA, VB, sel[2 * ns + 1]));
        uint32_t mn = __vimin3_u16x2(__vimin3_u16x2(n[0], n[1], n[2]), __vimin3_u16x2(n[3], n[4], n[5]), __vminu2(n[6], n[7]));
        if (NS == 16) mn = __vimin3_u16x2(mn, __vimin3_u16x2(n[8], n[9], n[10]), __vimin3_u16x2(__vimin3_u16x2(n[11], n[12], n[13]), n[14], n[15]));
#pragma unroll
        for (int s = 0; s < NS; ++s) Q[s] = NORM ? n[s] - mn : n[s];            // Eq. 5 (or deferred)
        uint32_t k[NS / 4];
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) {
            k[i] = ((Q[4 * i + 3] * 8u + Q[4 * i + 2]) * 8u + Q[4 * i + 1]) * 8u + Q[4 * i];
            if (!NORM) k[i] -= 585u * mn;
        }
        // low halves = trial A, high halves = trial B
        sxA = lookup(__byte_perm(k[0], k[1], 0x5410), NS == 16 ? __byte_perm(k[NS / 4 - 2], k[NS / 4 - 1], 0x5410) : 0u);
        sxB = lookup(__byte_perm(k[0], k[1], 0x7632), NS == 16 ? __byte_perm(k[NS / 4 - 2], k[NS / 4 - 1], 0x7632) : 0u);
    }
};

// ll_slot[t][slot * 4 + r] = ll[t][state(slot) * 4 + r]: the log rows of every table in hash-slot order
__global__ void slot_rows_kernel(const double2* __restrict__ ll, const uint32_t* __restrict__ pht, uint32_t slots, uint32_t SR,
                                 uint32_t ntables, double2* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)slots * 4u) return;
    const uint32_t st = pht[idx >> 2];                       // state * R, or MVD_EMPTY (never looked up)
    for (uint32_t t = 0; t < ntables; ++t)
        out[(size_t)t * slots * 4u + idx] = st == MVD_EMPTY ? make_double2(0.0, 0.0) : ll[(size_t)t * SR + st + (idx & 3u)];
}

// Device-side packing of the log-likelihood tables (mvd_set_loglik): the host uploads log P1 [ntables][SR] and
// log Tref [SR] as they are; the interleaved {log P1, log Tref} rows and (large S) the 16-byte entries of the one-load
// NEXT walk {log P1, next row byte offset, c} are written here instead of by single-threaded host loops.
__global__ void pack_ll_kernel(const double* __restrict__ lp1, const double* __restrict__ ltref, const uint32_t* __restrict__ nxt,
                               const uint32_t* __restrict__ tcode, uint32_t SR, uint32_t ntables, double2* __restrict__ ll,
                               uint4* __restrict__ gfsm1) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)SR * ntables) return;
    const uint32_t e = (uint32_t)(idx % SR);
    const double a = lp1[idx];
    ll[idx] = make_double2(a, ltref[e]);
    if (gfsm1)
        gfsm1[idx] = make_uint4((uint32_t)__double2loint(a), (uint32_t)__double2hiint(a), nxt[e] << 4, tcode[e]);
}

template <int M, bool GT, int PHILOX, bool DS = false>
__global__ void __launch_bounds__(DET2P_BLOCK, M == 3 ? 3 : 2) detect3p_kernel(const __grid_constant__ Params P,
                                                                               const __grid_constant__ SegBatch B) {
    constexpr int NS = 1 << M;
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
    const uint32_t slots = P.fp.ph_slots, tbytes = GT ? 0u : slots * 8u;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t a_T = GT ? ((sbase + 1023u) & ~1023u) : ((sbase + tbytes - 1u) & ~(tbytes - 1u));
    const uint32_t a_D = a_T + tbytes;                       // tbytes >= 2 KB keeps the 1 KB alignment
    const uint32_t a_tb = a_D + 1024u;
    const uint32_t a_V = a_tb + 128u;                        // 16-byte table V(r), 16 B aligned
    const uint32_t a_ll = a_V + 128u;
    unsigned char* g = smem_raw - sbase;                     // generic pointer of shared address 0

    if (threadIdx.x < 32u)
        *reinterpret_cast<uint32_t*>(g + a_tb + 4u * threadIdx.x) = 0u - ((sg.threshold >> (31u - threadIdx.x)) & 1u);
    if (!GT) {
        const double2* llg = P.ll + (size_t)sg.table * SR;
        for (uint32_t i = threadIdx.x; i < SR * 2u; i += BS)
            *reinterpret_cast<double2*>(g + a_ll + ((i >> 1) << 5) + ((i & 1u) << 4)) = __ldg(llg + (i >> 1));
    }
    if (threadIdx.x < 4u) {                                  // V(r) = bytes popc(L ^ r), L = 0..3
        const uint32_t r = threadIdx.x;
        uint32_t v = 0;
        for (uint32_t L = 0; L < 4u; ++L) v |= (uint32_t)__popc(L ^ r) << (8u * L);
        *reinterpret_cast<uint32_t*>(g + a_V + 4u * r) = v;
    }
    if (GT && DS)                                            // bucket displacements: ph_nb x 4 B after the V table
        for (uint32_t i = threadIdx.x; i < P.fp.ph_nb; i += BS) *reinterpret_cast<uint32_t*>(g + a_ll + 4u * i) = P.fp.ph_d[i];
    if (!GT) {
        for (uint32_t i = threadIdx.x; i < 256u; i += BS) *reinterpret_cast<uint32_t*>(g + a_D + 4u * i) = P.fp.ph_d[i] << 3;
        for (uint32_t i = threadIdx.x; i < slots * 2u; i += BS) {
            const uint32_t row = P.fp.ph_t[i >> 1];          // state * 4, or MVD_EMPTY (never looked up)
            *reinterpret_cast<uint32_t*>(g + a_T + 4u * i) = a_ll + ((row == MVD_EMPTY ? 0u : row) << 5) + ((i & 1u) << 4);
        }
    }
    __syncthreads();

    PairEngineN<M, GT, DS> eng;
#pragma unroll
    for (int s = 0; s < NS; ++s) eng.Q[s] = 0u;
    eng.sxA = eng.sxB = GT ? P.fp.ph_slot0 << 2 : a_ll + ((lane & 1u) << 4); // state 0 = the all-zero vector
    eng.gD = P.fp.ph_d;
    eng.gT = P.fp.ph_t;
    eng.gll = GT ? P.fp.ll_slot + (size_t)sg.table * slots * 4u : P.ll + (size_t)sg.table * SR;
    eng.bshift = P.fp.ph_bshift;
    eng.gmask = slots - 1u;
    eng.c2 = P.fp.ph_c2;
    eng.c4 = P.fp.ph_c4;
    eng.kV = a_V;
    eng.kD = (GT && DS) ? a_ll : a_D;
    // label of branch (ns, b) from the branch-metric table of the decoder (P.bm[r][2 g + b] = distances to ns = 2g and
    // 2g + 1 from predecessor g + 4 b): (d(L,0), d(L,1)) = (0,1), (1,0), (1,2), (2,1) for L = 0, 1, 2, 3
#pragma unroll
    for (int ns = 0; ns < NS; ++ns)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const uint32_t w0 = P.bm[0 * NS + 2 * (ns >> 1) + b], w1 = P.bm[1 * NS + 2 * (ns >> 1) + b];
            const uint32_t d0 = (ns & 1) ? (w0 >> 16) : (w0 & 0xFFFFu), d1 = (ns & 1) ? (w1 >> 16) : (w1 & 0xFFFFu);
            const uint32_t L = d0 == 0u ? 0u : (d0 == 2u ? 3u : (d1 == 0u ? 1u : 2u));
            eng.sel[2 * ns + b] = L | 0x80u | ((4u + L) << 8) | 0x8000u;
        }
    eng.kT = a_T + ((lane & 1u) << 2);
    eng.tmask = (slots - 1u) << 3;
    eng.a1A = eng.a0A = eng.a1B = eng.a0B = 0.0;

    const uint32_t N = sg.N;
    const int ncalls = sg.dmin > 31u ? 0 : (int)((31u - sg.dmin) / 4u + 1u);
    constexpr bool philox = PHILOX != 0;              // bit source as a template parameter (see detect2p_kernel)
    const unsigned long long trA = sg.trial_begin + tlA, trB = sg.trial_begin + tlB;
    // Philox counter words, activity masks, stream id and mask-table address pinned in registers (see detect2p_kernel)
    uint32_t c1A = (uint32_t)trA, c2A = (uint32_t)(trA >> 32), c1B = (uint32_t)trB, c2B = (uint32_t)(trB >> 32);
    uint32_t mA = actA ? 0xFFFFFFFFu : 0u, mB = actB ? 0xFFFFFFFFu : 0u, c3 = sg.stream, a_tbp = a_tb;
    asm volatile("" : "+r"(c1A), "+r"(c2A), "+r"(c1B), "+r"(c2B), "+r"(mA), "+r"(mB), "+r"(c3), "+r"(a_tbp));
    const uint32_t taps0 = sg.enc_taps[0], taps1 = sg.enc_taps[1];
    uint32_t tm0[M + 1], tm1[M + 1];
#pragma unroll
    for (int i = 0; i <= M; ++i) {
        tm0[i] = 0u - ((taps0 >> i) & 1u);
        tm1[i] = 0u - ((taps1 >> i) & 1u);
    }
    uint32_t prevUA = 0, prevUB = 0;
    const uint32_t nsb = (N + 127u) >> 7;
    for (uint32_t sb = 0; sb < nsb; ++sb) {
        uint4 UA = make_uint4(0, 0, 0, 0), UB = UA, EA0 = UA, EA1 = UA, EB0 = UA, EB1 = UA;
        if (philox) {
            UA = philox10(((4u * sb) << 6) | 32u, c1A, c2A, c3, P);
            UB = philox10(((4u * sb) << 6) | 32u, c1B, c2B, c3, P);
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
            uint32_t wev[2], wod[2];                    // received pairs of the even / odd steps
            uint32_t eA0, eA1, eB0, eB1;
            if (philox) {
                // four flip words (2 trials x 2 outputs) through one copy of the lazy loop, results rotate
                uint32_t cb = (4u * sb + (uint32_t)w) << 6, vm = vmask;
                asm volatile("" : "+r"(cb), "+r"(vm));
                eA0 = eA1 = eB0 = eB1 = 0u;
#pragma unroll 1
                for (int j = 0; j < 4; ++j) {
                    const bool second = j >= 2;
                    const uint32_t e = lazy_bernoulli_a(cb | (((uint32_t)j & 1u) << 3), second ? c1B : c1A, second ? c2B : c2A, c3, a_tbp,
                                                        ncalls, (second ? mB : mA) & vm, P);
                    eA0 = eA1;
                    eA1 = eB0;
                    eB0 = eB1;
                    eB1 = e;
                }
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
                uint32_t o0 = U & tm0[0], o1 = U & tm1[0];              // m + 1 tap masks per output
#pragma unroll
                for (int i = 1; i <= M; ++i) {
                    const uint32_t sh = __funnelshift_l(pu, U, i);
                    o0 ^= sh & tm0[i];
                    o1 ^= sh & tm1[i];
                }
                if (x) prevUB = U; else prevUA = U;
                const uint32_t R0 = o0 ^ e0, R1 = o1 ^ e1;
                // r_t = (R0 bit t, R1 bit t): two bit-selects instead of a Morton interleave (see detect2p_kernel)
                wev[x] = bitsel(R1, R0 << 1, 0x55555555u);
                wod[x] = bitsel(R1 >> 1, R0, 0x55555555u);
            }
            UA = make_uint4(UA.y, UA.z, UA.w, 0u);
            UB = make_uint4(UB.y, UB.z, UB.w, 0u);
            // 8 steps = bits 0..7 of the even-step words (ea, eb) and of the odd-step words (oa, ob)
            auto oct = [&](uint32_t ea, uint32_t oa, uint32_t eb, uint32_t ob) {
                eng.step<false>(ea << 5, eb << 5, ea << 2, eb << 2);
                eng.step<false>(oa << 5, ob << 5, oa << 2, ob << 2);
                eng.step<false>(ea << 3, eb << 3, ea, eb);
                eng.step<false>(oa << 3, ob << 3, oa, ob);
                eng.step<false>(ea << 1, eb << 1, ea >> 2, eb >> 2);
                eng.step<false>(oa << 1, ob << 1, oa >> 2, ob >> 2);
                eng.step<false>(ea >> 1, eb >> 1, ea >> 4, eb >> 4);
                eng.step<true>(oa >> 1, ob >> 1, oa >> 4, ob >> 4);
            };
#pragma unroll 1
            for (uint32_t c = 0; c < valid; c += 8u) {
                const uint32_t ea = wev[0] >> c, oa = wod[0] >> c, eb = wev[1] >> c, ob = wod[1] >> c;
                if (c + 8u <= valid) {
                    oct(ea, oa, eb, ob);
                } else {
                    for (uint32_t j = 0; j < valid - c; ++j) {
                        const uint32_t sh = j & ~1u;
                        const uint32_t ra = (((j & 1u) ? oa : ea) >> sh) & 3u, rb = (((j & 1u) ? ob : eb) >> sh) & 3u;
                        eng.step<true>(ra << 5, rb << 5, ra << 2, rb << 2);
                    }
                }
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
