// mvd_detect3p.cuh -- ACS engine for memory m = 3 and m = 4 rate-1/2 codes (8 / 16 trellis states; the demo's
// second predefined pair, S = 435; m = 4: S = 25 751 ... 232 567): TWO trials per thread, the 16x2 SIMD lanes
// run across the two trials as in detect2p_kernel (mvd_detect2.cuh), and the metric-vector -> Markov-state
// lookup (the state_index dict of Pd_plotter.py:139) is a *perfect hash* built on the host for the closed set:
//
//     w0   = k0 | k1 << 16,  k_q = sum_i 16^i (D[4q+i] - D[0] + 8)          (|D[s] - D[0]| <= 7; m = 4: w1 from k2, k3)
//     slot = (((w0 c2 + w1 c4) >> 11) + disp[(w0 C1 + w1 C3) >> bshift]) & (slots - 1)        (hash, displace;
//                                                     c2, c4: odd multipliers the host build settles on)
//
// The key is OFFSET-INVARIANT (differences to D[0]), so it is the same for the un-normalised vector D' of Eq. 4 and for
// D' - min D' of Eq. 5: the minimum is not needed to find the Markov state, and Eq. 5 is applied once per 32-step block
// (it only keeps the 16-bit lanes small).  k_q is a linear form, sum_i 16^i Q[4q+i] - 4369 Q[0] + 0x8888, evaluated on
// the packed pair by IMADs (FMA pipe) with multipliers taken from the kernel parameters; both lanes' true results are in
// [0, 2^16), so the 32-bit arithmetic is exact whatever the low lane carries in between.  Round 1 packed the
// normalised metrics (base 8) and paid a VIMNMX3 tree plus a correction of 585 min per step, all on the ALU pipe.
//
// m = 3 keeps every table in shared memory.  m = 4 (template GT; S = 25 751 ... 232 567) cannot: there the log rows are
// stored BY SLOT in global memory (FastPlan::ll_slot, built once per mvd_set_loglik by slot_rows_kernel), so the slot
// itself is the row index and the slot -> state table is never read: one displacement read (shared memory when the
// bucket table fits, template DS; else L2) and one 16-byte row read per trial-step.  The first version read
// displacement, slot and row from L2 -- three dependent gathers per trial-step, L1 wavefront-bound at 74 % with
// `long_scoreboard` the top stall (profiles/r02a_m4_pair_ncu_full.txt).
//
// Results are the same bit for bit as the one-trial kernels and the oracle (same Eq. 4-5 arithmetic, same sums in
// step order).
//
// Shared memory, m = 3 (absolute addresses; alignment lets one LOP3 form an address).  The kernel is bound by shared-memory
// wavefronts (ncu: 93 % of the LSU data pipe, 56 % of the wavefronts bank conflicts with two row copies and plain hash tables,
// profiles/r03d_m3_pair_ncu_full.txt), so ONE block of 768 threads per SM shares conflict-poor replicas of every table:
//   [straggler queues 1 KB per warp][threshold masks 128 B][branch table 4 x 16 B (four banks), 256 B aligned]
//   [displacements 256 x 32 lanes x 4 B at a 32 KB-aligned address: one copy per lane = per bank, no conflicts]
//   [hash slots: slots x 16 copies x 4 B, aligned to its size: two lanes per copy]
//   [log rows S x 4 x 4 copies x 16 B, 256 B aligned]
// m = 4 keeps blocks of 256 threads: [queues][masks][branch table][displacements ph_nb x 4 B, 1 KB aligned].
#pragma once
#include "mvd_detect2.cuh"

#define PH3_C1 0x9E3779B1u
#define PH3_C3 0xC2B2AE35u

// Branch metrics without a table read per branch: for a rate-1/2 code the metric of a branch with label L
// against received word r is popc(L ^ r).
//   ANTI (every generator has its first and its last tap set, e.g. the demo pair (17,13) / (13,17)): the four branches of
//   butterfly g carry the labels X, ~X, ~X, X, so with x_g = d(X, r) the step is
//       D'[2g] = min(D[g] + x_g, D[g+H] + n - x_g),   D'[2g+1] = min(D[g] + n - x_g, D[g+H] + x_g):
//   the block stages, per received word r, the bytes x_g of all 2^(m-1) butterflies (one LDS.32 / LDS.64 per trial, rows
//   at 16-byte pitch); PRMT(row_A, row_B, constant selector) packs (x_g | r_A, x_g | r_B) -- ONE pick per butterfly instead of
//   one per branch; n - x_g of the pair is 0x00020002 - pair, and that and the two additions per butterfly are IMADs
//   (multipliers -1 / 1 from the kernel parameters, FMA pipe); the VIADDMNMX pair stays on the ALU pipe.
//   General decoders: the 2^(m+1) branch words of a step are picks from the four bytes V(r) = (d(0,r), d(1,r), d(2,r),
//   d(3,r)) with per-decoder selectors held in registers (selector nibbles with bit 3 set yield the replicated sign = 0).
// PTX prmt in its default mode: selector nibble bit 3 replicates the sign of the selected byte (0 for the small bytes used
// here).  __byte_perm with a compile-time selector is folded with bit 3 masked off, so the constant picks go through this.
__device__ __forceinline__ uint32_t prmt_s(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

template <int M, bool GT, bool DS, bool ANTI>
struct PairEngineN {
    static constexpr int NS = 1 << M, HALF = NS / 2;
    uint32_t Q[NS];                   // (trial A, trial B) metrics of the trellis states
    uint32_t sxA, sxB;                // smem: absolute address of this lane's copy of the current log row; GT: state * R
    uint32_t kV, kD, kT, tmask;       // branch table (64-byte rows) ; smem: displacement table | lane * 4 ; slot table | copy * 4 ; (slots - 1) << 6
    uint32_t sel[ANTI ? 1 : 2 * NS];  // general decoders: PRMT selector of branch (ns, b) at [2 ns + b]
    const uint32_t* gD;               // GT: displacements, slots and log rows in global memory
    const double2* gll;
    uint32_t bshift, gmask, c2, c4;   // c2, c4: multipliers of the second hash, chosen by the host build
    double a1A, a0A, a1B, a0B;
    double2 pA, pB;                   // GT: the terms read by the previous step, not yet added

    // GT: add the terms of the last step
    __device__ __forceinline__ void flush() {
        if (GT) {
            a1A += pA.x;
            a0A += pA.y;
            a1B += pB.x;
            a0B += pB.y;
        }
    }

    __device__ __forceinline__ uint32_t lookup(uint32_t w0, uint32_t w1) const {
        if (GT) {
            const uint32_t b = (w0 * PH3_C1 + w1 * PH3_C3) >> bshift;
            const uint32_t d = DS ? lds_u32(kD + (b << 2)) : __ldg(gD + b);
            return ((((w0 * c2 + w1 * c4) >> 11) + d) & gmask) << 2;           // slot * R: the row index of ll_slot
        }
        const uint32_t d6 = lds_u32(kD | (((w0 * PH3_C1) >> 17) & 0x7F80u));          // displacement * 64, this lane's copy
        return lds_u32(kT | ((((w0 * c2) >> 15) + d6) & tmask));
    }

    // (f & 0xC0) | base as ONE LOP3 (left to itself the compiler shares the AND between the two tables of a trial: three)
    static __device__ __forceinline__ uint32_t radr(uint32_t f, uint32_t base) {
        uint32_t d;
        asm("lop3.b32 %0, %1, 0xC0, %2, 0xEA;" : "=r"(d) : "r"(f), "r"(base));
        return d;
    }

    // the branch table has 16-byte rows (r at bits 4..5): its four rows lie in four different banks, where rows at the 64-byte
    // pitch of the log rows put r = 0, 2 (and 1, 3) into the same bank -- two wavefronts per read
    static __device__ __forceinline__ uint32_t xadr(uint32_t x, uint32_t base) {
        uint32_t d;
        asm("lop3.b32 %0, %1, 0x30, %2, 0xEA;" : "=r"(d) : "r"(x), "r"(base));
        return d;
    }

    // fA / fB: r_A / r_B at bits 6..7 (log rows: 64-byte entries); xA / xB: the same at bits 4..5 (branch table); other bits arbitrary.
    // Eq. 5 is deferred (see the header): the caller normalises once per 32-step block.
    __device__ __forceinline__ void step(uint32_t fA, uint32_t fB, uint32_t xA, uint32_t xB, const Params& P) {
        if (GT) {
            // the rows come from L2 (~300 cycles): the terms of a step are added one step later (same order of additions, so the
            // same sums bit for bit), which takes the read off the path of the DADD chain (long_scoreboard was the top stall)
            const double2 vA = __ldg(gll + sxA + ((fA >> 6) & 3u)), vB = __ldg(gll + sxB + ((fB >> 6) & 3u));
            a1A += pA.x;
            a0A += pA.y;
            a1B += pB.x;
            a0B += pB.y;
            pA = vA;
            pB = vB;
        } else {
            const double2 vA = lds_d2(radr(fA, sxA)), vB = lds_d2(radr(fB, sxB));
            a1A += vA.x;
            a0A += vA.y;
            a1B += vB.x;
            a0B += vB.y;
        }
        const uint32_t one = P.fp.kq[4];
        uint32_t n[NS];
        if (ANTI) {
            // row r of the branch table: bytes x_g; the packed pair (n - x_g | r_A, n - x_g | r_B) = 0x00020002 - (x_g pair), one IMAD by -1
            // (no borrow: x_g <= 2) instead of a second PRMT
            uint32_t xa[2], xb[2];
            if (M == 3) {
                xa[0] = lds_u32(xadr(xA, kV));
                xb[0] = lds_u32(xadr(xB, kV));
                xa[1] = xb[1] = 0u;
            } else {
                const uint2 ra = lds_v2(xadr(xA, kV)), rb = lds_v2(xadr(xB, kV));
                xa[0] = ra.x; xa[1] = ra.y;
                xb[0] = rb.x; xb[1] = rb.y;
            }
#pragma unroll
            for (int g = 0; g < HALF; ++g) {
                const uint32_t s4 = (uint32_t)(g & 3) | 0x80u | ((4u + (uint32_t)(g & 3)) << 8) | 0x8000u;
                const uint32_t x = prmt_s(xa[g >> 2], xb[g >> 2], s4), nx = madlo(x, P.fma_km1, 0x00020002u);
                const uint32_t t0 = madlo(Q[g], one, x), t1 = madlo(Q[g + HALF], one, x);
                n[2 * g] = __viaddmin_u16x2(Q[g + HALF], nx, t0);              // Eq. 4, both trials
                n[2 * g + 1] = __viaddmin_u16x2(Q[g], nx, t1);
            }
        } else {
            const uint32_t VA = lds_u32(xadr(xA, kV)), VB = lds_u32(xadr(xB, kV));
#pragma unroll
            for (int ns = 0; ns < NS; ++ns)             // new state ns from predecessors ns >> 1 and (ns >> 1) + HALF: Eq. 4, both trials
                n[ns] = __viaddmin_u16x2(Q[ns >> 1], prmt_s(VA, VB, sel[2 * ns]), Q[(ns >> 1) + HALF] + prmt_s(VA, VB, sel[2 * ns + 1]));
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) Q[s] = n[s];
        // offset-invariant key words: low halves = trial A, high halves = trial B
        uint32_t k[NS / 4];
#pragma unroll
        for (int q = 0; q < NS / 4; ++q) {
            uint32_t acc = madlo(Q[0], q == 0 ? P.fp.kq[5] : P.fp.kq[3], 0x88888888u);      // - 4369 D[0] (q = 0: - 4368 D[0]) + 8 per digit
            if (q > 0) acc = madlo(Q[4 * q], one, acc);
            acc = madlo(Q[4 * q + 1], P.fp.kq[0], acc);
            acc = madlo(Q[4 * q + 2], P.fp.kq[1], acc);
            k[q] = madlo(Q[4 * q + 3], P.fp.kq[2], acc);
        }
        sxA = lookup(__byte_perm(k[0], k[1], 0x5410), NS == 16 ? __byte_perm(k[NS / 4 - 2], k[NS / 4 - 1], 0x5410) : 0u);
        sxB = lookup(__byte_perm(k[0], k[1], 0x7632), NS == 16 ? __byte_perm(k[NS / 4 - 2], k[NS / 4 - 1], 0x7632) : 0u);
    }

    // Eq. 5 (keeps the lanes small; the key does not see it)
    __device__ __forceinline__ void normalise() {
        uint32_t mn = __vimin3_u16x2(__vimin3_u16x2(Q[0], Q[1], Q[2]), __vimin3_u16x2(Q[3], Q[4], Q[5]), __vminu2(Q[6], Q[7]));
        if (NS == 16) mn = __vimin3_u16x2(mn, __vimin3_u16x2(Q[8], Q[9], Q[10]), __vimin3_u16x2(__vimin3_u16x2(Q[11], Q[12], Q[13]), Q[14], Q[15]));
#pragma unroll
        for (int s = 0; s < NS; ++s) Q[s] -= mn;
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

template <int M, bool GT, int PHILOX, bool DS, bool ANTI>
__global__ void __launch_bounds__(M == 3 ? DET3P_BLOCK : DET2P_BLOCK, M == 3 ? 1 : 2) detect3p_kernel(const __grid_constant__ Params P,
                                                                                                      const __grid_constant__ SegBatch B) {
    constexpr int NS = 1 << M, HALF = NS / 2;
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
    const uint32_t slots = P.fp.ph_slots, tbytes = GT ? 0u : slots * 64u;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
    uint32_t a_sq = sbase + (threadIdx.x >> 5) * 1024u;      // this warp's straggler queue (flip_words4)
    const uint32_t a_tb = (sbase + (GT ? (uint32_t)DET2P_QUEUES : (uint32_t)DET3P_QUEUES) + 255u) & ~255u;
    const uint32_t a_V = a_tb + 256u;                        // branch table: 4 rows x 16 B, 256 B aligned
    // m = 3: 256 displacements x 32 lanes (32 KB, 32 KB aligned); m = 4 (DS): ph_nb of them (1 KB aligned)
    const uint32_t a_D = GT ? ((a_V + 256u + 1023u) & ~1023u) : ((a_V + 256u + 32767u) & ~32767u);
    const uint32_t a_T = GT ? 0u : ((a_D + 32768u + tbytes - 1u) & ~(tbytes - 1u));
    const uint32_t a_ll = a_T + tbytes;                      // log rows (m = 3), 256 B aligned
    unsigned char* g = smem_raw - sbase;                     // generic pointer of shared address 0
    {
        uint32_t dyn;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        const uint32_t need = GT ? (DS ? a_D + 4u * P.fp.ph_nb : a_V + 256u) : a_ll + (SR << 6);
        if (need - sbase > dyn) {                             // the host sized the window for another base address
            if (threadIdx.x == 0) atomicOr(P.error_flag, 4);
            return;
        }
    }

    if (threadIdx.x < 32u)
        *reinterpret_cast<uint32_t*>(g + a_tb + 4u * threadIdx.x) = 0u - ((sg.threshold >> (31u - threadIdx.x)) & 1u);
    if (!GT) {
        const double2* llg = P.ll + (size_t)sg.table * SR;
        for (uint32_t i = threadIdx.x; i < SR; i += BS) {   // one global read per entry, four shared copies
            const double2 v = __ldg(llg + i);
#pragma unroll
            for (uint32_t c = 0; c < 4u; ++c) *reinterpret_cast<double2*>(g + a_ll + (i << 6) + (c << 4)) = v;
        }
    }
    if (threadIdx.x < 4u) {
        const uint32_t r = threadIdx.x;
        uint32_t* row = reinterpret_cast<uint32_t*>(g + a_V + 16u * r);
        if (ANTI) {
            // bytes x_g = d(g -> 2g | r) of the butterflies g = 0 .. 2^(m-1) - 1 (P.bm[r][2 g + b]: low half = d(g + HALF b -> 2g))
            uint32_t x[2] = {0u, 0u};
            for (uint32_t gg = 0; gg < (uint32_t)HALF; ++gg) x[gg >> 2] |= (P.bm[r * NS + 2u * gg] & 0xFFFFu) << (8u * (gg & 3u));
            row[0] = x[0];
            row[1] = x[1];
        } else {                                             // V(r) = bytes popc(L ^ r), L = 0..3
            uint32_t v = 0;
            for (uint32_t L = 0; L < 4u; ++L) v |= (uint32_t)__popc(L ^ r) << (8u * L);
            row[0] = v;
        }
    }
    if (GT && DS)                                            // bucket displacements
        for (uint32_t i = threadIdx.x; i < P.fp.ph_nb; i += BS) *reinterpret_cast<uint32_t*>(g + a_D + 4u * i) = P.fp.ph_d[i];
    if (!GT) {
        for (uint32_t i = threadIdx.x; i < 256u * 4u; i += BS) {              // a quarter of a bucket's 32 lane copies per thread
            const uint4 d = make_uint4(P.fp.ph_d[i >> 2] << 6, P.fp.ph_d[i >> 2] << 6, P.fp.ph_d[i >> 2] << 6, P.fp.ph_d[i >> 2] << 6);
            *reinterpret_cast<uint4*>(g + a_D + 32u * i) = d;
            *reinterpret_cast<uint4*>(g + a_D + 32u * i + 16u) = d;
        }
        for (uint32_t i = threadIdx.x; i < slots * 4u; i += BS) {             // four of a slot's 16 copies per thread
            const uint32_t row = P.fp.ph_t[i >> 2];          // state * 4, or MVD_EMPTY (never looked up)
            const uint32_t a = a_ll + ((row == MVD_EMPTY ? 0u : row) << 6);
            *reinterpret_cast<uint4*>(g + a_T + 16u * i) = make_uint4(a, a + 16u, a + 32u, a + 48u);
        }
    }
    __syncthreads();

    PairEngineN<M, GT, DS, ANTI> eng;
#pragma unroll
    for (int s = 0; s < NS; ++s) eng.Q[s] = 0u;
    // state 0 = the all-zero vector
    eng.sxA = eng.sxB = GT ? P.fp.ph_slot0 << 2 : a_ll + ((lane & 3u) << 4);
    eng.gD = P.fp.ph_d;
    eng.gll = GT ? P.fp.ll_slot + (size_t)sg.table * slots * 4u : P.ll + (size_t)sg.table * SR;
    eng.bshift = P.fp.ph_bshift;
    eng.gmask = slots - 1u;
    eng.c2 = P.fp.ph_c2;
    eng.c4 = P.fp.ph_c4;
    eng.kV = a_V;
    eng.kD = GT ? a_D : a_D + (lane << 2);
    if (!ANTI) {
        // label of branch (ns, b) from the branch-metric table of the decoder (P.bm[r][2 g + b] = distances to ns = 2g and
        // 2g + 1 from predecessor g + HALF b): (d(L,0), d(L,1)) = (0,1), (1,0), (1,2), (2,1) for L = 0, 1, 2, 3
#pragma unroll
        for (int ns = 0; ns < NS; ++ns)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const uint32_t w0 = P.bm[0 * NS + 2 * (ns >> 1) + b], w1 = P.bm[1 * NS + 2 * (ns >> 1) + b];
                const uint32_t d0 = (ns & 1) ? (w0 >> 16) : (w0 & 0xFFFFu), d1 = (ns & 1) ? (w1 >> 16) : (w1 & 0xFFFFu);
                const uint32_t L = d0 == 0u ? 0u : (d0 == 2u ? 3u : (d1 == 0u ? 1u : 2u));
                eng.sel[ANTI ? 0 : 2 * ns + b] = L | 0x80u | ((4u + L) << 8) | 0x8000u;
            }
    }
    eng.kT = a_T + ((lane & 15u) << 2);
    eng.tmask = (slots - 1u) << 6;
    eng.a1A = eng.a0A = eng.a1B = eng.a0B = 0.0;
    eng.pA = eng.pB = make_double2(0.0, 0.0);

    const uint32_t N = sg.N;
    const int ncalls = sg.dmin > 31u ? 0 : (int)((31u - sg.dmin) / 4u + 1u);
    constexpr bool philox = PHILOX != 0;              // bit source as a template parameter (see detect2p_kernel)
    const unsigned long long trA = sg.trial_begin + tlA, trB = sg.trial_begin + tlB;
    // Philox counter words, stream id and mask-table address pinned in registers (see detect2p_kernel); an inactive trial of
    // the last block draws bits like any other, nothing of it is counted or stored
    uint32_t c1A = (uint32_t)trA, c2A = (uint32_t)(trA >> 32), c1B = (uint32_t)trB, c2B = (uint32_t)(trB >> 32);
    uint32_t c3 = sg.stream, a_tbp = a_tb;
    asm volatile("" : "+r"(c1A), "+r"(c2A), "+r"(c1B), "+r"(c2B), "+r"(c3), "+r"(a_tbp), "+r"(a_sq));
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
                // the four flip words of this 32-step block (2 trials x 2 outputs): two calls each by everybody, the
                // undecided rest through the warp's straggler queue (flip_words4, mvd_detect2.cuh)
                uint32_t cb = (4u * sb + (uint32_t)w) << 6, vm = vmask;
                asm volatile("" : "+r"(cb), "+r"(vm));
                const FlipWords f = flip_words4(cb, c1A, c2A, c1B, c2B, c3, vm, a_tbp, ncalls, a_sq, lane, BS, P);
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
                // (the branch-table field of a step is the log-row field of the step after next: one more shift per word and oct)
                const uint32_t ea6 = ea << 6, eb6 = eb << 6, oa6 = oa << 6, ob6 = ob << 6, ea4 = ea << 4, eb4 = eb << 4, oa4 = oa << 4, ob4 = ob << 4;
                const uint32_t ea2 = ea << 2, eb2 = eb << 2, oa2 = oa << 2, ob2 = ob << 2;
                eng.step(ea6, eb6, ea4, eb4, P);
                eng.step(oa6, ob6, oa4, ob4, P);
                eng.step(ea4, eb4, ea2, eb2, P);
                eng.step(oa4, ob4, oa2, ob2, P);
                eng.step(ea2, eb2, ea, eb, P);
                eng.step(oa2, ob2, oa, ob, P);
                eng.step(ea, eb, ea >> 2, eb >> 2, P);
                eng.step(oa, ob, oa >> 2, ob >> 2, P);
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
                        eng.step(ra << 6, rb << 6, ra << 4, rb << 4, P);
                    }
                }
            }
            eng.normalise();                            // Eq. 5 once per block: a lane grows by at most n = 2 per step
        }
    }

    eng.flush();
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
