// mvd_kernels.cuh -- sm_100a kernels of the hybrid-detector hot path.
//
// One thread = one Monte-Carlo trial (or one learning chain).  Per 32 trellis steps a thread
//   1. obtains 32 info bits and n x 32 BSC flips  (on-device Philox4x32-10 "MVD-PHILOX-2", lazily evaluated
//      32-lane Bernoulli words -- or coalesced 128-bit loads of host-supplied bitstreams),
//   2. encodes the 32 steps bit-parallel (XOR of funnel-shifted info words = GF(2) convolution,
//      the bit-sliced form of viterbi_markov.py:82-106 for k = 1),
//   3. walks the 32 steps through one of two engines
//        ACS : Eq. 4-5 (viterbi_markov.py:139-159) on 16x2-packed path metrics held in
//              registers (VIADDMNMX.U16x2 / VIMNMX3.U16x2), then metric-vector -> Markov-state
//              lookup through a shared-memory hash table (the state_index dict of
//              Pd_plotter.py:139),
//        FSM : the same chain walked through the precomputed NEXT[state][r] table,
//      and accumulates either two float64 log-likelihood sums in step order
//      (Pd_plotter.py:106-116), a transition histogram (Pd_plotter.py:158-163) or a trace.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mvd.h"
#include "mvd_types.h"

// ------------------------------------------------------------------------------------------ Philox
__device__ __forceinline__ uint4 philox10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const Params& P) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ P.rk0[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ P.rk1[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

// 32 Bernoulli(T / 2^32) lanes at once: MSB-first comparison of a lazily drawn uniform with T
// (MVD-PHILOX-2: call k of output j of 32-step block b has counter word c0 = (b << 6) | (8 j + k)).
// The loop is warp-uniform (vote); calls are addressed by position, so skipping one for a warp whose
// lanes are all decided changes nothing else.
__device__ __forceinline__ uint32_t lazy_bernoulli(uint32_t c0base, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t T,
                                                   int dmin, uint32_t vmask, const Params& P) {
    uint32_t und = vmask, e = 0;
    int d = 31;
    uint32_t k = 0;
    while (d >= dmin) {
        if (!__any_sync(0xFFFFFFFFu, und != 0u)) break;
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

__device__ __forceinline__ uint32_t pick(const uint4& v, int w) {
    return w == 0 ? v.x : (w == 1 ? v.y : (w == 2 ? v.z : v.w));
}

// Find the segment a block belongs to (block_begin is ascending).
__device__ __forceinline__ uint32_t find_segment(const Params& P, uint32_t blk) {
    uint32_t lo = 0, hi = P.nsegs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (P.segs[mid].block_begin <= blk) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------ driver
// Feeds an engine with (absolute step index, received word) for every step of one trial.
template <int NOUT, class Eng>
__device__ __forceinline__ void run_trial(const Params& P, const DevSeg& sg, bool active, unsigned long long trial,
                                          unsigned long long tl, unsigned long long ntr, Eng& eng) {
    const int n = NOUT ? NOUT : P.n;
    const int m = P.m;
    const uint32_t N = sg.N;
    const uint32_t T = sg.threshold;
    const int dmin = (int)sg.dmin;
    const bool philox = P.src_mode == MVD_SRC_PHILOX;
    const uint32_t c1 = (uint32_t)trial, c2 = (uint32_t)(trial >> 32), c3 = sg.stream;
    uint32_t prevU = 0;
    const uint32_t nsb = (N + 127u) >> 7;
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
                if (j < n) Ew[j] = __ldg(base + (unsigned long long)(1 + j) * ntr);
        }
        if (!sg.random_input) Uw = make_uint4(0, 0, 0, 0);
#pragma unroll 1
        for (int w = 0; w < 4; ++w) {
            const uint32_t t0 = sb * 128u + (uint32_t)w * 32u;
            if (t0 >= N) break;
            const uint32_t valid = min(32u, N - t0);
            const uint32_t vmask = valid == 32u ? 0xFFFFFFFFu : ((1u << valid) - 1u);
            const uint32_t U = pick(Uw, w);
            uint32_t Rw[MVD_MAX_N];
#pragma unroll
            for (int j = 0; j < MVD_MAX_N; ++j) {
                Rw[j] = 0;
                if (j < n) {
                    uint32_t E;
                    if (philox) E = lazy_bernoulli(((4u * sb + (uint32_t)w) << 6) | (8u * (uint32_t)j), c1, c2, c3, T, dmin,
                                                   active ? vmask : 0u, P);
                    else E = pick(Ew[j], w);
                    const uint32_t taps = sg.enc_taps[j];
                    uint32_t o = (taps & 1u) ? U : 0u;
#pragma unroll
                    for (int i = 1; i <= MVD_MAX_M; ++i)
                        if (i <= m && ((taps >> i) & 1u)) o ^= __funnelshift_l(prevU, U, i);
                    Rw[j] = o ^ E;
                }
            }
            prevU = U;
            if (valid == 32u) {
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    uint32_t r;
                    if (NOUT == 2) {
                        r = (((Rw[0] >> t) & 1u) << 1) | ((Rw[1] >> t) & 1u);
                    } else {
                        r = 0;
#pragma unroll
                        for (int j = 0; j < MVD_MAX_N; ++j)
                            if (j < n) r = (r << 1) | ((Rw[j] >> t) & 1u);
                    }
                    eng.step(t0 + (uint32_t)t, r);
                }
            } else {
                for (uint32_t t = 0; t < valid; ++t) {
                    uint32_t r = 0;
#pragma unroll
                    for (int j = 0; j < MVD_MAX_N; ++j)
                        if (j < n) r = (r << 1) | ((Rw[j] >> t) & 1u);
                    eng.step(t0 + t, r);
                }
            }
        }
    }
}

// The same driver for codes given as tables (mvd_set_code_tables / mvd_set_encoders: any k <= MVD_MAX_K inputs per step,
// viterbi_markov.py:82-106 is generic in k): the encoder of the segment's hypothesis is the table
// enc[(state << k) | u] = next state << 8 | output label, u = (u_0 .. u_{k-1}) with u_0 the most significant bit (the position
// of the tuple in itertools.product([0,1], repeat=k)); the info word of input i comes from slot 32 + i of the superblock's first
// block (bit streams: 128-bit word c = i of the k + n per superblock).
template <class Eng>
__device__ __forceinline__ void run_trial_tab(const Params& P, const DevSeg& sg, bool active, unsigned long long trial,
                                              unsigned long long tl, unsigned long long ntr, Eng& eng) {
    const int n = P.n, k = P.k;
    const uint32_t N = sg.N;
    const uint32_t T = sg.threshold;
    const int dmin = (int)sg.dmin;
    const bool philox = P.src_mode == MVD_SRC_PHILOX;
    const uint32_t c1 = (uint32_t)trial, c2 = (uint32_t)(trial >> 32), c3 = sg.stream;
    const uint16_t* enc = P.enc_tab + ((size_t)sg.enc_taps[0] << (P.m + k));
    uint32_t es = 0;                                              // encoder state 0 (alpha_exponent.py:123)
    const uint32_t nsb = (N + 127u) >> 7;
    for (uint32_t sb = 0; sb < nsb; ++sb) {
        uint4 Uw[MVD_MAX_K], Ew[MVD_MAX_N];
#pragma unroll
        for (int i = 0; i < MVD_MAX_K; ++i) Uw[i] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < MVD_MAX_N; ++j) Ew[j] = make_uint4(0, 0, 0, 0);
        if (philox) {
#pragma unroll
            for (int i = 0; i < MVD_MAX_K; ++i)
                if (i < k) Uw[i] = philox10(((4u * sb) << 6) | (32u + (uint32_t)i), c1, c2, c3, P);
        } else if (active) {
            const uint4* base = P.bits + sg.bits_offset + (unsigned long long)sb * (unsigned)(k + n) * ntr + tl;
#pragma unroll
            for (int i = 0; i < MVD_MAX_K; ++i)
                if (i < k) Uw[i] = __ldg(base + (unsigned long long)i * ntr);
#pragma unroll
            for (int j = 0; j < MVD_MAX_N; ++j)
                if (j < n) Ew[j] = __ldg(base + (unsigned long long)(k + j) * ntr);
        }
        if (!sg.random_input) {
#pragma unroll
            for (int i = 0; i < MVD_MAX_K; ++i) Uw[i] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll 1
        for (int w = 0; w < 4; ++w) {
            const uint32_t t0 = sb * 128u + (uint32_t)w * 32u;
            if (t0 >= N) break;
            const uint32_t valid = min(32u, N - t0);
            const uint32_t vmask = valid == 32u ? 0xFFFFFFFFu : ((1u << valid) - 1u);
            uint32_t Ui[MVD_MAX_K], Ej[MVD_MAX_N];
#pragma unroll
            for (int i = 0; i < MVD_MAX_K; ++i) Ui[i] = i < k ? pick(Uw[i], w) : 0u;
#pragma unroll
            for (int j = 0; j < MVD_MAX_N; ++j) {
                Ej[j] = 0u;
                if (j < n) {
                    if (philox) Ej[j] = lazy_bernoulli(((4u * sb + (uint32_t)w) << 6) | (8u * (uint32_t)j), c1, c2, c3, T, dmin,
                                                       active ? vmask : 0u, P);
                    else Ej[j] = pick(Ew[j], w);
                }
            }
#pragma unroll 1
            for (uint32_t t = 0; t < valid; ++t) {
                uint32_t u = 0, f = 0;
#pragma unroll
                for (int i = 0; i < MVD_MAX_K; ++i)
                    if (i < k) u = (u << 1) | ((Ui[i] >> t) & 1u);
#pragma unroll
                for (int j = 0; j < MVD_MAX_N; ++j)
                    if (j < n) f = (f << 1) | ((Ej[j] >> t) & 1u);
                const uint32_t br = __ldg(enc + ((es << k) | u));
                es = br >> 8;
                eng.step(t0 + t, (br & 0xFFu) ^ f);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ FSM engine
template <int MODE, bool SMEM>
struct FsmEngine {
    const uint16_t* nx16;      // SMEM
    const uint32_t* nx32;      // global
    const double2* ll;
    uint32_t* hist32;          // SMEM learn histogram
    unsigned long long* hist64;
    uint32_t* tr_idx;
    uint32_t sx;               // state * R
    uint32_t burn;
    int nshift;
    bool active;
    double a1, a0;

    __device__ __forceinline__ void step(uint32_t t, uint32_t r) {
        const uint32_t e = sx + r;
        if (MODE == MODE_DETECT) {
            const double2 v = SMEM ? ll[e] : __ldg(ll + e);
            a1 += v.x;
            a0 += v.y;
        } else if (MODE == MODE_LEARN) {
            if (active && t >= burn) {
                if (SMEM) atomicAdd(hist32 + e, 1u);
                else atomicAdd(hist64 + e, 1ull);
            }
        }
        sx = SMEM ? (uint32_t)nx16[e] : __ldg(nx32 + e);
        if (MODE == MODE_TRACE) {
            if (active) tr_idx[t + 1] = sx >> nshift;
        }
    }
};

// grid: one block per (segment, chunk of MVD_BLOCK trials)
// TAB: the code was given as tables (run_trial_tab)
template <int MODE, int NOUT, bool SMEM, bool TAB = false>
__global__ void __launch_bounds__(MVD_BLOCK) fsm_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t seg = find_segment(P, blockIdx.x);
    const DevSeg sg = P.segs[seg];
    const unsigned long long ntr = sg.trial_end - sg.trial_begin;
    const unsigned long long tl = (unsigned long long)(blockIdx.x - sg.block_begin) * MVD_BLOCK + threadIdx.x;
    const bool active = tl < ntr;
    const unsigned long long trial = sg.trial_begin + tl;

    FsmEngine<MODE, SMEM> eng;
    eng.nx16 = nullptr;
    eng.nx32 = P.nxt;
    eng.ll = P.ll + (size_t)sg.table * P.SR;
    eng.hist32 = nullptr;
    eng.hist64 = P.counts ? P.counts + (size_t)seg * P.SR : nullptr;
    eng.tr_idx = P.trace_idx ? P.trace_idx + tl * ((unsigned long long)sg.N + 1ull) : nullptr;
    eng.sx = 0;
    eng.burn = P.burn;
    eng.nshift = P.n;
    eng.active = active;
    eng.a1 = 0.0;
    eng.a0 = 0.0;
    if (SMEM) {
        // layout: [double2 ll[SR]] [uint32 hist[SR] (learn)] [uint16 nx[SR]]
        double2* s_ll = reinterpret_cast<double2*>(smem_raw);
        size_t off = (MODE == MODE_DETECT) ? sizeof(double2) * (size_t)P.SR : 0;
        uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw + off);
        if (MODE == MODE_LEARN) off += sizeof(uint32_t) * (size_t)P.SR;
        uint16_t* s_nx = reinterpret_cast<uint16_t*>(smem_raw + off);
        for (uint32_t i = threadIdx.x; i < P.SR; i += MVD_BLOCK) {
            s_nx[i] = (uint16_t)P.nxt[i];
            if (MODE == MODE_DETECT) s_ll[i] = eng.ll[i];
            if (MODE == MODE_LEARN) s_hist[i] = 0u;
        }
        __syncthreads();
        eng.nx16 = s_nx;
        eng.ll = s_ll;
        eng.hist32 = s_hist;
    }
    if (MODE == MODE_TRACE && active) eng.tr_idx[0] = 0;

    if (TAB) run_trial_tab(P, sg, active, trial, tl, ntr, eng);
    else run_trial<NOUT>(P, sg, active, trial, tl, ntr, eng);

    if (MODE == MODE_DETECT) {
        const bool win = active && (sg.decide == 0 ? (eng.a1 > eng.a0) : (eng.a1 <= eng.a0));
        const int c = __syncthreads_count(win ? 1 : 0);
        if (threadIdx.x == 0 && c) {
            atomicAdd(P.tallies + seg, (unsigned long long)c);
            if (P.tallies2) atomicAdd(P.tallies2 + seg, (unsigned long long)c);
        }
        if (P.logp && active) {
            double2* o = reinterpret_cast<double2*>(P.logp) + sg.out_offset + tl;
            *o = make_double2(eng.a1, eng.a0);
        }
    }
    if (MODE == MODE_LEARN && SMEM) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < P.SR; i += MVD_BLOCK) {
            const uint32_t c = eng.hist32[i];
            if (c) atomicAdd(eng.hist64 + i, (unsigned long long)c);
        }
    }
}

// ------------------------------------------------------------------------------------------ ACS engine
// Path metrics of the 2^M trellis states live in NP = 2^(M-1) registers, two 16-bit lanes each:
// register g holds states 2g (low half) and 2g+1 (high half).  For k = 1 the trellis is the
// shift-register butterfly: next state ns has predecessors ns>>1 and (ns>>1) + 2^(M-1), so the
// output register g (ns = 2g, 2g+1) needs D[g] and D[g + 2^(M-1)] broadcast to both halves.
template <int M>
struct AcsCore {
    static constexpr int NSTATE = 1 << M;
    static constexpr int NP = NSTATE / 2;
    static constexpr int HALF = NSTATE / 2;
    static constexpr int KW = (NSTATE + 7) / 8;      // 32-bit key words, 4 bits per state
    uint32_t D[NP];

    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int g = 0; g < NP; ++g) D[g] = 0u;
    }

    // bm = 2*NP words for the current received word: {BM0[g], BM1[g]} interleaved
    __device__ __forceinline__ void step(const uint32_t* bm) {
        uint32_t Mx[NP];
#pragma unroll
        for (int g = 0; g < NP; ++g) {
            const uint32_t A = __byte_perm(D[g >> 1], 0u, (g & 1) ? 0x3232u : 0x1010u);
            const uint32_t B = __byte_perm(D[(g + HALF) >> 1], 0u, ((g + HALF) & 1) ? 0x3232u : 0x1010u);
            const uint32_t Y = B + bm[2 * g + 1];
            Mx[g] = __viaddmin_u16x2(A, bm[2 * g], Y);          // min(A + BM0, B + BM1)   (Eq. 4)
        }
        uint32_t mn = Mx[0];
#pragma unroll
        for (int g = 1; g + 1 < NP; g += 2) mn = __vimin3_u16x2(mn, Mx[g], Mx[g + 1]);
        if (NP > 1 && (NP % 2 == 0)) mn = __vminu2(mn, Mx[NP - 1]);
        mn = __vminu2(mn, __byte_perm(mn, 0u, 0x1032u));        // both halves = global minimum
#pragma unroll
        for (int g = 0; g < NP; ++g) D[g] = Mx[g] - mn;          // Eq. 5 (no borrow: every lane >= mn)
    }

    // nibble-packed key; returns false if a metric does not fit 4 bits
    __device__ __forceinline__ bool key(uint32_t* kw) const {
        uint32_t over = 0;
#pragma unroll
        for (int i = 0; i < KW; ++i) kw[i] = 0u;
#pragma unroll
        for (int g = 0; g < NP; ++g) {
            over |= D[g] & 0xFFF0FFF0u;
            const uint32_t b = (D[g] | (D[g] >> 12)) & 0xFFu;
            kw[g >> 2] |= b << (8 * (g & 3));
        }
        return over == 0u;
    }
};

__device__ __forceinline__ uint32_t key_hash(const uint32_t* kw, int nkw) {
    uint32_t h = 0x9E3779B1u;
    for (int i = 0; i < nkw; ++i) {
        h ^= kw[i];
        h *= 0x85EBCA6Bu;
        h ^= h >> 13;
    }
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

template <int MODE, int M>
struct AcsEngine {
    AcsCore<M> core;
    const uint32_t* bm;        // shared memory, [R][2*NP]
    const uint32_t* hkeys;     // shared or global
    const uint32_t* hvals;
    uint32_t hcap;
    const double2* ll;
    uint32_t* hist32;
    unsigned long long* hist64;
    uint32_t* tr_idx;
    uint8_t* tr_met;
    int* error_flag;
    uint32_t sx, burn;
    int nshift;
    bool active, hist_smem;
    double a1, a0;
    unsigned long long h;

    __device__ __forceinline__ void step(uint32_t t, uint32_t r) {
        const uint32_t e = sx + r;
        if (MODE == MODE_DETECT) {
            const double2 v = ll[e];
            a1 += v.x;
            a0 += v.y;
        } else if (MODE == MODE_LEARN) {
            if (active && t >= burn) {
                if (hist_smem) atomicAdd(hist32 + e, 1u);
                else atomicAdd(hist64 + e, 1ull);
            }
        }
        core.step(bm + r * (2 * AcsCore<M>::NP));
        uint32_t kw[AcsCore<M>::KW];
        const bool ok = core.key(kw);
        if (MODE == MODE_HASH) {
#pragma unroll
            for (int i = 0; i < AcsCore<M>::KW; ++i) h = (h ^ (unsigned long long)kw[i]) * 0x100000001B3ull;
            if (!ok && active) atomicOr(error_flag, 2);
        } else {
            // metric vector -> Markov state: the state_index[...] lookup of Pd_plotter.py:112-113
            uint32_t slot = key_hash(kw, AcsCore<M>::KW) & (hcap - 1u);
            uint32_t val = MVD_EMPTY;
            for (uint32_t probe = 0; probe < hcap; ++probe) {
                val = hvals[slot];
                if (val == MVD_EMPTY) break;
                bool same = true;
#pragma unroll
                for (int i = 0; i < AcsCore<M>::KW; ++i) same = same && (hkeys[(size_t)i * hcap + slot] == kw[i]);
                if (same) break;
                val = MVD_EMPTY;
                slot = (slot + 1u) & (hcap - 1u);
            }
            if (val == MVD_EMPTY || !ok) {
                if (active) atomicOr(error_flag, 1);     // KeyError analogue
                val = 0u;
            }
            sx = val;
            if (MODE == MODE_TRACE && active) {
                tr_idx[t + 1] = sx >> nshift;
                if (tr_met) {
                    uint8_t* o = tr_met + (size_t)(t + 1) * AcsCore<M>::NSTATE;
#pragma unroll
                    for (int g = 0; g < AcsCore<M>::NP; ++g) {
                        o[2 * g] = (uint8_t)(core.D[g] & 0xFFu);
                        o[2 * g + 1] = (uint8_t)((core.D[g] >> 16) & 0xFFu);
                    }
                }
            }
        }
    }
};

template <int MODE, int NOUT, int M>
__global__ void __launch_bounds__(MVD_BLOCK) acs_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NP = AcsCore<M>::NP;
    constexpr int KW = AcsCore<M>::KW;
    const uint32_t seg = find_segment(P, blockIdx.x);
    const DevSeg sg = P.segs[seg];
    const unsigned long long ntr = sg.trial_end - sg.trial_begin;
    const unsigned long long tl = (unsigned long long)(blockIdx.x - sg.block_begin) * MVD_BLOCK + threadIdx.x;
    const bool active = tl < ntr;
    const unsigned long long trial = sg.trial_begin + tl;

    // shared layout: [bm: R*2*NP u32] then, when tables_in_smem:
    //   [ll double2[SR] (detect)] [hist u32[SR] (learn)] [hvals u32[hcap]] [hkeys u32[KW*hcap]]
    uint32_t* s_bm = reinterpret_cast<uint32_t*>(smem_raw);
    const uint32_t nbm = (uint32_t)P.R * 2u * NP;
    for (uint32_t i = threadIdx.x; i < nbm; i += MVD_BLOCK) s_bm[i] = P.bm[i];
    size_t off = ((size_t)nbm * 4 + 15) & ~(size_t)15;

    AcsEngine<MODE, M> eng;
    eng.core.reset();
    eng.bm = s_bm;
    eng.hkeys = P.hkeys;
    eng.hvals = P.hvals;
    eng.hcap = P.hcap;
    eng.ll = (MODE == MODE_DETECT) ? P.ll + (size_t)sg.table * P.SR : nullptr;
    eng.hist32 = nullptr;
    eng.hist64 = P.counts ? P.counts + (size_t)seg * P.SR : nullptr;
    eng.hist_smem = false;
    eng.tr_idx = P.trace_idx ? P.trace_idx + tl * ((unsigned long long)sg.N + 1ull) : nullptr;
    eng.tr_met = P.trace_met ? P.trace_met + tl * ((unsigned long long)sg.N + 1ull) * AcsCore<M>::NSTATE : nullptr;
    eng.error_flag = P.error_flag;
    eng.sx = 0;
    eng.burn = P.burn;
    eng.nshift = P.n;
    eng.active = active;
    eng.a1 = 0.0;
    eng.a0 = 0.0;
    eng.h = 0xCBF29CE484222325ull;
    if (MODE != MODE_HASH && P.tables_in_smem) {
        if (MODE == MODE_DETECT) {
            double2* s_ll = reinterpret_cast<double2*>(smem_raw + off);
            for (uint32_t i = threadIdx.x; i < P.SR; i += MVD_BLOCK) s_ll[i] = eng.ll[i];
            eng.ll = s_ll;
            off += sizeof(double2) * (size_t)P.SR;
        }
        if (MODE == MODE_LEARN) {
            uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw + off);
            for (uint32_t i = threadIdx.x; i < P.SR; i += MVD_BLOCK) s_hist[i] = 0u;
            eng.hist32 = s_hist;
            eng.hist_smem = true;
            off += sizeof(uint32_t) * (size_t)P.SR;
        }
        uint32_t* s_hv = reinterpret_cast<uint32_t*>(smem_raw + off);
        uint32_t* s_hk = s_hv + P.hcap;
        for (uint32_t i = threadIdx.x; i < P.hcap; i += MVD_BLOCK) {
            s_hv[i] = P.hvals[i];
#pragma unroll
            for (int w = 0; w < KW; ++w) s_hk[(size_t)w * P.hcap + i] = P.hkeys[(size_t)w * P.hcap + i];
        }
        eng.hvals = s_hv;
        eng.hkeys = s_hk;
    }
    __syncthreads();
    if (MODE == MODE_TRACE && active) {
        eng.tr_idx[0] = 0;
        if (eng.tr_met)
            for (int s = 0; s < AcsCore<M>::NSTATE; ++s) eng.tr_met[s] = 0;
    }

    run_trial<NOUT>(P, sg, active, trial, tl, ntr, eng);

    if (MODE == MODE_DETECT) {
        const bool win = active && (sg.decide == 0 ? (eng.a1 > eng.a0) : (eng.a1 <= eng.a0));
        const int c = __syncthreads_count(win ? 1 : 0);
        if (threadIdx.x == 0 && c) {
            atomicAdd(P.tallies + seg, (unsigned long long)c);
            if (P.tallies2) atomicAdd(P.tallies2 + seg, (unsigned long long)c);
        }
        if (P.logp && active) {
            double2* o = reinterpret_cast<double2*>(P.logp) + sg.out_offset + tl;
            *o = make_double2(eng.a1, eng.a0);
        }
    }
    if (MODE == MODE_LEARN && eng.hist_smem) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < P.SR; i += MVD_BLOCK) {
            const uint32_t c = eng.hist32[i];
            if (c) atomicAdd(eng.hist64 + i, (unsigned long long)c);
        }
    }
    if (MODE == MODE_HASH && active) {
        P.hashes[sg.out_offset + tl] = eng.h;
        if (P.final_met) {
            uint8_t* o = P.final_met + (sg.out_offset + tl) * AcsCore<M>::NSTATE;
#pragma unroll
            for (int g = 0; g < NP; ++g) {
                o[2 * g] = (uint8_t)(eng.core.D[g] & 0xFFu);
                o[2 * g + 1] = (uint8_t)((eng.core.D[g] >> 16) & 0xFFu);
            }
        }
    }
}

