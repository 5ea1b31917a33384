// mvd_bfs.cuh -- Markov-state enumeration on the GPU: the breadth-first closure of
// the all-zero metric vector under all 2^n received words, enumerate_markov_states_allzero
// (viterbi_markov.py:166-195), with the reference's *discovery order* as state index.
//
// The reference pops states in index order and, for each received word r in product order, appends
// an unseen successor at the end of the queue (viterbi_markov.py:183-193).  So the index of a state
// is the rank of its first appearance in the sequence of candidates ordered by (parent index, r).
// That order survives parallel expansion if ties are broken by candidate rank:
//
//   the queue is consumed in chunks of consecutive states [lo, lo + P); candidate c = (parent - lo) * R + r
//   1. children: one thread per parent computes Eq. 4-5 on the nibble-packed vector for all R received words
//                and writes the child keys to ckeys[c].
//      probe   : every candidate probes the open-addressing table.  A slot holds EMPTY,
//                a final state index, or TENT | rank of the lowest-ranked candidate that claimed it in
//                this chunk (atomicCAS to claim, atomicMin to lower the rank); equal keys always meet
//                in the same slot because a slot never changes its key once claimed.  Lanes pull their
//                next candidate as soon as the current one is settled (no waiting for the warp's
//                longest probe chain).
//   2. count   : winners (slot == TENT | own rank) per block
//   3. scan    : exclusive scan of the block counts (one block)
//   4. commit  : winner c gets index S + (number of winners of lower rank), stores its key there and
//                turns its slot into that final index
//   5. link    : NEXT[parent][r] = slot value (now final) for every candidate not resolved in step 1
//
// Earlier chunks are final before later ones start, so the result is the sequential BFS order for
// any chunk size.  Everything is HBM-bound hash work: per candidate one key write and read (streamed),
// ~1.5 random slot reads and ~1.5 random key reads (32-byte sectors in 64-byte DRAM granules), one NEXT write.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define BFS_EMPTY 0xFFFFFFFFu
#define BFS_TENT 0x80000000u
#define BFS_RESOLVED 0xFFFFFFFFu
#define BFS_BLOCK 256
#define BFS_SCAN_BLOCK 1024

struct BfsParams {
    uint32_t* keys;        // [max_states][KW] nibble-packed metric vectors, index = state
    uint32_t* nxt;         // [max_states][R] or null (count-only)
    uint32_t* slots;       // [mask + 1]
    uint32_t mask;
    uint32_t* ckeys;       // [chunk candidates][KW]
    uint32_t* cslot;       // [chunk candidates]
    uint32_t* winbits;     // [chunk candidates / 32, padded to whole scan blocks] winner ballots of count
    uint32_t* blocksum;    // [scan blocks + 1]; commit reads the exclusive offsets
    uint32_t* total;       // [2]: winners of this chunk, running maximum metric
    int* err;              // 1 = a relative metric exceeded 15
    uint32_t lo, nparents, S;
    int n, R;
    uint8_t lab0[64], lab1[64];   // branch labels into ns from ps = ns >> 1 and ps = (ns >> 1) + 2^(m-1)
};

template <int KW>
__device__ __forceinline__ uint32_t bfs_hash(const uint32_t* kw) {
    unsigned long long h = 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int i = 0; i < KW; ++i) {
        h ^= kw[i];
        h *= 0xD6E8FEB86659FD93ull;
        h ^= h >> 29;
    }
    h *= 0xBF58476D1CE4E5B9ull;
    return (uint32_t)(h >> 32);
}

template <int M>
struct BfsShape {
    static constexpr int NS = 1 << M;
    static constexpr int KW = NS >= 8 ? NS / 8 : 1;
};

template <int KW>
__device__ __forceinline__ void bfs_load_key(const uint32_t* p, uint32_t* k) {
    if (KW == 1) {
        k[0] = __ldcg(p);
    } else if (KW == 2) {
        const uint2 v = __ldcg(reinterpret_cast<const uint2*>(p));
        k[0] = v.x;
        k[1] = v.y;
    } else {
#pragma unroll
        for (int q = 0; q < KW / 4; ++q) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4*>(p) + q);
            k[4 * q] = v.x;
            k[4 * q + 1] = v.y;
            k[4 * q + 2] = v.z;
            k[4 * q + 3] = v.w;
        }
    }
}

template <int KW>
__device__ __forceinline__ void bfs_store_key(uint32_t* p, const uint32_t* k) {
    if (KW == 1) {
        p[0] = k[0];
    } else if (KW == 2) {
        *reinterpret_cast<uint2*>(p) = make_uint2(k[0], k[1]);
    } else {
#pragma unroll
        for (int q = 0; q < KW / 4; ++q)
            reinterpret_cast<uint4*>(p)[q] = make_uint4(k[4 * q], k[4 * q + 1], k[4 * q + 2], k[4 * q + 3]);
    }
}

// 1a. children: one thread per parent unpacks the vector once and writes the keys of its R successors.
template <int M>
__global__ void __launch_bounds__(BFS_BLOCK) bfs_child_kernel(const __grid_constant__ BfsParams P) {
    constexpr int NS = BfsShape<M>::NS, KW = BfsShape<M>::KW, HALF = NS / 2;
    const uint32_t pi = blockIdx.x * BFS_BLOCK + threadIdx.x;
    if (pi >= P.nparents) return;
    const uint32_t parent = P.lo + pi;
    uint32_t pk[KW];
#pragma unroll
    for (int w = 0; w < KW; ++w) pk[w] = P.keys[(size_t)parent * KW + w];
    int D[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) D[s] = (int)((pk[s >> 3] >> (4 * (s & 7))) & 15u);
    bool bad = false;
    for (uint32_t r = 0; r < (uint32_t)P.R; ++r) {
        int nd[NS], lo = 1 << 30;
#pragma unroll
        for (int ns = 0; ns < NS; ++ns) {
            const int a = D[ns >> 1] + __popc((uint32_t)P.lab0[ns] ^ r);            // Eq. 4
            const int b = D[(ns >> 1) + HALF] + __popc((uint32_t)P.lab1[ns] ^ r);
            nd[ns] = min(a, b);
            lo = min(lo, nd[ns]);
        }
        uint32_t ck[KW];
#pragma unroll
        for (int w = 0; w < KW; ++w) ck[w] = 0u;
        int hi = 0;
#pragma unroll
        for (int ns = 0; ns < NS; ++ns) {
            const int v = nd[ns] - lo;                                               // Eq. 5
            hi = max(hi, v);
            ck[ns >> 3] |= (uint32_t)(v & 15) << (4 * (ns & 7));
        }
        bad = bad || hi > 15;                            // does not fit a nibble: the host aborts after this pass
        bfs_store_key<KW>(P.ckeys + ((size_t)pi * P.R + r) * KW, ck);
    }
    if (bad) *P.err = 1;
}

// 1b. probe: a warp owns a tile of 32 * BFS_CPL candidates; every lane pulls its next candidate as soon as
// its current one is settled, so the lanes' probe chains (slot word -> key of the state or claimant found
// there) stay in flight together instead of idling behind the longest chain of the warp.
// (Measured alternatives, both slower on B200: 4 lock-stepped chains per thread -- fewer chains in flight;
// slot records that carry the key inline -- one granule per probe instead of two, but an 8x larger randomly
// accessed table: 1.58 ms instead of 0.99 ms per 1.68e7 candidates, 1.4 TB/s instead of 3.8 TB/s of DRAM reads.)
#define BFS_CPL 8

template <int KW>
__global__ void __launch_bounds__(BFS_BLOCK) bfs_probe_kernel(const __grid_constant__ BfsParams P) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * BFS_BLOCK + threadIdx.x) >> 5;
    const uint32_t nc = P.nparents * (uint32_t)P.R;
    const uint32_t base = warp * (32u * BFS_CPL);
    if (base >= nc) return;
    uint32_t ck[KW], c = 0, pos = 0, k = 0;
    bool have = false;
    for (;;) {
        if (!have && k < BFS_CPL) {
            c = base + k * 32u + lane;
            ++k;
            if (c < nc) {
                bfs_load_key<KW>(P.ckeys + (size_t)c * KW, ck);
                pos = bfs_hash<KW>(ck) & P.mask;
                have = true;
            }
        }
        if (!__any_sync(0xFFFFFFFFu, have || k < BFS_CPL)) break;
        if (!have) continue;
        uint32_t* rec = P.slots + pos;
        uint32_t v = __ldcg(rec);
        if (v == BFS_EMPTY) {
            v = atomicCAS(rec, BFS_EMPTY, BFS_TENT | c);
            if (v == BFS_EMPTY) {
                P.cslot[c] = pos;
                have = false;
                continue;
            }
        }
        uint32_t kk[KW];                                 // key of the claimant of this pass / of the final state
        bfs_load_key<KW>((v & BFS_TENT) ? P.ckeys + (size_t)(v & ~BFS_TENT) * KW : P.keys + (size_t)v * KW, kk);
        bool same = true;
#pragma unroll
        for (int w = 0; w < KW; ++w) same = same && (kk[w] == ck[w]);
        if (!same) {
            pos = (pos + 1u) & P.mask;
        } else if (v & BFS_TENT) {
            if ((BFS_TENT | c) < v) atomicMin(rec, BFS_TENT | c);
            P.cslot[c] = pos;
            have = false;
        } else {
            if (P.nxt) P.nxt[(size_t)P.lo * P.R + c] = v;
            P.cslot[c] = BFS_RESOLVED;
            have = false;
        }
    }
}

__device__ __forceinline__ bool bfs_is_winner(const BfsParams& P, uint32_t c, uint32_t nc) {
    if (c >= nc) return false;
    const uint32_t s = P.cslot[c];
    return s != BFS_RESOLVED && P.slots[s] == (BFS_TENT | c);
}

__global__ void __launch_bounds__(BFS_SCAN_BLOCK) bfs_count_kernel(const __grid_constant__ BfsParams P) {
    const uint32_t c = blockIdx.x * BFS_SCAN_BLOCK + threadIdx.x;
    const bool win = bfs_is_winner(P, c, P.nparents * (uint32_t)P.R);
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, win);
    if ((threadIdx.x & 31u) == 0) P.winbits[c >> 5] = bal;       // commit reads the ballots instead of probing again
    const int n = __syncthreads_count(win ? 1 : 0);
    if (threadIdx.x == 0) P.blocksum[blockIdx.x] = (uint32_t)n;
}

// exclusive scan of blocksum[0 .. nb) in place; blocksum[nb] and total[0] = sum
__global__ void __launch_bounds__(BFS_SCAN_BLOCK) bfs_scan_kernel(uint32_t* blocksum, uint32_t nb, uint32_t* total) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nb; base += BFS_SCAN_BLOCK) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nb ? blocksum[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= (uint32_t)d) x += y;
        }
        if (lane == 31u) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, d);
                if (lane >= (uint32_t)d) w += y;
            }
            wsum[lane] = w;                              // inclusive over warps
        }
        __syncthreads();
        const uint32_t before = carry + (wid ? wsum[wid - 1] : 0u) + (x - v);
        if (i < nb) blocksum[i] = before;
        __syncthreads();
        if (threadIdx.x == BFS_SCAN_BLOCK - 1) carry += wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        blocksum[nb] = carry;
        total[0] = carry;
    }
}

template <int KW>
__global__ void __launch_bounds__(BFS_SCAN_BLOCK) bfs_commit_kernel(const __grid_constant__ BfsParams P) {
    __shared__ uint32_t wsum[32];
    const uint32_t c = blockIdx.x * BFS_SCAN_BLOCK + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t bal = P.winbits[c >> 5];
    const bool win = (bal >> lane) & 1u;
    if (lane == 0) wsum[wid] = (uint32_t)__popc(bal);
    __syncthreads();
    if (wid == 0) {
        uint32_t w = wsum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if (lane >= (uint32_t)d) w += y;
        }
        wsum[lane] = w;
    }
    __syncthreads();
    if (!win) return;
    const uint32_t idx = P.S + P.blocksum[blockIdx.x] + (wid ? wsum[wid - 1] : 0u) + (uint32_t)__popc(bal & ((1u << lane) - 1u));
    uint32_t mx = 0;
#pragma unroll
    for (int w = 0; w < KW; ++w) {
        const uint32_t kw = P.ckeys[(size_t)c * KW + w];
        P.keys[(size_t)idx * KW + w] = kw;
#pragma unroll
        for (int b = 0; b < 8; ++b) mx = max(mx, (kw >> (4 * b)) & 15u);
    }
    P.slots[P.cslot[c]] = idx;
    if (mx > P.total[1]) atomicMax(P.total + 1, mx);
}

__global__ void __launch_bounds__(BFS_BLOCK) bfs_link_kernel(const __grid_constant__ BfsParams P) {
    const uint32_t c = blockIdx.x * BFS_BLOCK + threadIdx.x;
    if (c >= P.nparents * (uint32_t)P.R) return;
    const uint32_t s = P.cslot[c];
    if (s == BFS_RESOLVED) return;
    P.nxt[(size_t)P.lo * P.R + c] = P.slots[s];
}
