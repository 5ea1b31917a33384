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
//   1. expand  : one thread per candidate computes Eq. 4-5 on the nibble-packed parent vector, writes
//                the child key to ckeys[c] and probes the open-addressing table.  A slot holds EMPTY,
//                a final state index, or TENT | rank of the lowest-ranked candidate that claimed it in
//                this chunk (atomicCAS to claim, atomicMin to lower the rank); equal keys always meet
//                in the same slot because a slot never changes its key once claimed.
//   2. count   : winners (slot == TENT | own rank) per block
//   3. scan    : exclusive scan of the block counts (one block)
//   4. commit  : winner c gets index S + (number of winners of lower rank), stores its key there and
//                turns its slot into that final index
//   5. link    : NEXT[parent][r] = slot value (now final) for every candidate not resolved in step 1
//
// Earlier chunks are final before later ones start, so the result is the sequential BFS order for
// any chunk size.  Everything is HBM/L2-bound hash work: per candidate one key write, ~1.5 random
// slot reads, one random key read, one NEXT write.
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

template <int M>
__global__ void __launch_bounds__(BFS_BLOCK) bfs_expand_kernel(const __grid_constant__ BfsParams P) {
    constexpr int NS = BfsShape<M>::NS, KW = BfsShape<M>::KW, HALF = NS / 2;
    const uint32_t c = blockIdx.x * BFS_BLOCK + threadIdx.x;
    const uint32_t nc = P.nparents * (uint32_t)P.R;
    if (c >= nc) return;
    const uint32_t parent = P.lo + c / (uint32_t)P.R, r = c % (uint32_t)P.R;
    uint32_t pk[KW];
#pragma unroll
    for (int w = 0; w < KW; ++w) pk[w] = P.keys[(size_t)parent * KW + w];
    int D[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) D[s] = (int)((pk[s >> 3] >> (4 * (s & 7))) & 15u);
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
    if (hi > 15) {                                       // does not fit a nibble: the host aborts after this kernel
        *P.err = 1;
        P.cslot[c] = BFS_RESOLVED;
        return;
    }
#pragma unroll
    for (int w = 0; w < KW; ++w) P.ckeys[(size_t)c * KW + w] = ck[w];
    __threadfence();                                     // the key is visible before the claim is
    uint32_t slot = bfs_hash<KW>(ck) & P.mask;
    for (;;) {
        uint32_t v = *reinterpret_cast<volatile uint32_t*>(P.slots + slot);
        if (v == BFS_EMPTY) {
            v = atomicCAS(P.slots + slot, BFS_EMPTY, BFS_TENT | c);
            if (v == BFS_EMPTY) {
                P.cslot[c] = slot;
                return;
            }
        }
        bool same = true;
        if (v & BFS_TENT) {
            const uint32_t c2 = v & ~BFS_TENT;
#pragma unroll
            for (int w = 0; w < KW; ++w) same = same && (__ldcg(P.ckeys + (size_t)c2 * KW + w) == ck[w]);
            if (same) {
                if (c < c2) atomicMin(P.slots + slot, BFS_TENT | c);
                P.cslot[c] = slot;
                return;
            }
        } else {
#pragma unroll
            for (int w = 0; w < KW; ++w) same = same && (P.keys[(size_t)v * KW + w] == ck[w]);
            if (same) {
                if (P.nxt) P.nxt[(size_t)parent * P.R + r] = v;
                P.cslot[c] = BFS_RESOLVED;
                return;
            }
        }
        slot = (slot + 1u) & P.mask;
    }
}

__device__ __forceinline__ bool bfs_is_winner(const BfsParams& P, uint32_t c, uint32_t nc) {
    if (c >= nc) return false;
    const uint32_t s = P.cslot[c];
    return s != BFS_RESOLVED && P.slots[s] == (BFS_TENT | c);
}

__global__ void __launch_bounds__(BFS_SCAN_BLOCK) bfs_count_kernel(const __grid_constant__ BfsParams P) {
    const uint32_t c = blockIdx.x * BFS_SCAN_BLOCK + threadIdx.x;
    const int n = __syncthreads_count(bfs_is_winner(P, c, P.nparents * (uint32_t)P.R) ? 1 : 0);
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
    const bool win = bfs_is_winner(P, c, P.nparents * (uint32_t)P.R);
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, win);
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
