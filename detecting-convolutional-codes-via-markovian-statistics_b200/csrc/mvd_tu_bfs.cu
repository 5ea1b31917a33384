// translation unit: GPU breadth-first enumeration of the Markov states (mvd_bfs.cuh)
#include "mvd_bfs.cuh"
#include "mvd_launch.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace {

struct Buf {
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max(bytes, (size_t)256)); }
    ~Buf() {
        if (p) cudaFree(p);
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

template <int M>
cudaError_t expand(const BfsParams& P, uint32_t nc, cudaStream_t st) {
    constexpr int KW = BfsShape<M>::KW;
    bfs_child_kernel<M><<<(P.nparents + BFS_BLOCK - 1) / BFS_BLOCK, BFS_BLOCK, 0, st>>>(P);
    const uint32_t per_block = BFS_BLOCK * BFS_CPL;
    bfs_probe_kernel<KW><<<(nc + per_block - 1) / per_block, BFS_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

template <int KW>
cudaError_t commit(const BfsParams& P, uint32_t nb, cudaStream_t st) {
    bfs_commit_kernel<KW><<<nb, BFS_SCAN_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

}  // namespace

#define BCK(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            snprintf(res.error, sizeof res.error, "%s failed: %s", #call, cudaGetErrorString(e__));   \
            return e__ == cudaErrorMemoryAllocation ? MVD_E_NOMEM : MVD_E_CUDA;                       \
        }                                                                                             \
    } while (0)

int mvd_bfs_run(const MvdBfsConfig& cfg, cudaStream_t st, MvdBfsResult& res) {
    const int m = cfg.m, n = cfg.n, R = 1 << n, NS = 1 << m, HALF = NS / 2, KW = NS >= 8 ? NS / 8 : 1;
    res = MvdBfsResult();
    if (cfg.max_states == 0 || cfg.max_states > MVD_BFS_MAX_STATES) {
        snprintf(res.error, sizeof res.error, "max_states %u outside [1, %u]", cfg.max_states, MVD_BFS_MAX_STATES);
        return MVD_E_INVALID;
    }
    uint32_t chunk = cfg.chunk_parents ? cfg.chunk_parents : (1u << 22);
    chunk = std::min(chunk, (1u << 28) / (uint32_t)R);
    chunk = std::min(chunk, cfg.max_states);
    const size_t maxc = (size_t)chunk * R;
    uint64_t cap = 1024;
    while (cap < 2ull * cfg.max_states && cap < (1ull << 31)) cap <<= 1;      // load <= 1/2 (<= 3/4 beyond 2^30 states)
    const size_t nb_max = (maxc + BFS_SCAN_BLOCK - 1) / BFS_SCAN_BLOCK;
    const size_t need = (size_t)cfg.max_states * KW * 4 + (cfg.keep_next ? (size_t)cfg.max_states * R * 4 : 0) + cap * 4 +
                        maxc * KW * 4 + maxc * 4 + (nb_max + 1) * 4;
    size_t free_b = 0, total_b = 0;
    BCK(cudaMemGetInfo(&free_b, &total_b));
    if (need + (1ull << 30) > free_b) {
        snprintf(res.error, sizeof res.error, "enumeration of up to %u states needs %.1f GB of device memory, %.1f GB free",
                 cfg.max_states, need / 1e9, free_b / 1e9);
        return MVD_E_NOMEM;
    }
    Buf keys, nxt, slots, ckeys, cslot, bsum, misc, wbits;
    BCK(keys.alloc((size_t)cfg.max_states * KW * 4));
    if (cfg.keep_next) BCK(nxt.alloc((size_t)cfg.max_states * R * 4));
    BCK(slots.alloc(cap * 4));
    BCK(ckeys.alloc(maxc * KW * 4));
    BCK(cslot.alloc(maxc * 4));
    BCK(bsum.alloc((nb_max + 1) * 4));
    BCK(wbits.alloc(nb_max * (BFS_SCAN_BLOCK / 32) * 4));
    BCK(misc.alloc(16));
    cudaEvent_t e0, e1;
    BCK(cudaEventCreate(&e0));
    BCK(cudaEventCreate(&e1));
    BCK(cudaEventRecord(e0, st));
    BCK(cudaMemsetAsync(slots.p, 0xFF, cap * 4, st));
    BCK(cudaMemsetAsync(keys.p, 0, (size_t)KW * 4, st));          // state 0 = the all-zero vector (viterbi_markov.py:177)
    BCK(cudaMemsetAsync(misc.p, 0, 16, st));

    BfsParams P;
    memset(&P, 0, sizeof P);
    P.keys = keys.as<uint32_t>();
    P.nxt = nxt.as<uint32_t>();
    P.slots = slots.as<uint32_t>();
    P.mask = (uint32_t)(cap - 1);
    P.ckeys = ckeys.as<uint32_t>();
    P.cslot = cslot.as<uint32_t>();
    P.blocksum = bsum.as<uint32_t>();
    P.winbits = wbits.as<uint32_t>();
    P.total = misc.as<uint32_t>();
    P.err = reinterpret_cast<int*>(misc.as<uint32_t>() + 2);
    P.n = n;
    P.R = R;
    for (int ns = 0; ns < NS; ++ns) {
        auto lab = [&](uint32_t ps, uint32_t u) {
            const uint32_t reg = (u & 1u) | (ps << 1);
            int l = 0;
            for (int j = 0; j < n; ++j) l = (l << 1) | (__builtin_popcount(reg & cfg.dec_taps[j]) & 1);
            return (uint8_t)l;
        };
        P.lab0[ns] = lab((uint32_t)(ns >> 1), (uint32_t)(ns & 1));
        P.lab1[ns] = lab((uint32_t)((ns >> 1) + HALF), (uint32_t)(ns & 1));
    }
    // state 0 enters the table through the same hash the kernels use: seed it with a 1-candidate
    // pseudo-chunk is not possible (no parent), so insert it on the host side of the protocol:
    // the slot index of the all-zero key is computed by a tiny kernel-free replica of bfs_hash.
    {
        unsigned long long h = 0x9E3779B97F4A7C15ull;
        for (int i = 0; i < KW; ++i) {
            h ^= 0u;
            h *= 0xD6E8FEB86659FD93ull;
            h ^= h >> 29;
        }
        h *= 0xBF58476D1CE4E5B9ull;
        const uint32_t slot = (uint32_t)(h >> 32) & P.mask;
        BCK(cudaMemsetAsync(P.slots + slot, 0, 4, st));                        // state 0 is final from the start
    }

    uint32_t S = 1, lo = 0, level_end = 1;
    uint32_t host_misc[4] = {0, 0, 0, 0};
    int rc = MVD_OK;
    res.levels.push_back(1);                              // level 0 = {all-zero}
    while (lo < S) {
        if (lo == level_end) {                           // the queue reached the first state of the next BFS level
            res.levels.push_back(S - level_end);
            level_end = S;
        }
        const uint32_t np = std::min(chunk, level_end - lo), nc = np * (uint32_t)R;
        const uint32_t nb = (nc + BFS_SCAN_BLOCK - 1) / BFS_SCAN_BLOCK;
        P.lo = lo;
        P.nparents = np;
        P.S = S;
        cudaError_t e;
        switch (m) {
            case 1: e = expand<1>(P, nc, st); break;
            case 2: e = expand<2>(P, nc, st); break;
            case 3: e = expand<3>(P, nc, st); break;
            case 4: e = expand<4>(P, nc, st); break;
            case 5: e = expand<5>(P, nc, st); break;
            case 6: e = expand<6>(P, nc, st); break;
            default: e = cudaErrorInvalidValue;
        }
        BCK(e);
        bfs_count_kernel<<<nb, BFS_SCAN_BLOCK, 0, st>>>(P);
        bfs_scan_kernel<<<1, BFS_SCAN_BLOCK, 0, st>>>(P.blocksum, nb, P.total);
        BCK(cudaGetLastError());
        BCK(cudaMemcpyAsync(host_misc, misc.p, 12, cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        res.launches += 4;
        res.iterations += 1;
        res.candidates += nc;
        if (host_misc[2]) {
            snprintf(res.error, sizeof res.error, "a relative metric exceeds 15 (nibble-packed keys)");
            rc = MVD_E_UNSUPPORTED;
            break;
        }
        const uint32_t nnew = host_misc[0];
        if ((uint64_t)S + nnew > cfg.max_states) {
            snprintf(res.error, sizeof res.error, "more than %u Markov states (%u enumerated, queue at %u)", cfg.max_states, S, lo);
            rc = MVD_E_NOMEM;
            break;
        }
        switch (KW) {
            case 1: e = commit<1>(P, nb, st); break;
            case 2: e = commit<2>(P, nb, st); break;
            case 4: e = commit<4>(P, nb, st); break;
            default: e = commit<8>(P, nb, st); break;
        }
        BCK(e);
        res.launches += 1;
        if (cfg.keep_next) {
            bfs_link_kernel<<<(nc + BFS_BLOCK - 1) / BFS_BLOCK, BFS_BLOCK, 0, st>>>(P);
            BCK(cudaGetLastError());
            res.launches += 1;
        }
        S += nnew;
        lo += np;
    }
    BCK(cudaMemcpyAsync(host_misc, misc.p, 12, cudaMemcpyDeviceToHost, st));
    BCK(cudaEventRecord(e1, st));
    BCK(cudaStreamSynchronize(st));
    BCK(cudaEventElapsedTime(&res.ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    res.S = S;
    res.frontier = lo;
    res.closed = rc == MVD_OK;
    res.max_metric = (int)host_misc[1];
    if (rc == MVD_OK && cfg.copy_out) {
        res.metrics.resize((size_t)S * NS);
        std::vector<uint32_t> hk((size_t)S * KW);
        BCK(cudaMemcpy(hk.data(), keys.p, hk.size() * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < (size_t)S; ++i)
            for (int s = 0; s < NS; ++s) res.metrics[i * NS + s] = (uint8_t)((hk[i * KW + (s >> 3)] >> (4 * (s & 7))) & 15u);
        if (cfg.keep_next) {
            res.next.resize((size_t)S * R);
            BCK(cudaMemcpy(res.next.data(), nxt.p, res.next.size() * 4, cudaMemcpyDeviceToHost));
        }
    }
    return rc;
}
