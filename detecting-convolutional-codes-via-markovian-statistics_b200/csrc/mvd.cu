// mvd.cu -- C ABI (include/mvd.h) over the sm_100a kernels in mvd_kernels.cuh.
//
// Host responsibilities: validate arguments, turn the decoder code into branch-metric constants,
// build the metric-vector hash table and the premultiplied NEXT table, stage segments, launch
// one kernel per call (a whole sweep = one launch), read tallies / counts back.
// There is no CPU fallback: every compute entry point needs a CUDA device.
#include "mvd_launch.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max(bytes, (size_t)256);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct mvd_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaDeviceProp prop{};
    std::string err;
    uint64_t launches = 0;
    float last_ms = 0.f;
    int last_fast = 0;              // 0 = generic kernel, else 1 + lookup kind + 16 * log2(log-row stride)
    bool force_generic = false, no_pair = false, force_pair = false, no_fsm1 = false, no_antipodal = false;
    int split_mode = 0;             // 0 = automatic, 1 = always split long trials along the time axis, 2 = never
    bool split_sequential = false;  // MVD_OPT_SPLIT_SEQUENTIAL: the split path adds every term one by one (no re-association)
    bool table_code = false;        // mvd_set_code_tables: trellis / encoders as tables (any k), generic NEXT-walk kernels only
    std::vector<uint8_t> h_dec_prev, h_dec_lab;   // [2^m][2^k]
    uint32_t nenc = 0;
    uint32_t split_chunk = 0;       // MVD_OPT_SPLIT_CHUNK: 0 = automatic, else 256 / 512 / 1024 steps per chunk
    SplitClasses split_cls{};       // log Tref has <= 3 distinct non-zero values: the split path counts them (class mode)
    bool no_split_classes = false;  // MVD_OPT_SPLIT_SEQUENTIAL = 2: re-association without the class mode (both sums as recurrences)
    bool split_tables_ready = false;// tie binades / float32 terms of the current log-likelihood tables are on the device
    unsigned long long last_split_sub = 0, last_split_seq = 0;   // sub-chunks of the last split launch / of them added term by term
    bool have_gfsm1 = false;
    bool tref_packed = false;       // log Tref = c * unit with c in {0, 2^j}: one-load NEXT walk possible
    double tref_unit = 0.0;
    std::vector<uint32_t> bfs_levels;   // level sizes of the last GPU enumeration
    uint32_t last_dirty = 0;        // chunk-parallel learning: chunks that needed the fix-up pass
    void* pin[2] = {nullptr, nullptr};      // pinned staging halves of the large copies (h2d / d2h)
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    uint64_t h2d_bytes = 0, d2h_bytes = 0;   // counted at every host<->device copy this context issues (mvd_copy_stats)
    // asynchronous detection launches (MVD_OPT_ASYNC_DETECT): mvd_detect with a device tally destination only returns
    // once the work is queued; mvd_synchronize drains: waits, reads the error flag, adds up the kernel times
    bool async_detect = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> async_ev;   // event pairs of the launches in flight (ring, <= 64)
    uint32_t async_pending = 0;
    double async_ms_sum = 0.0;
    uint64_t async_launches = 0;
    uint32_t learn_warm = LEARN_WARM;
    bool learn_warm_set = false;    // MVD_OPT_LEARN_WARM given: mvd_set_code keeps it

    // code
    bool have_code = false;
    int k = 0, n = 0, m = 0;
    uint32_t dec_taps[MVD_MAX_N] = {0, 0, 0, 0};
    // states
    bool have_states = false, acs_ok = false;
    bool closed = false;            // every successor of every state is in the table (fast kernels)
    int max_metric = 0;
    uint32_t nkeys = 0;             // direct metric-vector -> state table (m <= 2), 0 = none
    bool linkey_ok = false;         // m = 2: offset-invariant linear key of the pair kernel found (d_dstate2)
    int32_t kc[4] = {0, 0, 0, 0};   //   key = sum kc[s] D[s] + kbias, sum kc = 0, injective on the states, in [0, 256)
    int32_t kbias = 0;
    uint32_t S = 0;
    std::vector<uint8_t> h_metrics;
    std::vector<uint32_t> h_next;
    uint32_t hcap = 0;
    uint32_t ph_slots = 0, ph_bshift = 24, ph_c2 = 0, ph_c4 = 0;   // m = 3 / 4 perfect hash (mvd_detect3p.cuh), 0 = none
    uint32_t ph_nb = 0, ph_slot0 = 0;
    bool have_llslot = false;       // m = 4: log rows stored by hash slot (d_llslot) for the loaded tables
    // loglik
    uint32_t ntables = 0;

    DevBuf d_bm, d_nxt, d_ll, d_hkeys, d_hvals, d_segs, d_tallies, d_counts, d_logp, d_trace_idx, d_trace_met,
        d_hashes, d_final, d_err, d_bits, d_peak, d_dstate, d_dstate2, d_stage, d_lspec, d_lend, d_ldirty, d_tcode, d_gfsm1, d_smeta, d_sedges, d_phd, d_pht, d_llslot, d_sapx, d_splan, d_sres, d_stie, d_sapxtab, d_sflags, d_stiek, d_enc;
};

namespace {

int fail(mvd_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

// every host<->device copy of the library goes through these two: the byte counters are what bench.py reports
// Large copies (edge counts and log-likelihood tables of an m = 4 code: 34 MB each way) go through two pinned 8 MB
// buffers of the context: DMA at link speed into / out of pinned memory while host threads move the other half to /
// from the caller's pageable array (a plain cudaMemcpyAsync on pageable memory ran at ~4 GB/s: 8.5 ms per table set).
constexpr size_t PIN_CHUNK = (size_t)8 << 20;
constexpr size_t PIN_MIN = (size_t)2 << 20;

// fn(lo, hi) over [0, count) on up to 16 host threads (ranges of at least `grain`); the element-wise work of the
// table helpers below does not depend on how the range is cut
template <class F>
void host_parallel(uint64_t count, uint64_t grain, F fn) {
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt == 0 ? 1u : (nt > 16u ? 16u : nt);
    const uint64_t want = (count + grain - 1) / grain;
    if (want < nt) nt = (unsigned)(want ? want : 1);
    if (nt <= 1) {
        fn(0, count);
        return;
    }
    std::vector<std::thread> th;
    const uint64_t per = (count + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        const uint64_t lo = (uint64_t)t * per, hi = lo + per < count ? lo + per : count;
        if (lo < hi) th.emplace_back([=, &fn] { fn(lo, hi); });
    }
    for (auto& x : th) x.join();
}

inline cudaError_t pin_ready(mvd_ctx* c) {
    if (c->pin[0]) return cudaSuccess;
    for (int i = 0; i < 2; ++i) {
        cudaError_t e = cudaHostAlloc(&c->pin[i], PIN_CHUNK, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->pin_ev[i], cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

inline void host_copy(void* dst, const void* src, size_t bytes) {
    host_parallel(bytes, (size_t)1 << 20, [&](uint64_t lo, uint64_t hi) {
        memcpy(static_cast<char*>(dst) + lo, static_cast<const char*>(src) + lo, hi - lo);
    });
}

inline cudaError_t h2d(mvd_ctx* c, void* dst, const void* src, size_t bytes) {
    c->h2d_bytes += bytes;
    if (bytes < PIN_MIN || pin_ready(c) != cudaSuccess) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream);
    int half = 0;
    bool used[2] = {false, false};
    for (size_t off = 0; off < bytes; off += PIN_CHUNK, half ^= 1) {
        const size_t len = bytes - off < PIN_CHUNK ? bytes - off : PIN_CHUNK;
        if (used[half]) {
            cudaError_t e = cudaEventSynchronize(c->pin_ev[half]);     // the DMA that last read this half
            if (e != cudaSuccess) return e;
        }
        host_copy(c->pin[half], static_cast<const char*>(src) + off, len);
        cudaError_t e = cudaMemcpyAsync(static_cast<char*>(dst) + off, c->pin[half], len, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->pin_ev[half], c->stream);
        if (e != cudaSuccess) return e;
        used[half] = true;
    }
    for (int i = 0; i < 2; ++i)
        if (used[i]) {
            cudaError_t e = cudaEventSynchronize(c->pin_ev[i]);        // the pinned halves are free again on return
            if (e != cudaSuccess) return e;
        }
    return cudaSuccess;
}
inline cudaError_t d2h(mvd_ctx* c, void* dst, const void* src, size_t bytes) {
    c->d2h_bytes += bytes;
    if (bytes < PIN_MIN || pin_ready(c) != cudaSuccess) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream);
    // chunk i + 1 is on the wire while chunk i is copied out of its pinned half; complete on return
    const size_t nchunks = (bytes + PIN_CHUNK - 1) / PIN_CHUNK;
    auto issue = [&](size_t i) -> cudaError_t {
        const size_t off = i * PIN_CHUNK, len = bytes - off < PIN_CHUNK ? bytes - off : PIN_CHUNK;
        cudaError_t e = cudaMemcpyAsync(c->pin[i & 1], static_cast<const char*>(src) + off, len, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->pin_ev[i & 1], c->stream);
        return e;
    };
    cudaError_t e = issue(0);
    if (e != cudaSuccess) return e;
    for (size_t i = 0; i < nchunks; ++i) {
        if (i + 1 < nchunks && (e = issue(i + 1)) != cudaSuccess) return e;
        if ((e = cudaEventSynchronize(c->pin_ev[i & 1])) != cudaSuccess) return e;
        const size_t off = i * PIN_CHUNK, len = bytes - off < PIN_CHUNK ? bytes - off : PIN_CHUNK;
        host_copy(static_cast<char*>(dst) + off, c->pin[i & 1], len);
    }
    return cudaSuccess;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(ctx, MVD_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

inline int label_of_branch(uint32_t ps, uint32_t u, const uint32_t* taps, int n) {
    const uint32_t reg = (u & 1u) | (ps << 1);
    int lab = 0;
    for (int j = 0; j < n; ++j) lab = (lab << 1) | (__builtin_popcount(reg & taps[j]) & 1);
    return lab;
}

// nibble-packed key of a metric vector (must match AcsCore<M>::key)
inline bool pack_key(const uint8_t* met, int nstate, uint32_t* kw, int nkw) {
    for (int i = 0; i < nkw; ++i) kw[i] = 0;
    bool ok = true;
    for (int s = 0; s < nstate; ++s) {
        if (met[s] > 15) ok = false;
        kw[s >> 3] |= (uint32_t)(met[s] & 15u) << (4 * (s & 7));
    }
    return ok;
}

inline uint32_t host_key_hash(const uint32_t* kw, int nkw) {
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

int install_states(mvd_ctx* ctx) {
    const int nstate = 1 << ctx->m, R = 1 << ctx->n;
    const uint32_t S = ctx->S;
    const size_t SR = (size_t)S * R;
    if (SR >= 0xFFFFFFF0ull / (size_t)R) return fail(ctx, MVD_E_UNSUPPORTED, "state table too large (S=%u)", S);
    std::vector<uint32_t> pre(SR);
    for (size_t i = 0; i < SR; ++i) {
        if (ctx->h_next[i] >= S) return fail(ctx, MVD_E_INVALID, "next[%zu]=%u out of range (S=%u)", i, ctx->h_next[i], S);
        pre[i] = ctx->h_next[i] * (uint32_t)R;
    }
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_nxt.reserve(SR * 4));
    CK(h2d(ctx, ctx->d_nxt.p, pre.data(), SR * 4));
    // hash table: metric vector -> state * R
    const int nkw = (nstate + 7) / 8;
    uint32_t cap = 64;
    while (cap < 2ull * S) cap <<= 1;
    // empty slots carry the key 0xFFFFFFFF.. (no metric vector packs to it: min-normalised vectors contain a 0)
    std::vector<uint32_t> keys((size_t)nkw * cap, 0xFFFFFFFFu), vals(cap, MVD_EMPTY);
    bool ok = true;
    for (uint32_t i = 0; i < S && ok; ++i) {
        uint32_t kw[8];
        if (!pack_key(ctx->h_metrics.data() + (size_t)i * nstate, nstate, kw, nkw)) { ok = false; break; }
        uint32_t slot = host_key_hash(kw, nkw) & (cap - 1);
        while (vals[slot] != MVD_EMPTY) {
            bool same = true;
            for (int w = 0; w < nkw; ++w) same = same && keys[(size_t)w * cap + slot] == kw[w];
            if (same) return fail(ctx, MVD_E_INVALID, "duplicate metric vector at state %u", i);
            slot = (slot + 1) & (cap - 1);
        }
        vals[slot] = i * (uint32_t)R;
        for (int w = 0; w < nkw; ++w) keys[(size_t)w * cap + slot] = kw[w];
    }
    ctx->acs_ok = ok;
    ctx->hcap = cap;
    CK(ctx->d_hkeys.reserve(keys.size() * 4));
    CK(ctx->d_hvals.reserve(vals.size() * 4));
    CK(h2d(ctx, ctx->d_hkeys.p, keys.data(), keys.size() * 4));
    CK(h2d(ctx, ctx->d_hvals.p, vals.data(), vals.size() * 4));
    // ---- closure: metrics[next[i][r]] == Eq.4-5(metrics[i], r) for every (i, r), state 0 = all-zero.
    // Only closed tables may take the fast kernels, which do not check for unknown vectors.
    {
        const int HALF = nstate / 2;
        std::vector<int> lab0(nstate), lab1(nstate);
        for (int ns = 0; ns < nstate; ++ns) {
            lab0[ns] = label_of_branch((uint32_t)(ns >> 1), (uint32_t)(ns & 1), ctx->dec_taps, ctx->n);
            lab1[ns] = label_of_branch((uint32_t)((ns >> 1) + HALF), (uint32_t)(ns & 1), ctx->dec_taps, ctx->n);
        }
        bool closed = true;
        int mx = 0;
        for (int s = 0; s < nstate; ++s) closed = closed && ctx->h_metrics[s] == 0;
        for (uint32_t i = 0; i < S && closed; ++i) {
            const uint8_t* cur = ctx->h_metrics.data() + (size_t)i * nstate;
            for (int r = 0; r < R && closed; ++r) {
                int tmp[64], lo = 1 << 30;
                for (int ns = 0; ns < nstate; ++ns) {
                    if (ctx->table_code) {                       // 2^k incoming branches from the tables (viterbi_markov.py:150-155)
                        const int nb = 1 << ctx->k;
                        int best = 1 << 30;
                        for (int b = 0; b < nb; ++b) {
                            const int v = cur[ctx->h_dec_prev[(size_t)ns * nb + b]] + __builtin_popcount((unsigned)(ctx->h_dec_lab[(size_t)ns * nb + b] ^ r));
                            best = v < best ? v : best;
                        }
                        tmp[ns] = best;
                    } else {
                        const int a = cur[ns >> 1] + __builtin_popcount((unsigned)(lab0[ns] ^ r));
                        const int b = cur[(ns >> 1) + HALF] + __builtin_popcount((unsigned)(lab1[ns] ^ r));
                        tmp[ns] = a < b ? a : b;
                    }
                    lo = tmp[ns] < lo ? tmp[ns] : lo;
                }
                const uint8_t* nx = ctx->h_metrics.data() + (size_t)ctx->h_next[(size_t)i * R + r] * nstate;
                for (int ns = 0; ns < nstate; ++ns) closed = closed && (tmp[ns] - lo) == (int)nx[ns];
            }
        }
        for (size_t i = 0; i < ctx->h_metrics.size(); ++i) mx = std::max(mx, (int)ctx->h_metrics[i]);
        ctx->closed = closed;
        ctx->max_metric = mx;
    }
    // ---- direct table for m <= 2: key = hi + (lo << b), lo = s0 | s2 << 2b, hi = s1 | s3 << 2b
    // (must match Acs2Engine::step), b = 4 (m = 1) or 2 (m = 2)
    ctx->nkeys = 0;
    if (ctx->closed && ((ctx->m == 1 && ctx->max_metric <= 15) || (ctx->m == 2 && ctx->max_metric <= 3))) {
        const int b = ctx->m == 1 ? 4 : 2;
        const uint32_t nkeys = 256;
        std::vector<uint16_t> dst(nkeys, (uint16_t)0xFFFF);
        for (uint32_t i = 0; i < S; ++i) {
            const uint8_t* v = ctx->h_metrics.data() + (size_t)i * nstate;
            uint32_t lo = v[0], hi = v[1];
            if (ctx->m == 2) {
                lo |= (uint32_t)v[2] << (2 * b);
                hi |= (uint32_t)v[3] << (2 * b);
            }
            dst[hi + (lo << b)] = (uint16_t)i;
        }
        CK(ctx->d_dstate.reserve(nkeys * 2));
        CK(h2d(ctx, ctx->d_dstate.p, dst.data(), nkeys * 2));
        ctx->nkeys = S <= 0xFFFE ? nkeys : 0;
    }
    // ---- m = 2 pair kernel: offset-invariant key.  key(D) = sum_s c_s D[s] + bias with sum_s c_s = 0 is the same for
    // D' and D' - min(D'), so un-normalised steps need no minimum at all (mvd_detect2.cuh, PairEngine::step).  The
    // coefficients are searched here: injective on this decoder's S metric vectors with all keys inside [0, 256).
    ctx->linkey_ok = false;
    if (ctx->nkeys && ctx->m == 2) {
        const int spans[2] = {12, 48};
        for (int pass = 0; pass < 2 && !ctx->linkey_ok; ++pass) {
            const int Rg = spans[pass];
            for (int c1 = -Rg; c1 <= Rg && !ctx->linkey_ok; ++c1)
                for (int c2 = -Rg; c2 <= Rg && !ctx->linkey_ok; ++c2)
                    for (int c3 = -Rg; c3 <= Rg; ++c3) {
                        const int c0 = -(c1 + c2 + c3);
                        int lo = 1 << 30, hi = -(1 << 30);
                        for (uint32_t i = 0; i < S; ++i) {
                            const uint8_t* v = ctx->h_metrics.data() + (size_t)i * nstate;
                            const int key = c0 * v[0] + c1 * v[1] + c2 * v[2] + c3 * v[3];
                            lo = key < lo ? key : lo;
                            hi = key > hi ? key : hi;
                        }
                        if (hi - lo > 255) continue;
                        uint64_t seen[4] = {0, 0, 0, 0};
                        bool inj = true;
                        for (uint32_t i = 0; i < S && inj; ++i) {
                            const uint8_t* v = ctx->h_metrics.data() + (size_t)i * nstate;
                            const int key = c0 * v[0] + c1 * v[1] + c2 * v[2] + c3 * v[3] - lo;
                            inj = !((seen[key >> 6] >> (key & 63)) & 1ull);
                            seen[key >> 6] |= 1ull << (key & 63);
                        }
                        if (!inj) continue;
                        ctx->kc[0] = c0; ctx->kc[1] = c1; ctx->kc[2] = c2; ctx->kc[3] = c3;
                        ctx->kbias = -lo;
                        ctx->linkey_ok = true;
                        break;
                    }
        }
        if (ctx->linkey_ok) {
            std::vector<uint16_t> dst(256, (uint16_t)0xFFFF);
            for (uint32_t i = 0; i < S; ++i) {
                const uint8_t* v = ctx->h_metrics.data() + (size_t)i * nstate;
                dst[ctx->kc[0] * v[0] + ctx->kc[1] * v[1] + ctx->kc[2] * v[2] + ctx->kc[3] * v[3] + ctx->kbias] = (uint16_t)i;
            }
            CK(ctx->d_dstate2.reserve(256 * 2));
            CK(h2d(ctx, ctx->d_dstate2.p, dst.data(), 256 * 2));
        }
    }
    // ---- perfect hash for m = 3 / 4, n = 2 (two-trials-per-thread ACS kernels, mvd_detect3p.cuh): hash, displace.
    // w0 = k0 | k1 << 16 (and w1 = k2 | k3 << 16 for m = 4) with the OFFSET-INVARIANT digits
    // k_q = sum_i 16^i (D[4q+i] - D[0] + 8)   (metrics <= 7, so every digit is in 1..15): the kernel forms the same words
    // from un-normalised metrics; bucket = (w0 C1 + w1 C3) >> bshift, slot = (((w0 C2 + w1 C4) >> h2shift) + disp[bucket]) &
    // (slots - 1).  Buckets are placed largest first, each with the smallest displacement that lands all its keys
    // on free slots (single keys take the next free slot).  m = 3 (tables in shared memory) first tries the smallest
    // power of two >= S / 0.85 slots, then twice that.
    ctx->ph_slots = 0;
    if (ctx->closed && (ctx->m == 3 || ctx->m == 4) && ctx->n == 2 && ctx->max_metric <= 7 && (ctx->m == 4 || S <= 1024) && S <= (1u << 20)) {
      uint32_t slots0 = 256;
      if (ctx->m == 3) { while (slots0 * 0.85 < S) slots0 <<= 1; } else { while (slots0 < 2 * S) slots0 <<= 1; }
      for (uint32_t slots = slots0; slots <= (ctx->m == 3 ? 2 * slots0 : slots0) && !ctx->ph_slots; slots <<= 1) {
        uint32_t nb = 256;
        if (ctx->m == 4) while (nb < S / 4) nb <<= 1;
        uint32_t lb = 0;
        while ((1u << lb) < nb) ++lb;
        const uint32_t bshift = 32 - lb, h2shift = ctx->m == 3 ? 21 : 11;
        std::vector<uint32_t> h1(S), h2(S), w0(S), w1(S, 0u);
        std::vector<uint32_t> bcount(nb + 1, 0u);
        for (uint32_t i = 0; i < S; ++i) {
            const uint8_t* v = ctx->h_metrics.data() + (size_t)i * nstate;
            uint32_t w[2] = {0u, 0u};
            for (int q = 0; q < nstate / 4; ++q) {
                uint32_t kq = 0;
                for (int i = 0; i < 4; ++i) kq |= (uint32_t)((int)v[4 * q + i] - (int)v[0] + 8) << (4 * i);
                w[q >> 1] |= kq << (16 * (q & 1));
            }
            w0[i] = w[0];
            w1[i] = w[1];
            h1[i] = (w[0] * 0x9E3779B1u + w[1] * 0xC2B2AE35u) >> bshift;
            ++bcount[h1[i] + 1];
        }
        for (uint32_t b = 0; b < nb; ++b) bcount[b + 1] += bcount[b];          // bucket b = members[bcount[b] .. bcount[b+1])
        std::vector<uint32_t> members(S), fill(bcount.begin(), bcount.end() - 1);
        for (uint32_t i = 0; i < S; ++i) members[fill[h1[i]]++] = i;
        std::vector<uint32_t> order(nb);
        for (uint32_t b = 0; b < nb; ++b) order[b] = b;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            return bcount[a + 1] - bcount[a] > bcount[b + 1] - bcount[b];
        });
        std::vector<uint32_t> disp(nb, 0u), table(slots, MVD_EMPTY);
        bool built = false;
        uint32_t c2 = 0x85EBCA6Bu, c4 = 0x27D4EB2Fu;
        // two keys of one bucket with the same second hash cannot be separated by a displacement: try the next
        // pair of (odd) multipliers -- about every second attempt succeeds at these sizes
        for (uint32_t attempt = 0; attempt < 64 && !built; ++attempt, c2 += 0xC6574B56u, c4 += 0x3C6EF372u) {
            for (uint32_t i = 0; i < S; ++i) h2[i] = (w0[i] * c2 + w1[i] * c4) >> h2shift;
            std::fill(disp.begin(), disp.end(), 0u);
            std::fill(table.begin(), table.end(), MVD_EMPTY);
            built = true;
            uint32_t next_free = 0;
            for (uint32_t oi = 0; oi < nb && built; ++oi) {
                const uint32_t b = order[oi], lo = bcount[b], hi = bcount[b + 1];
                if (lo == hi) break;
                if (hi - lo == 1) {                                            // any free slot will do
                    while (next_free < slots && table[next_free] != MVD_EMPTY) ++next_free;
                    if (next_free == slots) { built = false; break; }
                    disp[b] = (next_free - h2[members[lo]]) & (slots - 1);
                    table[next_free] = members[lo] * (uint32_t)R;
                    continue;
                }
                bool placed = false;
                for (uint32_t d = 0; d < slots && !placed; ++d) {
                    bool ok2 = true;
                    for (uint32_t a = lo; a < hi && ok2; ++a) {
                        const uint32_t sa = (h2[members[a]] + d) & (slots - 1);
                        ok2 = table[sa] == MVD_EMPTY;
                        for (uint32_t c = lo; c < a && ok2; ++c) ok2 = sa != ((h2[members[c]] + d) & (slots - 1));
                    }
                    if (ok2) {
                        for (uint32_t a = lo; a < hi; ++a) table[(h2[members[a]] + d) & (slots - 1)] = members[a] * (uint32_t)R;
                        disp[b] = d;
                        placed = true;
                    }
                }
                built = placed;
            }
            if (built) {
                ctx->ph_c2 = c2;
                ctx->ph_c4 = c4;
            }
        }
        if (built) {
            CK(ctx->d_phd.reserve((size_t)nb * 4));
            CK(ctx->d_pht.reserve((size_t)slots * 4));
            CK(h2d(ctx, ctx->d_phd.p, disp.data(), (size_t)nb * 4));
            CK(h2d(ctx, ctx->d_pht.p, table.data(), (size_t)slots * 4));
            CK(cudaStreamSynchronize(ctx->stream));
            ctx->ph_slots = slots;
            ctx->ph_bshift = bshift;
            ctx->ph_nb = nb;
            ctx->ph_slot0 = (h2[0] + disp[h1[0]]) & (slots - 1);   // Markov state 0 = the all-zero vector
        }
      }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_states = true;
    ctx->ntables = 0;
    ctx->have_llslot = false;
    return MVD_OK;
}

struct LaunchOut {
    uint64_t* tallies = nullptr;
    double* logp = nullptr;
    void* d_tallies = nullptr;
    uint64_t* counts = nullptr;
    uint32_t burn = 0;
    uint32_t* trace_idx = nullptr;
    uint8_t* trace_met = nullptr;
    uint64_t* hashes = nullptr;
    uint8_t* final_met = nullptr;
};

// Shared-memory plan of the fast detection kernels; returns false if they do not apply.
bool plan_det2(const mvd_ctx* ctx, int engine, int* lk_out, int* lls_out, FastPlan* fp, size_t* smem_out, bool* gt_out,
               bool allow_big = false, bool* big_out = nullptr) {
    *gt_out = false;
    if (big_out) *big_out = false;
    // n = 2: every engine; n = 3 (rate 1/3): the NEXT-walk engines, which take any number of received words
    if (!(ctx->n == 2 || (ctx->n == 3 && engine == MVD_ENGINE_FSM)) || !ctx->closed) return false;
    const int m = ctx->m, nstate = 1 << m, NP = nstate / 2, R = 1 << ctx->n;
    const size_t SR = (size_t)ctx->S * R;
    int lk;
    if (engine == MVD_ENGINE_FSM && ctx->tref_packed && !ctx->no_fsm1 && 128 + (SR << 4) + 64 <= ctx->prop.sharedMemPerBlockOptin) {
        // one-load NEXT walk: [masks][S*R entries x 2^(LLS-4) copies x 16 B].  Eight copies make every row read
        // conflict-free; when that leaves one block per SM (S = 435: 222 KB) fewer copies and more resident blocks win
        // (measured at S = 435: 8 / 4 / 2 / 1 copies = 7.06e11 / 8.46e11 / 8.44e11 / 8.17e11 steps/s)
        int lls = 7;
        const char* force = getenv("MVD_FSM1_LLS");                 // experiments only
        if (force && *force >= '4' && *force <= '7') lls = *force - '0';
        else {
            while (lls > 4 && 128 + (SR << lls) + 64 > 75 * 1024) --lls;              // three blocks of 512 threads per SM
            // ... unless the launch is large enough for TWO blocks of 768 threads per SM (as many threads) and a table with twice
            // the copies fits twice: half the bank conflicts of the row reads (measured at S = 435: 8.5e11 -> 9.0e11 steps/s)
            if (allow_big && big_out && ctx->n == 2 && lls < 7 && lls >= 4 && 128 + (SR << (lls + 1)) + 64 <= 112 * 1024) {
                *big_out = true;
                ++lls;
            }
        }
        while (lls > 4 && 128 + (SR << lls) + 64 > ctx->prop.sharedMemPerBlockOptin) --lls;
        fp->off_tb = 0;
        fp->off_bm = fp->off_st = 128;
        fp->off_ll = 128;
        fp->key_mul = 0;
        fp->nkeys = 0;
        fp->dstate = nullptr;
        fp->tcode = ctx->d_tcode.as<uint32_t>();
        fp->tref_unit = ctx->tref_unit;
        *lk_out = LK_FSM1;
        *lls_out = lls;
        *smem_out = 128 + (SR << lls);
        return true;
    }
    if (engine == MVD_ENGINE_FSM) lk = LK_FSM;
    else if (ctx->nkeys) lk = LK_DIRECT;
    else if ((m == 2 || m == 3) && ctx->acs_ok) lk = LK_HASH;
    else lk = -1;                                   // no shared-memory variant: maybe the global-table one below
    const size_t budget2 = 110 * 1024, smem_max = ctx->prop.sharedMemPerBlockOptin;   // 2 blocks / SM if possible
    for (int pass = 0; pass < 2 && lk >= 0; ++pass) {
        for (int lls = 7; lls >= 4; --lls) {
            size_t off = 128, st_bytes;                 // 32 threshold-bit masks
            fp->off_tb = 0;
            fp->off_bm = 128;
            if (lk != LK_FSM) off += (size_t)(NP >= 2 ? NP / 2 : 1) * ((size_t)R << lls);   // planes x rows x copies
            off = (off + 15) & ~(size_t)15;
            fp->off_st = (uint32_t)off;
            if (lk == LK_FSM) st_bytes = lls == 7 ? SR * 128 : SR * 4;
            else if (lk == LK_DIRECT) st_bytes = (size_t)ctx->nkeys * 128;
            else st_bytes = (size_t)ctx->hcap * 4 * (1 + (nstate + 7) / 8);
            off = (off + st_bytes + 15) & ~(size_t)15;
            fp->off_ll = (uint32_t)off;
            off += SR << lls;
            if (lk == LK_DIRECT && fp->off_st + ((size_t)ctx->nkeys << 7) + 128 >= 65536) continue;
            if (off + 64 <= (pass == 0 ? budget2 : smem_max)) {
                const int b = m == 1 ? 4 : 2;
                fp->key_mul = ((1u << (16 + b)) + 1u) << 7;
                fp->nkeys = ctx->nkeys;
                fp->dstate = ctx->d_dstate.as<uint16_t>();
                *lk_out = lk;
                *lls_out = lls;
                *smem_out = off;
                return true;
            }
        }
    }
    // tables too large for shared memory: leave them in global memory (L2), stage only masks + branch metrics
    fp->off_tb = 0;
    fp->off_bm = fp->off_st = fp->off_ll = 128;
    fp->key_mul = 0;
    fp->nkeys = 0;
    fp->dstate = nullptr;
    fp->tcode = ctx->d_tcode.as<uint32_t>();
    fp->tref_unit = ctx->tref_unit;
    fp->gfsm1 = ctx->d_gfsm1.as<uint4>();
    if (engine == MVD_ENGINE_FSM && ctx->have_gfsm1) {
        *lk_out = LK_FSM1;
        *lls_out = 4;
        *smem_out = 128;
        *gt_out = true;
        return true;
    }
    if (engine == MVD_ENGINE_ACS && (m == 3 || m == 4) && ctx->acs_ok) {
        *lk_out = LK_HASH;
        *lls_out = 4;
        *smem_out = 128 + (size_t)(NP / 2) * ((size_t)R << 4);
        *gt_out = true;
        return true;
    }
    return false;
}

// Wait for the asynchronous detection launches in flight, add up their kernel times, read and clear the error flag.
int drain_async(mvd_ctx* ctx) {
    if (!ctx->async_pending) return MVD_OK;
    int herr = 0;
    CK(d2h(ctx, &herr, ctx->d_err.p, sizeof(int)));
    CK(cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < ctx->async_pending; ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->async_ev[i].first, ctx->async_ev[i].second));
        ctx->async_ms_sum += ms;
        ctx->last_ms = ms;
    }
    ctx->async_launches += ctx->async_pending;
    ctx->async_pending = 0;
    if (herr) {
        CK(cudaMemsetAsync(ctx->d_err.p, 0, sizeof(int), ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (herr & 1) return fail(ctx, MVD_E_UNKNOWN_STATE, "a relative-metric vector was not in the state table (KeyError)");
    if (herr & 2) return fail(ctx, MVD_E_UNKNOWN_STATE, "relative metric exceeded 15 (hash key overflow)");
    if (herr & 4) return fail(ctx, MVD_E_CUDA, "pair kernel: the dynamic shared window does not start where the host's layout assumed");
    return MVD_OK;
}

int run(mvd_ctx* ctx, int mode, int engine, const mvd_src* src, const mvd_segment* segs, uint32_t nsegs,
        const LaunchOut& out) {
    if (!ctx) return MVD_E_INVALID;
    if (!src || !segs || nsegs == 0) return fail(ctx, MVD_E_INVALID, "null source / segments");
    if (!ctx->have_code) return fail(ctx, MVD_E_STATE, "mvd_set_code has not been called");
    if (mode != MODE_HASH && !ctx->have_states) return fail(ctx, MVD_E_STATE, "no Markov state table (mvd_set_states)");
    if (mode == MODE_DETECT && ctx->ntables == 0) return fail(ctx, MVD_E_STATE, "no log-likelihood tables (mvd_set_loglik)");
    if (src->mode != MVD_SRC_PHILOX && src->mode != MVD_SRC_BITSTREAM) return fail(ctx, MVD_E_INVALID, "bad source mode");
    CK(cudaSetDevice(ctx->device));
    const int n = ctx->n, m = ctx->m, R = 1 << n, nstate = 1 << m;
    const uint32_t S = ctx->S;
    const uint32_t SR = mode == MODE_HASH ? 0u : S * (uint32_t)R;

    if (engine == MVD_ENGINE_AUTO) engine = (mode == MODE_HASH) ? MVD_ENGINE_ACS : MVD_ENGINE_FSM;
    if (mode == MODE_HASH) engine = MVD_ENGINE_ACS;
    if (ctx->table_code) {
        if (engine != MVD_ENGINE_FSM) return fail(ctx, MVD_E_UNSUPPORTED, "a code given as tables runs on the Markov-state walk (MVD_ENGINE_FSM / AUTO) only");
        if (ctx->nenc == 0) return fail(ctx, MVD_E_STATE, "no encoder tables (mvd_set_encoders)");
        if (!ctx->closed) return fail(ctx, MVD_E_INVALID, "the state table is not closed under Eq. 4-5 on the given trellis");
    }
    if (engine != MVD_ENGINE_ACS && engine != MVD_ENGINE_FSM) return fail(ctx, MVD_E_INVALID, "bad engine %d", engine);
    if (engine == MVD_ENGINE_ACS && mode != MODE_HASH && !ctx->acs_ok)
        return fail(ctx, MVD_E_UNSUPPORTED, "ACS engine needs relative metrics <= 15");

    // ---- kernel choice: the fast detection kernels when they apply, else the generic ones
    FastPlan fplan{};
    int det2_lk = -1, det2_lls = 0;
    size_t det2_smem = 0;
    bool det2_gt = false;
    uint64_t all_trials = 0;
    for (uint32_t i = 0; i < nsegs; ++i) all_trials += segs[i].trial_end >= segs[i].trial_begin ? segs[i].trial_end - segs[i].trial_begin : 0;
    const uint64_t sms = (uint64_t)ctx->prop.multiProcessorCount;
    // blocks of 768 threads (one or two per SM around larger shared tables) pay off when the LONG segments of the launch fill the
    // GPU by themselves: in an N sweep at few trials per point the tail is a handful of long blocks, and fewer, larger blocks
    // lengthen it (m = 3 Pd-vs-N sweep at 10^4 trials per point: 10.5 ms with blocks of 512, 13.0 ms with blocks of 768)
    uint32_t longest = 0;
    for (uint32_t i = 0; i < nsegs; ++i) longest = std::max(longest, segs[i].N);
    uint64_t long_trials = 0;
    for (uint32_t i = 0; i < nsegs; ++i)
        if (2ull * segs[i].N >= longest && segs[i].trial_end >= segs[i].trial_begin) long_trials += segs[i].trial_end - segs[i].trial_begin;
    const bool big_ok = long_trials >= 2ull * DET2_BIG_BLOCK * sms && !getenv("MVD_NO_BIG_BLOCK");
    bool det2_big = false;
    const bool fast = mode == MODE_DETECT && !ctx->force_generic && !ctx->table_code &&
                      plan_det2(ctx, engine, &det2_lk, &det2_lls, &fplan, &det2_smem, &det2_gt, big_ok, &det2_big);
    // two trials per thread: ACS engine, m = 2, direct table, log rows replicated 8 x
    // (layout of detect2p_kernel: straggler queues, masks, branch metrics, log rows, then the state table at a 32 KB-aligned
    // absolute shared address; the dynamic window starts after the 1 KB the system reserves)
    const size_t pair_sbase = 1024;
    const size_t pair_st = (pair_sbase + DET2P_QUEUES + 128 + 2048 + 4096 + ((size_t)ctx->S * 4 << 7) + 32767) & ~(size_t)32767;
    const size_t pair_smem = pair_st + 32768 - pair_sbase;
    // pairing halves the thread count: only when the GPU stays full (3 blocks of 256 pair-threads per SM)
    const bool pair = fast && !ctx->no_pair && det2_lk == LK_DIRECT && ctx->m == 2 && ctx->linkey_ok && pair_smem <= 110 * 1024 &&
                      (ctx->force_pair || all_trials >= 2ull * DET2P_BLOCK * 3ull * sms);
    // m = 3: two trials per thread with the perfect-hash lookup
    // (layout of detect3p_kernel, m = 3: straggler queues of 24 warps, masks, branch table, lane-replicated displacements at a
    // 32 KB-aligned and the 16-copy slot table at a (slots x 64)-aligned absolute shared address, log rows in 4 copies; the dynamic
    // window starts after the 1 KB the system reserves; one block of 768 pair-threads per SM)
    const size_t pair3_tb = (size_t)ctx->ph_slots * 64;
    const size_t pair3_D = (((pair_sbase + DET3P_QUEUES + 255) & ~(size_t)255) + 256 + 256 + 32767) & ~(size_t)32767;
    const size_t pair3_T = pair3_tb ? (pair3_D + 32768 + pair3_tb - 1) & ~(pair3_tb - 1) : 0;
    const size_t pair3_smem = pair3_T + pair3_tb + ((size_t)ctx->S * 4 << 6) - pair_sbase;
    const bool pair3s = fast && !pair && !ctx->no_pair && engine == MVD_ENGINE_ACS && det2_lk == LK_HASH && !det2_gt && ctx->m == 3 &&
                        ctx->ph_slots && pair3_smem <= 227 * 1024 &&
                        (ctx->force_pair || all_trials >= 2ull * DET3P_BLOCK * sms);
    // m = 4: the same kernel with displacements, slots and log rows in global memory
    const bool pair4 = fast && !pair && !ctx->no_pair && engine == MVD_ENGINE_ACS && det2_lk == LK_HASH && det2_gt && ctx->m == 4 &&
                       ctx->ph_slots && ctx->have_llslot && (ctx->force_pair || all_trials >= 2ull * DET2P_BLOCK * 2ull * sms);
    const bool pair3 = pair3s || pair4;
    // few trials: smaller blocks so that every SM gets work (the kernels read blockDim.x)
    uint32_t threads = pair3s ? DET3P_BLOCK : (pair || pair3) ? DET2P_BLOCK : det2_big ? DET2_BIG_BLOCK : DET2_BLOCK;
    if (fast && !(det2_big && !pair && !pair3)) {
        const uint64_t per_thread = (pair || pair3) ? 2 : 1;
        const uint32_t least = pair3s ? 96 : 64;             // 768 halves to 384, 192, 96: whole warps
        while (threads > least && (all_trials + threads * per_thread - 1) / (threads * per_thread) < (pair3s ? 1 : 4) * sms) threads >>= 1;
    }
    const uint32_t block = fast ? threads * ((pair || pair3) ? 2u : 1u) : MVD_BLOCK;

    // ---- segments
    std::vector<DevSeg> ds(nsegs);
    uint64_t blocks = 0, trials = 0, need_words = 0;
    for (uint32_t i = 0; i < nsegs; ++i) {
        const mvd_segment& s = segs[i];
        if (s.trial_end < s.trial_begin) return fail(ctx, MVD_E_INVALID, "segment %u: trial_end < trial_begin", i);
        if (mode == MODE_DETECT && s.table >= ctx->ntables) return fail(ctx, MVD_E_INVALID, "segment %u: table %u >= %u", i, s.table, ctx->ntables);
        if (s.decide > 1) return fail(ctx, MVD_E_INVALID, "segment %u: decide must be 0 or 1", i);
        // MVD-PHILOX-2 addresses a call as (32-step block << 6) | slot in one 32-bit counter word: block < 2^26
        if (src->mode == MVD_SRC_PHILOX && s.N >= 0x80000000u)
            return fail(ctx, MVD_E_INVALID, "segment %u: N = %u >= 2^31 overflows the position-addressed Philox counter", i, s.N);
        DevSeg& d = ds[i];
        d.N = s.N;
        d.threshold = s.threshold;
        d.stream = s.stream;
        d.table = s.table;
        for (int j = 0; j < MVD_MAX_N; ++j) {
            d.enc_taps[j] = j < n ? s.enc_taps[j] : 0u;
            if (!ctx->table_code && j < n && (s.enc_taps[j] >> (m + 1))) return fail(ctx, MVD_E_INVALID, "segment %u: encoder tap beyond memory m=%d", i, m);
        }
        if (ctx->table_code) {
            if (s.enc_taps[0] >= ctx->nenc) return fail(ctx, MVD_E_INVALID, "segment %u: encoder index %u >= %u (mvd_set_encoders)", i, s.enc_taps[0], ctx->nenc);
            d.enc_taps[0] = s.enc_taps[0];
        }
        d.decide = s.decide;
        d.random_input = s.random_input ? 1u : 0u;
        d.dmin = s.threshold ? (uint32_t)__builtin_ctz(s.threshold) : 32u;
        const uint64_t ntr = s.trial_end - s.trial_begin;
        if (blocks > 0x7FFFFFFFull) return fail(ctx, MVD_E_INVALID, "too many trials in one call");
        d.block_begin = (uint32_t)blocks;
        d.trial_begin = s.trial_begin;
        d.trial_end = s.trial_end;
        d.bits_offset = s.bits_offset;
        d.out_offset = trials;
        blocks += (ntr + block - 1) / block;
        trials += ntr;
        if (src->mode == MVD_SRC_BITSTREAM) {
            const uint64_t nsb = ((uint64_t)s.N + 127) / 128;
            need_words = std::max<uint64_t>(need_words, s.bits_offset + nsb * (uint64_t)(ctx->k + n) * ntr);
        }
    }
    if (blocks > 0x7FFFFFFFull) return fail(ctx, MVD_E_INVALID, "too many trials in one call");
    if ((mode == MODE_TRACE || mode == MODE_HASH) && nsegs != 1) return fail(ctx, MVD_E_INVALID, "trace/hash take exactly one segment");

    Params P{};
    P.n = n;
    P.m = m;
    P.k = ctx->k;
    P.enc_tab = ctx->table_code ? ctx->d_enc.as<uint16_t>() : nullptr;
    P.R = R;
    P.nstate = nstate;
    P.S = S;
    P.SR = SR;
    P.src_mode = src->mode;
    {
        uint32_t k0 = (uint32_t)src->seed, k1 = (uint32_t)(src->seed >> 32);
        for (int r = 0; r < 10; ++r) {
            P.rk0[r] = k0;
            P.rk1[r] = k1;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
    }
    if (src->mode == MVD_SRC_BITSTREAM) {
        if (!src->bits) return fail(ctx, MVD_E_INVALID, "bitstream source without bits");
        if (src->bits_words < need_words) return fail(ctx, MVD_E_INVALID, "bitstream too short: %llu words, need %llu",
                                                      (unsigned long long)src->bits_words, (unsigned long long)need_words);
        if (src->bits_on_device) {
            P.bits = reinterpret_cast<const uint4*>(src->bits);
        } else {
            CK(ctx->d_bits.reserve((size_t)need_words * 16));
            CK(h2d(ctx, ctx->d_bits.p, src->bits, (size_t)need_words * 16));
            P.bits = ctx->d_bits.as<uint4>();
        }
    }
    CK(ctx->d_segs.reserve(sizeof(DevSeg) * nsegs));
    CK(h2d(ctx, ctx->d_segs.p, ds.data(), sizeof(DevSeg) * nsegs));
    P.segs = ctx->d_segs.as<DevSeg>();
    P.nsegs = nsegs;
    P.nxt = ctx->d_nxt.as<uint32_t>();
    P.ll = ctx->d_ll.as<double2>();
    P.bm = ctx->d_bm.as<uint32_t>();
    P.bm_antipodal = !ctx->no_antipodal;
    for (int j = 0; j < ctx->n; ++j)
        if (!(ctx->dec_taps[j] & 1u) || !((ctx->dec_taps[j] >> ctx->m) & 1u)) P.bm_antipodal = 0;
    P.hkeys = ctx->d_hkeys.as<uint32_t>();
    P.hvals = ctx->d_hvals.as<uint32_t>();
    P.hcap = ctx->hcap;
    P.burn = out.burn;
    // asynchronous launch: device tallies only, nothing for the host to read; the error flag accumulates until the drain
    const bool async = ctx->async_detect && mode == MODE_DETECT && out.d_tallies && !out.tallies && !out.logp;
    if (ctx->async_pending && (!async || ctx->async_pending >= 64)) {
        const int rc = drain_async(ctx);
        if (rc != MVD_OK) return rc;
    }
    CK(ctx->d_err.reserve(sizeof(int)));
    if (!ctx->async_pending) CK(cudaMemsetAsync(ctx->d_err.p, 0, sizeof(int), ctx->stream));
    P.error_flag = ctx->d_err.as<int>();
    P.fp = fplan;
    P.fp.dstate2 = ctx->d_dstate2.as<uint16_t>();
    for (int i = 0; i < 4; ++i) P.fp.kc[i] = (uint32_t)ctx->kc[i];
    P.fp.kcb = ((uint32_t)ctx->kbias << 7) * 0x00010001u;
    P.fp.ph_d = ctx->d_phd.as<uint32_t>();
    P.fp.ph_t = ctx->d_pht.as<uint32_t>();
    P.fp.ph_slots = ctx->ph_slots;
    P.fp.ph_bshift = ctx->ph_bshift;
    P.fp.ph_c2 = ctx->ph_c2;
    P.fp.ph_c4 = ctx->ph_c4;
    P.fp.ph_nb = ctx->ph_nb;
    P.fp.ph_slot0 = ctx->ph_slot0;
    P.fp.kq[0] = 16u; P.fp.kq[1] = 256u; P.fp.kq[2] = 4096u; P.fp.kq[3] = 0u - 4369u; P.fp.kq[4] = 1u; P.fp.kq[5] = 0u - 4368u;
    P.fp.ll_slot = ctx->d_llslot.as<double2>();

    // ---- outputs
    if (mode == MODE_DETECT) {
        CK(ctx->d_tallies.reserve(8 * (size_t)nsegs));
        CK(cudaMemsetAsync(ctx->d_tallies.p, 0, 8 * (size_t)nsegs, ctx->stream));
        P.tallies = ctx->d_tallies.as<unsigned long long>();
        P.tallies2 = nullptr;               // d_tallies receives a device-to-device copy after the launch (below)
        if (out.logp) {
            CK(ctx->d_logp.reserve(16 * (size_t)trials));
            P.logp = ctx->d_logp.as<double>();
        }
    } else if (mode == MODE_LEARN) {
        CK(ctx->d_counts.reserve(8 * (size_t)nsegs * SR));
        CK(cudaMemsetAsync(ctx->d_counts.p, 0, 8 * (size_t)nsegs * SR, ctx->stream));
        P.counts = ctx->d_counts.as<unsigned long long>();
    } else if (mode == MODE_TRACE) {
        const size_t cells = (size_t)trials * ((size_t)segs[0].N + 1);
        CK(ctx->d_trace_idx.reserve(4 * cells));
        P.trace_idx = ctx->d_trace_idx.as<uint32_t>();
        if (out.trace_met && engine == MVD_ENGINE_ACS) {
            CK(ctx->d_trace_met.reserve(cells * nstate));
            P.trace_met = ctx->d_trace_met.as<uint8_t>();
        }
    } else if (mode == MODE_HASH) {
        CK(ctx->d_hashes.reserve(8 * (size_t)trials));
        P.hashes = ctx->d_hashes.as<unsigned long long>();
        if (out.final_met) {
            CK(ctx->d_final.reserve((size_t)trials * nstate));
            P.final_met = ctx->d_final.as<uint8_t>();
        }
    }

    // ---- shared memory plan
    const size_t smem_max = ctx->prop.sharedMemPerBlockOptin;
    size_t smem = 0;
    bool in_smem = false;
    if (engine == MVD_ENGINE_FSM) {
        size_t need = (size_t)SR * 2;
        if (mode == MODE_DETECT) need += (size_t)SR * 16;
        if (mode == MODE_LEARN) need += (size_t)SR * 4;
        bool learn_fits32 = mode != MODE_LEARN || (uint64_t)MVD_BLOCK * (uint64_t)segs[0].N < 0xFFFFFFFFull;
        for (uint32_t i = 0; i < nsegs && learn_fits32 && mode == MODE_LEARN; ++i)
            learn_fits32 = (uint64_t)MVD_BLOCK * (uint64_t)segs[i].N < 0xFFFFFFFFull;
        in_smem = (uint64_t)SR <= 65535ull && need + 64 <= smem_max && learn_fits32;
        smem = in_smem ? need : 0;
    } else {
        const int NP = nstate / 2, KW = (nstate + 7) / 8;
        size_t base = (((size_t)R * 2 * NP * 4) + 15) & ~(size_t)15;
        size_t need = base;
        if (mode == MODE_DETECT) need += (size_t)SR * 16;
        if (mode == MODE_LEARN) need += (size_t)SR * 4;
        need += (size_t)ctx->hcap * 4 * (1 + KW);
        bool learn_fits32 = true;
        for (uint32_t i = 0; i < nsegs && mode == MODE_LEARN; ++i)
            learn_fits32 = learn_fits32 && (uint64_t)MVD_BLOCK * (uint64_t)segs[i].N < 0xFFFFFFFFull;
        in_smem = mode != MODE_HASH && need + 64 <= smem_max && learn_fits32;
        smem = in_smem ? need : base;
    }
    P.tables_in_smem = in_smem ? 1 : 0;

    if (blocks == 0) {
        if (mode == MODE_DETECT && out.d_tallies) {
            CK(cudaMemsetAsync(out.d_tallies, 0, 8 * (size_t)nsegs, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        if (out.tallies) memset(out.tallies, 0, 8 * (size_t)nsegs);
        if (out.counts) memset(out.counts, 0, 8 * (size_t)nsegs * SR);
        return MVD_OK;
    }
    // few long trials (Pd-vs-N sweeps): split them along the time axis when that fills the GPU better
    bool split = mode == MODE_DETECT && engine == MVD_ENGINE_FSM && !ctx->force_generic && !ctx->table_code && ctx->split_mode != 2 &&
                 src->mode == MVD_SRC_PHILOX && ctx->closed && trials > 0 && trials <= 0x7FFFFFFFull;
    if (split) {
        unsigned long long steps = 0, ew = 0, work = 0;
        for (uint32_t i = 0; i < nsegs; ++i) {
            const unsigned long long ntr = ds[i].trial_end - ds[i].trial_begin;
            if (ds[i].N == 0) split = false;
            steps += ntr * ds[i].N;
            ew += ntr * (((unsigned long long)ds[i].N + 3ull) & ~3ull);
            work += ntr * (((unsigned long long)ds[i].N + 255ull) / 256ull);      // chunks at the smallest chunk size
        }
        split = split && ew * 4ull <= (6ull << 30) && work <= 0x7FFFFFFFull * (unsigned long long)SPLIT_BLOCK;
        if (ctx->split_mode == 0) {
            // automatic: compare the two paths with measured rates (B200, m = 2).  One thread per trial: the chain of a
            // trial costs ~59 ns per step until the GPU is full (~1e12 steps/s); split: walk + partial sums are work-bound at
            // ~5.5e11 steps/s (warm-up included), the in-order part costs ~2.2 ns per step of the longest trial (one record
            // per 128 steps plus the binade crossings), four launches ~50 us.
            uint32_t maxN = 0;
            for (uint32_t i = 0; i < nsegs; ++i) maxN = std::max(maxN, ds[i].N);
            const double t_plain = std::max((double)maxN * 59e-9, (double)steps / 1.0e12);
            const double t_split = (double)steps / 5.5e11 + (double)maxN * 2.2e-9 + 50e-6;
            split = split && maxN >= 2048u && t_split < 0.8 * t_plain;
        }
    }
    const dim3 grid((unsigned)blocks);
    const bool n2 = (n == 2);
    cudaError_t le = cudaErrorInvalidValue;
    const bool go_async = async && !split;
    cudaEvent_t ev0 = ctx->ev0, ev1 = ctx->ev1;
    if (go_async) {
        if (ctx->async_ev.size() <= ctx->async_pending) {
            cudaEvent_t a = nullptr, b = nullptr;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            ctx->async_ev.emplace_back(a, b);
        }
        ev0 = ctx->async_ev[ctx->async_pending].first;
        ev1 = ctx->async_ev[ctx->async_pending].second;
    }
    CK(cudaEventRecord(ev0, ctx->stream));
    bool plearn = mode == MODE_LEARN && engine == MVD_ENGINE_FSM && !ctx->force_generic && !ctx->table_code && src->mode == MVD_SRC_PHILOX;
    uint32_t maxL = 0;
    for (uint32_t i = 0; i < nsegs && plearn; ++i) {
        plearn = (segs[i].trial_end - segs[i].trial_begin) == 1;
        maxL = std::max(maxL, segs[i].N);
    }
    if (plearn && maxL > 0) {
        // chunk-parallel learning chains (mvd_learn2.cuh): speculate, check, fix
        LearnParams LP{};
        // steps per chunk: the longest of 128 / 256 / 512 / 1024 that still leaves every SM two rounds of 1 024 chunk-threads
        // (the walk is bound by the latency of its dependent table reads; at S = 150 743 a 384-step warm-up in front of
        // 128-step chunks quadruples the steps walked, in front of 512-step chunks it adds 75 %)
        uint64_t sumL = 0;
        for (uint32_t i = 0; i < nsegs; ++i) sumL += segs[i].N;
        LP.chunk = LEARN_CH;
        const char* fch = getenv("MVD_LEARN_CHUNK");               // tests / experiments: 128, 256, 512 or 1024 (identical counts)
        const uint32_t forced = fch ? (uint32_t)atoi(fch) : 0u;
        if (forced == 128u || forced == 256u || forced == 512u || forced == 1024u) LP.chunk = forced;
        else while (LP.chunk < 1024u && sumL / (2ull * LP.chunk) >= 2ull * 1024ull * (uint64_t)ctx->prop.multiProcessorCount) LP.chunk *= 2u;
        LP.nchunks = (maxL + LP.chunk - 1) / LP.chunk;
        LP.warm = ctx->learn_warm;
        const size_t cells = (size_t)nsegs * LP.nchunks;
        CK(ctx->d_lspec.reserve(cells * 4));
        CK(ctx->d_lend.reserve(cells * 4));
        CK(ctx->d_ldirty.reserve((size_t)nsegs * 4));
        CK(cudaMemsetAsync(ctx->d_ldirty.p, 0, (size_t)nsegs * 4, ctx->stream));
        LP.spec_start = ctx->d_lspec.as<uint32_t>();
        LP.end = ctx->d_lend.as<uint32_t>();
        LP.ndirty = ctx->d_ldirty.as<uint32_t>();
        const size_t lsmem = (size_t)SR * 8;
        const bool lin = lsmem <= 96 * 1024;
        LP.nxt_in_smem = lin ? 1 : 0;
        le = mvd_launch_learn(lin, lsmem, nsegs, ctx->stream, P, LP);
        ctx->launches += 2;
        ctx->last_fast = 1024;
    } else if (split) {
        // few long trials: one thread per (trial, chunk) walks the states, one per trial adds the log-likelihoods
        // in step order (mvd_split.cuh)
        std::vector<unsigned long long> meta(3 * (size_t)nsegs + 1);
        unsigned long long w = 0, ew = 0, subs = 0;
        // steps per chunk: the largest of 1024 / 512 / 256 that still gives the walk 1.5 x the threads the GPU holds at
        // 64 registers (the warm-up before every chunk is overhead: 128 steps per chunk)
        uint32_t chunk = SPLIT_CH_MAX;
        for (; chunk > 256u; chunk >>= 1) {
            unsigned long long items = 0;
            for (uint32_t i = 0; i < nsegs; ++i) items += (ds[i].trial_end - ds[i].trial_begin) * (((unsigned long long)ds[i].N + chunk - 1ull) / chunk);
            if (items * 2ull >= 3ull * 1024ull * sms) break;
        }
        if (ctx->split_chunk) chunk = ctx->split_chunk;
        const int eb = SR <= 256u ? 1 : (SR <= 65536u ? 2 : 4);
        const unsigned long long spg = 16 / eb;
        for (uint32_t i = 0; i < nsegs; ++i) {
            const unsigned long long ntr = ds[i].trial_end - ds[i].trial_begin;
            meta[i] = w;
            meta[nsegs + 1 + i] = ew;
            meta[2 * (size_t)nsegs + 1 + i] = subs;
            w += ntr * (((unsigned long long)ds[i].N + chunk - 1ull) / chunk);
            ew += ntr * (((unsigned long long)ds[i].N + spg - 1) / spg) * 4ull;      // 16-byte groups, in 32-bit words
            subs += ntr * (((unsigned long long)ds[i].N + SPLIT_SUB - 1ull) / SPLIT_SUB);
        }
        meta[nsegs] = w;
        SplitParams SP{};
        SP.edge_bytes = eb;
        SP.fast_walk = (n == 2 && ctx->learn_warm % 128u == 0u) ? 1 : 0;
        SP.chunk = chunk;
        for (uint32_t i = 0; i < nsegs; ++i) SP.max_chunks = std::max<unsigned long long>(SP.max_chunks, ((unsigned long long)ds[i].N + chunk - 1ull) / chunk);
        for (uint32_t i = 0; i < nsegs; ++i) SP.max_trials = std::max<unsigned long long>(SP.max_trials, ds[i].trial_end - ds[i].trial_begin);
        if (nsegs > 65535) return fail(ctx, MVD_E_INVALID, "too many segments for one call");
        SP.warm = ctx->learn_warm;
        SP.nchains = (uint32_t)trials;
        SP.nwork = w;
        SP.sequential = ctx->split_sequential ? 1 : 0;
        SP.cls = ctx->no_split_classes ? SplitClasses{} : ctx->split_cls;
        CK(ctx->d_smeta.reserve(meta.size() * 8 + 16));
        CK(h2d(ctx, ctx->d_smeta.p, meta.data(), meta.size() * 8));
        SP.work_begin = ctx->d_smeta.as<unsigned long long>();
        SP.edge_begin = SP.work_begin + nsegs + 1;
        SP.sub_begin = SP.work_begin + 2 * (size_t)nsegs + 1;
        CK(ctx->d_lspec.reserve((size_t)w * 4));
        CK(ctx->d_lend.reserve((size_t)w * 4));
        CK(ctx->d_ldirty.reserve(4));
        CK(cudaMemsetAsync(ctx->d_ldirty.p, 0, 4, ctx->stream));
        CK(ctx->d_sedges.reserve((size_t)ew * 4));
        CK(ctx->d_sapx.reserve((size_t)subs * 8));
        CK(ctx->d_splan.reserve((size_t)subs * 4));
        CK(ctx->d_sres.reserve((size_t)subs * 16));
        // per (table, edge): the binade in which each term is a round-half-even tie, the terms in float32, "some term > 0"
        CK(ctx->d_sflags.reserve(16));                          // [0] flags (u32), [8] sub-chunks added term by term (u64)
        if (!ctx->split_tables_ready) {
            const size_t cells = (size_t)SR * ctx->ntables;
            CK(ctx->d_stie.reserve(cells * 4));
            CK(ctx->d_sapxtab.reserve(cells * 8));
            CK(ctx->d_stiek.reserve(16 * (size_t)ctx->ntables));
            CK(mvd_launch_split_tables(ctx->d_ll.as<double2>(), (uint32_t)SR, ctx->ntables, ctx->d_stie.as<uint32_t>(), ctx->d_sapxtab.as<float2>(),
                                       ctx->d_sflags.as<uint32_t>(), ctx->d_stiek.as<unsigned long long>(),
                                       ctx->no_split_classes ? SplitClasses{} : ctx->split_cls, ctx->stream));
            ctx->launches += 1;
            ctx->split_tables_ready = true;
        }
        CK(cudaMemsetAsync(ctx->d_sflags.as<unsigned char>() + 8, 0, 8, ctx->stream));
        SP.spec_start = ctx->d_lspec.as<uint32_t>();
        SP.end = ctx->d_lend.as<uint32_t>();
        SP.ndirty = ctx->d_ldirty.as<uint32_t>();
        SP.edges = ctx->d_sedges.as<uint32_t>();
        SP.apx = ctx->d_sapx.as<float2>();
        SP.plan = ctx->d_splan.as<uint32_t>();
        SP.res = ctx->d_sres.as<double2>();
        SP.tietab = ctx->d_stie.as<uint32_t>();
        SP.apxtab = ctx->d_sapxtab.as<float2>();
        SP.flags = ctx->d_sflags.as<uint32_t>();
        SP.tiek = ctx->d_stiek.as<unsigned long long>();
        SP.nseq = reinterpret_cast<unsigned long long*>(ctx->d_sflags.as<unsigned char>() + 8);
        // shared memory: log-likelihood rows 8 x replicated (16-byte pitch), tie / float32 rows 16 x (8-byte pitch), when they fit
        SP.ll_rep_shift = (size_t)SR * 128 <= 64 * 1024 ? 3 : 0;
        SP.apx_rep_shift = (size_t)SR * 128 <= 64 * 1024 ? 4 : 0;
        const size_t nb = ((size_t)SR * 4 + 15) & ~(size_t)15, lb = ((size_t)SR * 16) << SP.ll_rep_shift;
        const size_t ab = ((size_t)SR * 8) << SP.apx_rep_shift, tb = ((size_t)SR * 4) << (SP.ll_rep_shift ? SP.ll_rep_shift + 2 : 0);
        SP.nxt_in_smem = nb + ab <= 96 * 1024;
        SP.ll_in_smem = lb + tb <= 128 * 1024;
        SP.walk_apx_offset = (uint32_t)nb;
        SP.isum_tie_offset = (uint32_t)lb;
        SP.chain_block = SPLIT_BLOCK;                            // scoring kernel: spread few chains over all SMs
        while (SP.chain_block > 32 && (trials + SP.chain_block - 1) / SP.chain_block < 2 * sms) SP.chain_block >>= 1;
        SP.score_ring_offset = SP.ll_in_smem ? (uint32_t)((lb + 127) & ~(size_t)127) : 0u;
        const size_t ring_bytes = 2 * (size_t)SPLIT_RB_HOST * SP.chain_block * 20;   // two batches of {partial sums 16 B, plan 4 B} per thread
        CK(cudaStreamSynchronize(ctx->stream));                 // meta is a stack-lifetime vector
        le = mvd_launch_split(SP.nxt_in_smem ? (SP.fast_walk ? nb + ab : nb) : 0, SP.ll_in_smem ? lb + tb : 0,
                              SP.score_ring_offset + ring_bytes, ctx->stream, P, SP);
        ctx->launches += SP.sequential ? 2 : 3;
        ctx->last_fast = 16384;
        ctx->last_split_sub = subs;
    } else if (fast) {
        // segments travel as kernel parameters, DET2_MAXSEG per launch; grid.y = segment
        le = cudaSuccess;
        uint32_t extra_launches = 0;
        for (uint32_t base = 0; base < nsegs && le == cudaSuccess; base += DET2_MAXSEG) {
            const uint32_t cnt = std::min<uint32_t>(DET2_MAXSEG, nsegs - base);
            SegBatch B{};
            uint64_t maxblocks = 0;
            for (uint32_t i = 0; i < cnt; ++i) {
                B.s[i] = ds[base + i];
                B.s[i].block_begin = base + i;                       // global segment index (tally slot)
                maxblocks = std::max<uint64_t>(maxblocks, (ds[base + i].trial_end - ds[base + i].trial_begin + block - 1) / block);
            }
            if (maxblocks == 0) continue;
            const dim3 g2((unsigned)maxblocks, cnt);
            // m = 4: the bucket displacements ride in shared memory when they fit beside two resident blocks
            const bool pair4_ds = (size_t)ctx->ph_nb * 4 <= 64 * 1024;       // bucket displacements in shared memory
            const size_t pair4_smem = ((((pair_sbase + DET2P_QUEUES + 255) & ~(size_t)255) + 256 + 256 + 1023) & ~(size_t)1023) + (pair4_ds ? (size_t)ctx->ph_nb * 4 : 0) - pair_sbase;
            if (pair3) le = mvd_launch_det3_pair(m, g2, threads, pair4 ? pair4_smem : pair3_smem, ctx->stream, P, B, pair4 && pair4_ds, P.bm_antipodal != 0);
            else le = mvd_launch_det2(det2_lk, m, det2_lls, det2_gt, pair, g2, threads, pair ? pair_smem : det2_smem, ctx->stream, P, B);
            extra_launches += 1;
        }
        ctx->last_fast = 1 + det2_lk + 16 * (pair3s ? 5 : det2_lls) + ((pair || pair3) ? 256 : 0) + (det2_gt ? 512 : 0);
        if (extra_launches > 1) ctx->launches += extra_launches - 1;  // the common increment below counts one
    } else {
        le = mvd_launch_generic(engine, mode, n2, m, in_smem, grid, smem, ctx->stream, P);
    }
    if (!fast && !split && !(plearn && maxL > 0)) ctx->last_fast = 0;
    if (le != cudaSuccess) return fail(ctx, MVD_E_CUDA, "kernel launch failed: %s", cudaGetErrorString(le));
    ctx->launches += 1;
    CK(cudaEventRecord(ev1, ctx->stream));
    if (mode == MODE_DETECT && out.d_tallies)
        CK(cudaMemcpyAsync(out.d_tallies, ctx->d_tallies.p, 8 * (size_t)nsegs, cudaMemcpyDeviceToDevice, ctx->stream));
    if (go_async) {
        ctx->async_pending += 1;
        return MVD_OK;
    }
    if (ctx->async_pending) {                      // (a split launch among asynchronous ones)
        const int rc = drain_async(ctx);
        if (rc != MVD_OK) return rc;
    }

    // ---- read back
    int herr = 0;
    std::vector<uint32_t> hdirty;
    CK(d2h(ctx, &herr, ctx->d_err.p, sizeof(int)));
    if (mode == MODE_DETECT) {
        if (split) {
            hdirty.assign(1, 0u);
            CK(d2h(ctx, hdirty.data(), ctx->d_ldirty.p, 4));
            CK(d2h(ctx, &ctx->last_split_seq, ctx->d_sflags.as<unsigned char>() + 8, 8));
        }
        if (out.tallies) CK(d2h(ctx, out.tallies, ctx->d_tallies.p, 8 * (size_t)nsegs));
        if (out.logp) CK(d2h(ctx, out.logp, ctx->d_logp.p, 16 * (size_t)trials));
    } else if (mode == MODE_LEARN) {
        if (out.counts) CK(d2h(ctx, out.counts, ctx->d_counts.p, 8 * (size_t)nsegs * SR));
        if (plearn && maxL > 0) {
            hdirty.assign(nsegs, 0u);
            CK(d2h(ctx, hdirty.data(), ctx->d_ldirty.p, 4 * (size_t)nsegs));
        }
    } else if (mode == MODE_TRACE) {
        const size_t cells = (size_t)trials * ((size_t)segs[0].N + 1);
        CK(d2h(ctx, out.trace_idx, ctx->d_trace_idx.p, 4 * cells));
        if (out.trace_met && engine == MVD_ENGINE_ACS)
            CK(d2h(ctx, out.trace_met, ctx->d_trace_met.p, cells * nstate));
    } else {
        CK(d2h(ctx, out.hashes, ctx->d_hashes.p, 8 * (size_t)trials));
        if (out.final_met) CK(d2h(ctx, out.final_met, ctx->d_final.p, (size_t)trials * nstate));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    ctx->last_dirty = 0;
    for (uint32_t v : hdirty) ctx->last_dirty += v;
    if (mode == MODE_TRACE && out.trace_met && engine == MVD_ENGINE_FSM) {
        // FSM engine: the metric vectors are a gather from the state table
        const size_t cells = (size_t)trials * ((size_t)segs[0].N + 1);
        for (size_t c = 0; c < cells; ++c)
            memcpy(out.trace_met + c * nstate, ctx->h_metrics.data() + (size_t)out.trace_idx[c] * nstate, nstate);
    }
    if (herr & 1) return fail(ctx, MVD_E_UNKNOWN_STATE, "a relative-metric vector was not in the state table (KeyError)");
    if (herr & 2) return fail(ctx, MVD_E_UNKNOWN_STATE, "relative metric exceeded 15 (hash key overflow)");
    if (herr & 4) return fail(ctx, MVD_E_CUDA, "pair kernel: the dynamic shared window does not start where the host's layout assumed");
    return MVD_OK;
}

}  // namespace

// =========================================================================================== C ABI
extern "C" {

int mvd_abi_version(void) { return MVD_ABI_VERSION; }

int mvd_create(mvd_ctx** out, int device) {
    mvd_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, MVD_E_INVALID, "null out pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, MVD_E_CUDA, "no CUDA device available (%s); libmvd has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, MVD_E_INVALID, "device %d out of range (count %d)", device, count);
    ctx = new mvd_ctx();
    ctx->device = device;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&ctx->prop, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
    if (e != cudaSuccess) {
        fail(nullptr, MVD_E_CUDA, "context creation failed: %s", cudaGetErrorString(e));
        delete ctx;
        return MVD_E_CUDA;
    }
    ctx->own_stream = true;
    *out = ctx;
    return MVD_OK;
}

int mvd_destroy(mvd_ctx* ctx) {
    if (!ctx) return MVD_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->d_bm, &ctx->d_nxt, &ctx->d_ll, &ctx->d_hkeys, &ctx->d_hvals, &ctx->d_segs, &ctx->d_tallies,
                      &ctx->d_counts, &ctx->d_logp, &ctx->d_trace_idx, &ctx->d_trace_met, &ctx->d_hashes, &ctx->d_final,
                      &ctx->d_err, &ctx->d_bits, &ctx->d_peak, &ctx->d_dstate, &ctx->d_dstate2, &ctx->d_stage, &ctx->d_lspec, &ctx->d_lend, &ctx->d_ldirty, &ctx->d_tcode, &ctx->d_gfsm1,
                      &ctx->d_smeta, &ctx->d_sedges, &ctx->d_phd, &ctx->d_pht, &ctx->d_llslot, &ctx->d_sapx, &ctx->d_splan,
                      &ctx->d_sres, &ctx->d_stie, &ctx->d_sapxtab, &ctx->d_sflags, &ctx->d_stiek, &ctx->d_enc};
    for (DevBuf* b : bufs) b->release();
    for (int i = 0; i < 2; ++i) {
        if (ctx->pin[i]) cudaFreeHost(ctx->pin[i]);
        if (ctx->pin_ev[i]) cudaEventDestroy(ctx->pin_ev[i]);
    }
    for (auto& pr : ctx->async_ev) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return MVD_OK;
}

const char* mvd_last_error(const mvd_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int mvd_set_stream(mvd_ctx* ctx, void* cuda_stream) {
    if (!ctx) return MVD_E_INVALID;
    cudaSetDevice(ctx->device);
    const int rc = drain_async(ctx);
    if (rc != MVD_OK) return rc;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
    return MVD_OK;
}

int mvd_synchronize(mvd_ctx* ctx) {
    if (!ctx) return MVD_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    const int rc = drain_async(ctx);
    if (rc != MVD_OK) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return MVD_OK;
}

int mvd_async_stats(mvd_ctx* ctx, double* kernel_ms_sum, uint64_t* launches) {
    if (!ctx) return MVD_E_INVALID;
    if (kernel_ms_sum) *kernel_ms_sum = ctx->async_ms_sum;
    if (launches) *launches = ctx->async_launches;
    ctx->async_ms_sum = 0.0;
    ctx->async_launches = 0;
    return MVD_OK;
}

int mvd_set_code(mvd_ctx* ctx, int k, int n, int m, const uint32_t* dec_taps) {
    if (!ctx || !dec_taps) return ctx ? fail(ctx, MVD_E_INVALID, "null taps") : MVD_E_INVALID;
    if (k != 1) return fail(ctx, MVD_E_UNSUPPORTED, "device path supports k = 1 codes only (got k=%d)", k);
    if (n < 1 || n > MVD_MAX_N) return fail(ctx, MVD_E_UNSUPPORTED, "n=%d outside [1,%d]", n, MVD_MAX_N);
    if (m < 1 || m > MVD_MAX_M) return fail(ctx, MVD_E_UNSUPPORTED, "m=%d outside [1,%d]", m, MVD_MAX_M);
    for (int j = 0; j < n; ++j)
        if (dec_taps[j] >> (m + 1)) return fail(ctx, MVD_E_INVALID, "decoder tap mask %d has bits beyond memory m=%d", j, m);
    CK(cudaSetDevice(ctx->device));
    ctx->k = k;
    ctx->n = n;
    ctx->m = m;
    ctx->table_code = false;
    ctx->nenc = 0;
    for (int j = 0; j < MVD_MAX_N; ++j) ctx->dec_taps[j] = j < n ? dec_taps[j] : 0u;
    // branch metrics, 16x2 packed: low half = input 0 (ns = 2g), high half = input 1 (ns = 2g+1)
    const int nstate = 1 << m, NP = nstate / 2, HALF = nstate / 2, R = 1 << n;
    std::vector<uint32_t> bm((size_t)R * 2 * NP);
    for (int r = 0; r < R; ++r)
        for (int g = 0; g < NP; ++g) {
            auto d = [&](uint32_t ps, uint32_t u) { return (uint32_t)__builtin_popcount((unsigned)(label_of_branch(ps, u, ctx->dec_taps, n) ^ r)); };
            bm[((size_t)r * NP + g) * 2 + 0] = d(g, 0) | (d(g, 1) << 16);
            bm[((size_t)r * NP + g) * 2 + 1] = d(g + HALF, 0) | (d(g + HALF, 1) << 16);
        }
    CK(ctx->d_bm.reserve(bm.size() * 4));
    CK(h2d(ctx, ctx->d_bm.p, bm.data(), bm.size() * 4));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_code = true;
    ctx->have_states = false;
    ctx->ntables = 0;
    // warm-up of the chunk-parallel chains: survivor paths of longer memories merge later.  Measured at m = 4
    // (S = 150 743, p = 1/2): 2 841 of 235 536 chunks mis-speculated with 128 steps (40 ms in the serial fix pass),
    // 21 with 256, none with 384 (0.7 ms).
    if (!ctx->learn_warm_set) ctx->learn_warm = m <= 3 ? LEARN_WARM : LEARN_WARM * (uint32_t)(m - 1);
    return MVD_OK;
}

int mvd_set_code_tables(mvd_ctx* ctx, int k, int n, int m, const uint8_t* dec_prev, const uint8_t* dec_label) {
    if (!ctx) return MVD_E_INVALID;
    if (!dec_prev || !dec_label) return fail(ctx, MVD_E_INVALID, "null trellis tables");
    if (k < 1 || k > MVD_MAX_K) return fail(ctx, MVD_E_UNSUPPORTED, "k=%d outside [1,%d]", k, MVD_MAX_K);
    if (n < 1 || n > MVD_MAX_N) return fail(ctx, MVD_E_UNSUPPORTED, "n=%d outside [1,%d]", n, MVD_MAX_N);
    if (m < 1 || m > MVD_MAX_M) return fail(ctx, MVD_E_UNSUPPORTED, "m=%d outside [1,%d]", m, MVD_MAX_M);
    const size_t cells = (size_t)1 << (m + k);
    for (size_t i = 0; i < cells; ++i) {
        if (dec_prev[i] >> m) return fail(ctx, MVD_E_INVALID, "dec_prev[%zu]=%u is not a trellis state (m=%d)", i, dec_prev[i], m);
        if (dec_label[i] >> n) return fail(ctx, MVD_E_INVALID, "dec_label[%zu]=%u has more than n=%d bits", i, dec_label[i], n);
    }
    CK(cudaSetDevice(ctx->device));
    ctx->k = k;
    ctx->n = n;
    ctx->m = m;
    ctx->table_code = true;
    ctx->nenc = 0;
    ctx->h_dec_prev.assign(dec_prev, dec_prev + cells);
    ctx->h_dec_lab.assign(dec_label, dec_label + cells);
    for (int j = 0; j < MVD_MAX_N; ++j) ctx->dec_taps[j] = 0u;
    ctx->have_code = true;
    ctx->have_states = false;
    ctx->ntables = 0;
    return MVD_OK;
}

int mvd_set_encoders(mvd_ctx* ctx, uint32_t nenc, const uint8_t* enc_next, const uint8_t* enc_out) {
    if (!ctx) return MVD_E_INVALID;
    if (!ctx->have_code || !ctx->table_code) return fail(ctx, MVD_E_STATE, "mvd_set_code_tables first");
    if (!enc_next || !enc_out || nenc == 0 || nenc > 65536u) return fail(ctx, MVD_E_INVALID, "null / empty encoder tables");
    const size_t cells = (size_t)1 << (ctx->m + ctx->k);
    std::vector<uint16_t> packed((size_t)nenc * cells);
    for (size_t i = 0; i < packed.size(); ++i) {
        if (enc_next[i] >> ctx->m) return fail(ctx, MVD_E_INVALID, "enc_next[%zu]=%u is not an encoder state (m=%d)", i, enc_next[i], ctx->m);
        if (enc_out[i] >> ctx->n) return fail(ctx, MVD_E_INVALID, "enc_out[%zu]=%u has more than n=%d bits", i, enc_out[i], ctx->n);
        packed[i] = (uint16_t)((uint16_t)enc_next[i] << 8 | enc_out[i]);
    }
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_enc.reserve(packed.size() * 2));
    CK(h2d(ctx, ctx->d_enc.p, packed.data(), packed.size() * 2));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->nenc = nenc;
    return MVD_OK;
}

int mvd_set_states(mvd_ctx* ctx, uint32_t S, const uint8_t* metrics, const uint32_t* next) {
    if (!ctx) return MVD_E_INVALID;
    if (!ctx->have_code) return fail(ctx, MVD_E_STATE, "mvd_set_code first");
    if (!metrics || !next || S == 0) return fail(ctx, MVD_E_INVALID, "null / empty state table");
    const int nstate = 1 << ctx->m, R = 1 << ctx->n;
    ctx->S = S;
    ctx->h_metrics.assign(metrics, metrics + (size_t)S * nstate);
    ctx->h_next.assign(next, next + (size_t)S * R);
    return install_states(ctx);
}

int mvd_enumerate_states(mvd_ctx* ctx, uint32_t max_states, uint32_t* S_out) {
    if (!ctx) return MVD_E_INVALID;
    if (!ctx->have_code) return fail(ctx, MVD_E_STATE, "mvd_set_code first");
    if (max_states == 0) return fail(ctx, MVD_E_INVALID, "max_states = 0");
    if (ctx->table_code) return fail(ctx, MVD_E_UNSUPPORTED, "mvd_enumerate_states needs the tap-mask form of the code (mvd_set_code)");
    const int n = ctx->n, m = ctx->m, nstate = 1 << m, R = 1 << n, HALF = nstate / 2;
    // branch labels of the butterfly
    std::vector<int> lab0(nstate), lab1(nstate);   // label of branch into ns from ps0 / ps1
    for (int ns = 0; ns < nstate; ++ns) {
        lab0[ns] = label_of_branch((uint32_t)(ns >> 1), (uint32_t)(ns & 1), ctx->dec_taps, n);
        lab1[ns] = label_of_branch((uint32_t)((ns >> 1) + HALF), (uint32_t)(ns & 1), ctx->dec_taps, n);
    }
    // discovery-ordered closure: the vector of states doubles as the BFS queue
    std::vector<uint8_t>& met = ctx->h_metrics;
    std::vector<uint32_t>& nxt = ctx->h_next;
    met.assign(nstate, 0);
    nxt.clear();
    size_t cap = 1024;
    std::vector<uint32_t> slots(cap, MVD_EMPTY);
    auto vhash = [&](const uint8_t* v) {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (int i = 0; i < nstate; ++i) h = (h ^ v[i]) * 0xD6E8FEB86659FD93ull;
        return (size_t)(h ^ (h >> 32));
    };
    auto insert_slot = [&](uint32_t idx) {
        size_t h = vhash(met.data() + (size_t)idx * nstate) & (cap - 1);
        while (slots[h] != MVD_EMPTY) h = (h + 1) & (cap - 1);
        slots[h] = idx;
    };
    insert_slot(0);
    uint32_t S = 1;
    std::vector<uint8_t> cand(nstate);
    for (uint32_t cur = 0; cur < S; ++cur) {
        for (int r = 0; r < R; ++r) {
            int lo = 1 << 30;
            int tmp[64];
            for (int ns = 0; ns < nstate; ++ns) {
                const int a = met[(size_t)cur * nstate + (ns >> 1)] + __builtin_popcount((unsigned)(lab0[ns] ^ r));
                const int b = met[(size_t)cur * nstate + (ns >> 1) + HALF] + __builtin_popcount((unsigned)(lab1[ns] ^ r));
                tmp[ns] = a < b ? a : b;
                lo = tmp[ns] < lo ? tmp[ns] : lo;
            }
            for (int ns = 0; ns < nstate; ++ns) {
                const int v = tmp[ns] - lo;
                if (v > 255) return fail(ctx, MVD_E_UNSUPPORTED, "relative metric exceeds 8 bits");
                cand[ns] = (uint8_t)v;
            }
            size_t h = vhash(cand.data()) & (cap - 1);
            uint32_t found = MVD_EMPTY;
            while (slots[h] != MVD_EMPTY) {
                if (memcmp(met.data() + (size_t)slots[h] * nstate, cand.data(), nstate) == 0) { found = slots[h]; break; }
                h = (h + 1) & (cap - 1);
            }
            if (found == MVD_EMPTY) {
                if (S >= max_states) {
                    ctx->have_states = false;
                    return fail(ctx, MVD_E_NOMEM, "more than %u Markov states", max_states);
                }
                met.insert(met.end(), cand.begin(), cand.end());
                found = S++;
                slots[h] = found;
                if ((size_t)S * 2 > cap) {          // grow + rehash
                    cap <<= 1;
                    slots.assign(cap, MVD_EMPTY);
                    for (uint32_t i = 0; i < S; ++i) insert_slot(i);
                }
            }
            nxt.push_back(found);
        }
    }
    ctx->S = S;
    if (S_out) *S_out = S;
    return install_states(ctx);
}

int mvd_enumerate_states_gpu(mvd_ctx* ctx, uint32_t max_states, uint32_t flags, uint32_t chunk_parents, mvd_bfs_stats* stats) {
    if (!ctx) return MVD_E_INVALID;
    if (!ctx->have_code) return fail(ctx, MVD_E_STATE, "mvd_set_code first");
    if ((flags & MVD_BFS_INSTALL) && (flags & MVD_BFS_COUNT_ONLY)) return fail(ctx, MVD_E_INVALID, "INSTALL and COUNT_ONLY exclude each other");
    if (ctx->table_code) return fail(ctx, MVD_E_UNSUPPORTED, "mvd_enumerate_states_gpu needs the tap-mask form of the code (mvd_set_code)");
    CK(cudaSetDevice(ctx->device));
    MvdBfsConfig cfg;
    cfg.n = ctx->n;
    cfg.m = ctx->m;
    for (int j = 0; j < MVD_MAX_N; ++j) cfg.dec_taps[j] = ctx->dec_taps[j];
    cfg.max_states = max_states;
    cfg.chunk_parents = chunk_parents;
    cfg.keep_next = !(flags & MVD_BFS_COUNT_ONLY);
    cfg.copy_out = (flags & MVD_BFS_INSTALL) != 0;
    MvdBfsResult res;
    const int rc = mvd_bfs_run(cfg, ctx->stream, res);
    ctx->launches += res.launches;
    ctx->last_ms = res.ms;
    ctx->last_fast = 2048;
    ctx->bfs_levels = res.levels;
    if (stats) {
        stats->S = res.S;
        stats->frontier = res.frontier;
        stats->iterations = res.iterations;
        stats->launches = res.launches;
        stats->candidates = res.candidates;
        stats->closed = res.closed ? 1 : 0;
        stats->max_metric = res.max_metric;
        stats->ms = res.ms;
        stats->reserved = 0;
    }
    if (rc != MVD_OK) return fail(ctx, rc, "%s", res.error);
    if (flags & MVD_BFS_INSTALL) {
        ctx->S = res.S;
        ctx->h_metrics.swap(res.metrics);
        ctx->h_next.swap(res.next);
        return install_states(ctx);
    }
    return MVD_OK;
}

int mvd_bfs_levels(mvd_ctx* ctx, uint32_t* sizes, uint32_t cap, uint32_t* nlevels) {
    if (!ctx) return MVD_E_INVALID;
    if (nlevels) *nlevels = (uint32_t)ctx->bfs_levels.size();
    if (sizes)
        for (uint32_t i = 0; i < cap && i < ctx->bfs_levels.size(); ++i) sizes[i] = ctx->bfs_levels[i];
    return MVD_OK;
}

int mvd_get_states(mvd_ctx* ctx, uint8_t* metrics, uint32_t* next) {
    if (!ctx) return MVD_E_INVALID;
    if (!ctx->have_states) return fail(ctx, MVD_E_STATE, "no state table");
    if (metrics) memcpy(metrics, ctx->h_metrics.data(), ctx->h_metrics.size());
    if (next) memcpy(next, ctx->h_next.data(), ctx->h_next.size() * 4);
    return MVD_OK;
}

int mvd_set_loglik(mvd_ctx* ctx, uint32_t ntables, const double* logP1, const double* logTref) {
    if (!ctx) return MVD_E_INVALID;
    if (!ctx->have_states) return fail(ctx, MVD_E_STATE, "mvd_set_states first");
    if (!logP1 || !logTref || ntables == 0) return fail(ctx, MVD_E_INVALID, "null / empty log-likelihood tables");
    CK(cudaSetDevice(ctx->device));
    const size_t SR = (size_t)ctx->S << ctx->n;
    // the tables go up as they are (log P1 [ntables][SR], log Tref [SR]); the interleaved rows the kernels read and the
    // packed NEXT-walk entries are written on the device (pack_ll_kernel) -- at S = 150 743 the host loops that used to
    // do this took 84 ms per call, 94 % of run_experiment's wall time together with the P1 / log tables
    CK(ctx->d_stage.reserve((SR * ntables + SR) * 8));
    double* d_lp1 = ctx->d_stage.as<double>();
    double* d_ltref = d_lp1 + SR * ntables;
    CK(h2d(ctx, d_lp1, logP1, SR * ntables * 8));
    CK(h2d(ctx, d_ltref, logTref, SR * 8));
    CK(ctx->d_ll.reserve(2 * SR * ntables * 8));
    // log Tref[e] == c[e] * unit exactly, c = 0 or a power of two?  (unit = the non-zero value of least magnitude)
    bool large = false;
    {
        // log Tref has a handful of distinct values (T(1/2) = mult / 2^n: at most 2^n of them), so the three scans below -- least
        // magnitude, value classes, the packed multiple c of every entry -- go through a memo of the distinct values in order of first
        // appearance instead of a division and a frexp per entry (S = 150 743: 603 000 entries per call); more than 16 distinct values
        // (or NaNs, which never compare equal) take the entry-by-entry form
        constexpr int MEMO = 16;
        double mv[MEMO];
        int nm = 0;
        bool many = false;
        for (size_t e = 0; e < SR && !many; ++e) {
            const double v = logTref[e];
            int j = 0;
            while (j < nm && mv[j] != v) ++j;
            if (j == nm) {
                if (nm == MEMO) many = true;
                else mv[nm++] = v;
            }
        }
        auto code_of = [](double v, double unit, bool& ok) -> uint32_t {
            double c = 0.0;
            if (v != 0.0) {
                c = v / unit;
                int ex = 0;
                ok = std::frexp(c, &ex) == 0.5 && c >= 1.0 && c <= 1048576.0 && c * unit == v;
            }
            uint64_t bits;
            memcpy(&bits, &c, 8);
            ok = ok && (uint32_t)bits == 0u;
            return (uint32_t)(bits >> 32);
        };
        double unit = 0.0;
        SplitClasses cls{};
        bool ok = true;
        std::vector<uint32_t> code(SR, 0u);
        if (!many) {
            for (int j = 0; j < nm; ++j)
                if (mv[j] != 0.0 && (unit == 0.0 || std::fabs(mv[j]) < std::fabs(unit))) unit = mv[j];
            // the distinct non-zero values of log Tref, if there are at most three (class mode of the split path, mvd_split.cuh)
            for (int j = 0; j < nm && cls.n >= 0; ++j) {
                const double v = mv[j];
                if (v == 0.0) continue;
                if (cls.n == 3 || !(v < 0.0)) cls.n = -1;               // a fourth value, a positive one: no class mode
                else cls.val[cls.n++] = v;
            }
            uint32_t mc[MEMO];
            for (int j = 0; j < nm; ++j) {
                bool okj = true;
                mc[j] = code_of(mv[j], unit, okj);
                ok = ok && okj;
            }
            if (ok)
                for (size_t e = 0; e < SR; ++e) {
                    const double v = logTref[e];
                    int j = 0;
                    while (mv[j] != v) ++j;                             // every value is in the memo
                    code[e] = mc[j];
                }
        } else {
            for (size_t e = 0; e < SR; ++e)
                if (logTref[e] != 0.0 && (unit == 0.0 || std::fabs(logTref[e]) < std::fabs(unit))) unit = logTref[e];
            for (size_t e = 0; e < SR && cls.n >= 0; ++e) {
                const double v = logTref[e];
                if (v == 0.0) continue;
                int j = 0;
                while (j < cls.n && cls.val[j] != v) ++j;
                if (j == cls.n) {
                    if (cls.n == 3 || !(v < 0.0)) cls.n = -1;           // a fourth value, a positive one or a NaN: no class mode
                    else cls.val[cls.n++] = v;
                }
            }
            for (size_t e = 0; e < SR && ok; ++e) {
                bool oke = true;
                code[e] = code_of(logTref[e], unit, oke);
                ok = oke;
            }
        }
        if (cls.n < 0) cls = SplitClasses{};
        ctx->split_cls = cls;
        ctx->tref_packed = ok;
        ctx->tref_unit = unit;
        ctx->have_gfsm1 = false;
        if (ok) {
            CK(ctx->d_tcode.reserve(SR * 4));
            CK(h2d(ctx, ctx->d_tcode.p, code.data(), SR * 4));
            // large S: packed NEXT-walk entries {log P1, next row byte offset, c} stay in global memory
            large = 128 + (SR << 7) + 64 > ctx->prop.sharedMemPerBlockOptin && SR < (1ull << 28);
            if (large) CK(ctx->d_gfsm1.reserve(16 * SR * ntables));
        }
    }
    CK(mvd_launch_pack_ll(d_lp1, d_ltref, ctx->d_nxt.as<uint32_t>(), ctx->d_tcode.as<uint32_t>(), (uint32_t)SR, ntables,
                          ctx->d_ll.as<double2>(), large ? ctx->d_gfsm1.as<uint4>() : nullptr, ctx->stream));
    ctx->launches += 1;
    ctx->have_gfsm1 = large;
    // m = 4 two-trials-per-thread ACS kernel: the same rows in hash-slot order (64 B per slot and table)
    ctx->have_llslot = false;
    if (ctx->ph_slots && ctx->m == 4 && (size_t)ctx->ph_slots * 64 * ntables <= ((size_t)24 << 30)) {
        CK(ctx->d_llslot.reserve((size_t)ctx->ph_slots * 64 * ntables));
        CK(mvd_launch_slot_rows(ctx->d_ll.as<double2>(), ctx->d_pht.as<uint32_t>(), ctx->ph_slots, (uint32_t)SR, ntables,
                                ctx->d_llslot.as<double2>(), ctx->stream));
        ctx->launches += 1;
        ctx->have_llslot = true;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->ntables = ntables;
    ctx->split_tables_ready = false;
    return MVD_OK;
}

int mvd_learn_counts(mvd_ctx* ctx, const mvd_src* src, const mvd_segment* segs, uint32_t nsegs, uint32_t burn, int engine,
                     uint64_t* edge_counts) {
    if (!ctx) return MVD_E_INVALID;
    if (!edge_counts) return fail(ctx, MVD_E_INVALID, "null edge_counts");
    LaunchOut o;
    o.counts = edge_counts;
    o.burn = burn;
    return run(ctx, MODE_LEARN, engine, src, segs, nsegs, o);
}

int mvd_detect(mvd_ctx* ctx, const mvd_src* src, const mvd_segment* segs, uint32_t nsegs, int engine, uint64_t* tallies,
               double* logp, void* d_tallies) {
    if (!ctx) return MVD_E_INVALID;
    if (!tallies && !d_tallies) return fail(ctx, MVD_E_INVALID, "no tally destination");
    LaunchOut o;
    o.tallies = tallies;
    o.logp = logp;
    o.d_tallies = d_tallies;
    return run(ctx, MODE_DETECT, engine, src, segs, nsegs, o);
}

int mvd_trace(mvd_ctx* ctx, const mvd_src* src, const mvd_segment* seg, int engine, uint32_t* state_idx, uint8_t* metrics) {
    if (!ctx) return MVD_E_INVALID;
    if (!state_idx) return fail(ctx, MVD_E_INVALID, "null state_idx");
    LaunchOut o;
    o.trace_idx = state_idx;
    o.trace_met = metrics;
    return run(ctx, MODE_TRACE, engine, src, seg, 1, o);
}

int mvd_acs_hash(mvd_ctx* ctx, const mvd_src* src, const mvd_segment* seg, uint64_t* hashes, uint8_t* final_metrics) {
    if (!ctx) return MVD_E_INVALID;
    if (!hashes) return fail(ctx, MVD_E_INVALID, "null hashes");
    LaunchOut o;
    o.hashes = hashes;
    o.final_met = final_metrics;
    return run(ctx, MODE_HASH, MVD_ENGINE_ACS, src, seg, 1, o);
}

int mvd_acs_final(mvd_ctx* ctx, const mvd_src* src, const mvd_segment* seg, uint8_t* final_metrics) {
    if (!ctx) return MVD_E_INVALID;
    if (!src || !seg || !final_metrics) return fail(ctx, MVD_E_INVALID, "null argument");
    if (!ctx->have_code) return fail(ctx, MVD_E_STATE, "mvd_set_code has not been called");
    if (ctx->table_code) return fail(ctx, MVD_E_UNSUPPORTED, "mvd_acs_final needs the tap-mask form of the code (mvd_set_code)");
    if (ctx->n != 2 || ctx->m < 2) return fail(ctx, MVD_E_UNSUPPORTED, "mvd_acs_final supports n = 2, m = 2..6 (use mvd_acs_hash)");
    if (src->mode != MVD_SRC_PHILOX) return fail(ctx, MVD_E_UNSUPPORTED, "mvd_acs_final takes the on-device bit source (use mvd_acs_hash)");
    if (seg->N >= 0x80000000u) return fail(ctx, MVD_E_INVALID, "N = %u >= 2^31 overflows the position-addressed Philox counter", seg->N);
    if (seg->trial_end < seg->trial_begin) return fail(ctx, MVD_E_INVALID, "trial_end < trial_begin");
    CK(cudaSetDevice(ctx->device));
    const int m = ctx->m, nstate = 1 << m, HALF = nstate / 2;
    const uint64_t ntr = seg->trial_end - seg->trial_begin;
    if (ntr == 0) return MVD_OK;
    uint32_t sel[128] = {0};
    for (int ns = 0; ns < nstate; ++ns)
        for (int b = 0; b < 2; ++b) {
            const uint32_t L = (uint32_t)label_of_branch((uint32_t)((ns >> 1) + b * HALF), (uint32_t)(ns & 1), ctx->dec_taps, 2);
            sel[2 * ns + b] = L | 0x80u | ((4u + L) << 8) | 0x8000u;
        }
    DevSeg d{};
    d.N = seg->N;
    d.threshold = seg->threshold;
    d.stream = seg->stream;
    for (int j = 0; j < 2; ++j) {
        if (seg->enc_taps[j] >> (m + 1)) return fail(ctx, MVD_E_INVALID, "encoder tap beyond memory m=%d", m);
        d.enc_taps[j] = seg->enc_taps[j];
    }
    d.random_input = seg->random_input ? 1u : 0u;
    d.dmin = seg->threshold ? (uint32_t)__builtin_ctz(seg->threshold) : 32u;
    d.trial_begin = seg->trial_begin;
    d.trial_end = seg->trial_end;
    Params P{};
    P.n = 2;
    P.m = m;
    P.src_mode = MVD_SRC_PHILOX;
    {
        uint32_t k0 = (uint32_t)src->seed, k1 = (uint32_t)(src->seed >> 32);
        for (int r = 0; r < 10; ++r) {
            P.rk0[r] = k0;
            P.rk1[r] = k1;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
    }
    CK(ctx->d_final.reserve((size_t)ntr * nstate));
    const uint64_t blocks = (ntr + 2 * DET2P_BLOCK - 1) / (2 * DET2P_BLOCK);
    if (blocks > 0x7FFFFFFFull) return fail(ctx, MVD_E_INVALID, "too many trials in one call");
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    CK(mvd_launch_acsp(m, dim3((unsigned)blocks), DET2P_BLOCK, ctx->stream, P, d, sel, ctx->d_final.as<uint8_t>()));
    ctx->launches += 1;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(d2h(ctx, final_metrics, ctx->d_final.p, (size_t)ntr * nstate));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    ctx->last_fast = 32768;
    return MVD_OK;
}

int mvd_chernoff_rho(mvd_ctx* ctx, uint32_t K, uint32_t R, const uint32_t* next, const double* lp1, const double* lp2,
                     const double* lb1, const double* lb2, const double* u_vals, uint32_t nu, double tol, uint32_t max_iter,
                     double* rho, uint32_t* iters) {
    if (!ctx) return MVD_E_INVALID;
    if (!next || !lp1 || !lp2 || !lb1 || !lb2 || !u_vals || !rho) return fail(ctx, MVD_E_INVALID, "null argument");
    if (K == 0 || R == 0 || nu == 0 || max_iter == 0) return fail(ctx, MVD_E_INVALID, "empty problem");
    const size_t KR = (size_t)K * R;
    for (size_t e = 0; e < KR; ++e)
        if (next[e] >= K) return fail(ctx, MVD_E_INVALID, "next[%zu]=%u out of range (K=%u)", e, next[e], K);
    CK(cudaSetDevice(ctx->device));
    // weights in scratch while it fits L2-ish sizes; beyond that they are recomputed in every product (mvd_chernoff.cuh)
    const bool recompute = (KR + (size_t)K) * 8 * nu > ((size_t)96 << 20);
    const size_t per_u = recompute ? 2 * (size_t)K * 8 : (KR + 3 * (size_t)K) * 8;
    uint32_t ub = (uint32_t)std::max<size_t>(1, std::min<size_t>(nu, ((size_t)8 << 30) / per_u));
    DevBuf d_in, d_scr, d_out;
    const size_t in_bytes = KR * 4 + 2 * KR * 8 + 2 * (size_t)K * 8 + (size_t)nu * 8;
    CK(d_in.reserve(in_bytes + 64));
    CK(d_scr.reserve((size_t)ub * per_u));
    CK(d_out.reserve((size_t)nu * 12));
    unsigned char* base = d_in.as<unsigned char>();
    double* g_lp1 = reinterpret_cast<double*>(base);
    double* g_lp2 = g_lp1 + KR;
    double* g_lb1 = g_lp2 + KR;
    double* g_lb2 = g_lb1 + K;
    double* g_u = g_lb2 + K;
    uint32_t* g_nxt = reinterpret_cast<uint32_t*>(g_u + nu);
    CK(h2d(ctx, g_lp1, lp1, KR * 8));
    CK(h2d(ctx, g_lp2, lp2, KR * 8));
    CK(h2d(ctx, g_lb1, lb1, (size_t)K * 8));
    CK(h2d(ctx, g_lb2, lb2, (size_t)K * 8));
    CK(h2d(ctx, g_u, u_vals, (size_t)nu * 8));
    CK(h2d(ctx, g_nxt, next, KR * 4));
    ChernoffParams P;
    P.K = K;
    P.R = R;
    P.max_iter = max_iter;
    P.nxt = g_nxt;
    P.lp1 = g_lp1;
    P.lp2 = g_lp2;
    P.lb1 = g_lb1;
    P.lb2 = g_lb2;
    P.tol = tol;
    if (recompute) {
        P.wd = nullptr;
        P.bgR = nullptr;
        P.xa = d_scr.as<double>();
    } else {
        P.wd = d_scr.as<double>();
        P.bgR = P.wd + (size_t)ub * KR;
        P.xa = P.bgR + (size_t)ub * K;
    }
    P.xb = P.xa + (size_t)ub * K;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    for (uint32_t u0 = 0; u0 < nu; u0 += ub) {
        P.nu = std::min(ub, nu - u0);
        P.u_vals = g_u + u0;
        P.rho = d_out.as<double>() + u0;
        P.iters = reinterpret_cast<uint32_t*>(d_out.as<double>() + nu) + u0;
        CK(mvd_launch_chernoff(P, ctx->stream));
        ctx->launches += 1;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(d2h(ctx, rho, d_out.p, (size_t)nu * 8));
    std::vector<uint32_t> it(nu);
    CK(d2h(ctx, it.data(), d_out.as<double>() + nu, (size_t)nu * 4));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    ctx->last_fast = 4096;
    if (iters) memcpy(iters, it.data(), (size_t)nu * 4);
    d_in.release();
    d_scr.release();
    d_out.release();
    return MVD_OK;
}

int mvd_chernoff_rho_dense(mvd_ctx* ctx, uint32_t K, uint32_t R, const double* logP1, const double* logP2, const double* u_vals,
                           uint32_t nu, double tol, uint32_t max_iter, double* rho, uint32_t* iters) {
    if (!ctx) return MVD_E_INVALID;
    if (!logP1 || !logP2 || !u_vals || !rho) return fail(ctx, MVD_E_INVALID, "null argument");
    if (K == 0 || R == 0 || nu == 0 || max_iter == 0) return fail(ctx, MVD_E_INVALID, "empty problem");
    CK(cudaSetDevice(ctx->device));
    const size_t KKR = (size_t)K * K * R;
    const size_t per_u = ((size_t)K * K + 2 * (size_t)K) * 8;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    if (2 * KKR * 8 + per_u + ((size_t)1 << 30) > free_b)
        return fail(ctx, MVD_E_NOMEM, "dense tensors of K=%u need %.1f GB; use the edge form (mvd_chernoff_rho)", K, 2 * KKR * 8 / 1e9);
    uint32_t ub = (uint32_t)std::max<size_t>(1, std::min<size_t>(nu, ((size_t)8 << 30) / per_u));
    DevBuf d_in, d_scr, d_out;
    CK(d_in.reserve(2 * KKR * 8 + (size_t)nu * 8));
    CK(d_scr.reserve((size_t)ub * per_u));
    CK(d_out.reserve((size_t)nu * 12));
    double* g_lp1 = d_in.as<double>();
    double* g_lp2 = g_lp1 + KKR;
    double* g_u = g_lp2 + KKR;
    CK(h2d(ctx, g_lp1, logP1, KKR * 8));
    CK(h2d(ctx, g_lp2, logP2, KKR * 8));
    CK(h2d(ctx, g_u, u_vals, (size_t)nu * 8));
    ChernoffParams P;
    memset(&P, 0, sizeof P);
    P.K = K;
    P.R = R;
    P.max_iter = max_iter;
    P.lp1 = g_lp1;
    P.lp2 = g_lp2;
    P.tol = tol;
    P.wd = d_scr.as<double>();
    P.xa = P.wd + (size_t)ub * K * K;
    P.xb = P.xa + (size_t)ub * K;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    for (uint32_t u0 = 0; u0 < nu; u0 += ub) {
        P.nu = std::min(ub, nu - u0);
        P.u_vals = g_u + u0;
        P.rho = d_out.as<double>() + u0;
        P.iters = reinterpret_cast<uint32_t*>(d_out.as<double>() + nu) + u0;
        CK(mvd_launch_chernoff_dense(P, ctx->stream));
        ctx->launches += 1;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(d2h(ctx, rho, d_out.p, (size_t)nu * 8));
    std::vector<uint32_t> it(nu);
    CK(d2h(ctx, it.data(), d_out.as<double>() + nu, (size_t)nu * 4));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    ctx->last_fast = 4097;
    if (iters) memcpy(iters, it.data(), (size_t)nu * 4);
    d_in.release();
    d_scr.release();
    d_out.release();
    return MVD_OK;
}

int mvd_parity_detect(mvd_ctx* ctx, const mvd_src* src, const mvd_parity_segment* segs, uint32_t nsegs, uint64_t* tallies,
                      uint32_t* satisfied) {
    if (!ctx) return MVD_E_INVALID;
    if (!src || !segs || nsegs == 0 || !tallies) return fail(ctx, MVD_E_INVALID, "null source / segments / tallies");
    if (src->mode != MVD_SRC_PHILOX && src->mode != MVD_SRC_BITSTREAM) return fail(ctx, MVD_E_INVALID, "bad source mode");
    CK(cudaSetDevice(ctx->device));
    std::vector<ParitySeg> ps(nsegs);
    uint64_t trials = 0, need_words = 0, max_blocks = 0;
    for (uint32_t i = 0; i < nsegs; ++i) {
        const mvd_parity_segment& s = segs[i];
        if (s.n < 1 || s.n > MVD_MAX_N) return fail(ctx, MVD_E_UNSUPPORTED, "segment %u: n=%u outside [1,%d]", i, s.n, MVD_MAX_N);
        if (s.m > 31) return fail(ctx, MVD_E_UNSUPPORTED, "segment %u: m=%u > 31", i, s.m);
        if (s.trial_end < s.trial_begin) return fail(ctx, MVD_E_INVALID, "segment %u: trial_end < trial_begin", i);
        if (s.decide > 1) return fail(ctx, MVD_E_INVALID, "segment %u: decide must be 0 or 1", i);
        if ((uint64_t)s.N + s.m > 0xFFFFFF00ull) return fail(ctx, MVD_E_INVALID, "segment %u: N too large", i);
        ParitySeg& d = ps[i];
        memset(&d, 0, sizeof d);
        d.N = s.N;
        d.m = s.m;
        d.n = s.n;
        d.threshold = s.threshold;
        d.stream = s.stream;
        d.dmin = s.threshold ? (uint32_t)__builtin_ctz(s.threshold) : 32u;
        d.decide = s.decide;
        uint32_t any = 0;
        for (uint32_t j = 0; j < s.n; ++j) {
            if (s.m < 31 && (s.enc_taps[j] >> (s.m + 1))) return fail(ctx, MVD_E_INVALID, "segment %u: encoder tap beyond memory m=%u", i, s.m);
            d.enc_taps[j] = s.enc_taps[j];
            d.tmpl[j] = s.tmpl[j];
            any |= s.tmpl[j];
        }
        if (!any) return fail(ctx, MVD_E_INVALID, "segment %u: empty parity template", i);
        d.max_delay = 31u - (uint32_t)__builtin_clz(any);                         // comp_parity.py:101
        d.gamma = s.gamma;
        d.trial_begin = s.trial_begin;
        d.trial_end = s.trial_end;
        d.bits_offset = s.bits_offset;
        d.out_offset = trials;
        d.seg_index = i;
        const uint64_t ntr = s.trial_end - s.trial_begin;
        trials += ntr;
        max_blocks = std::max<uint64_t>(max_blocks, (ntr + PARITY_BLOCK - 1) / PARITY_BLOCK);
        if (src->mode == MVD_SRC_BITSTREAM) {
            const uint64_t nsb = ((uint64_t)s.N + s.m + 127) / 128;
            need_words = std::max<uint64_t>(need_words, s.bits_offset + nsb * (uint64_t)(1 + s.n) * ntr);
        }
    }
    if (max_blocks > 0x7FFFFFFFull) return fail(ctx, MVD_E_INVALID, "too many trials in one segment");
    Params P{};
    P.src_mode = src->mode;
    {
        uint32_t k0 = (uint32_t)src->seed, k1 = (uint32_t)(src->seed >> 32);
        for (int r = 0; r < 10; ++r) {
            P.rk0[r] = k0;
            P.rk1[r] = k1;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
    }
    if (src->mode == MVD_SRC_BITSTREAM) {
        if (!src->bits) return fail(ctx, MVD_E_INVALID, "bitstream source without bits");
        if (src->bits_words < need_words) return fail(ctx, MVD_E_INVALID, "bitstream too short: %llu words, need %llu",
                                                      (unsigned long long)src->bits_words, (unsigned long long)need_words);
        if (src->bits_on_device) {
            P.bits = reinterpret_cast<const uint4*>(src->bits);
        } else {
            CK(ctx->d_bits.reserve((size_t)need_words * 16));
            CK(h2d(ctx, ctx->d_bits.p, src->bits, (size_t)need_words * 16));
            P.bits = ctx->d_bits.as<uint4>();
        }
    }
    CK(ctx->d_tallies.reserve(8 * (size_t)nsegs));
    CK(cudaMemsetAsync(ctx->d_tallies.p, 0, 8 * (size_t)nsegs, ctx->stream));
    P.tallies = ctx->d_tallies.as<unsigned long long>();
    uint32_t* d_sat = nullptr;
    if (satisfied) {
        CK(ctx->d_trace_idx.reserve(4 * (size_t)std::max<uint64_t>(trials, 1)));
        d_sat = ctx->d_trace_idx.as<uint32_t>();
    }
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    for (uint32_t s0 = 0; s0 < nsegs; s0 += PARITY_MAXSEG) {
        const uint32_t cnt = std::min<uint32_t>(PARITY_MAXSEG, nsegs - s0);
        ParityBatch B;
        memset(&B, 0, sizeof B);
        uint64_t mb = 0;
        for (uint32_t i = 0; i < cnt; ++i) {
            B.s[i] = ps[s0 + i];
            mb = std::max<uint64_t>(mb, (ps[s0 + i].trial_end - ps[s0 + i].trial_begin + PARITY_BLOCK - 1) / PARITY_BLOCK);
        }
        if (mb == 0) continue;
        CK(mvd_launch_parity(dim3((unsigned)mb, cnt), ctx->stream, P, B, d_sat));
        ctx->launches += 1;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(d2h(ctx, tallies, ctx->d_tallies.p, 8 * (size_t)nsegs));
    if (satisfied && trials) CK(d2h(ctx, satisfied, d_sat, 4 * (size_t)trials));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    ctx->last_fast = 8192;
    return MVD_OK;
}

int mvd_host_log_table(const double* values, double* out, uint64_t count) {
    if (!values || !out) return MVD_E_INVALID;
    // the same libm call per element whatever the thread count (bit-equal to the single-threaded loop)
    host_parallel(count, 1u << 15, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) out[i] = std::log(values[i] > 1e-300 ? values[i] : 1e-300);
    });
    return MVD_OK;
}

int mvd_host_p1_edge_tables(const uint64_t* edge_counts, const uint32_t* next, uint32_t S, uint32_t R, uint32_t ntables,
                            double laplace, double* P1) {
    if (!edge_counts || !next || !P1 || S == 0 || R == 0 || R > 16) return MVD_E_INVALID;
    const double lapS = laplace * (double)S;                       // one rounding, as `laplace * S` in Python
    host_parallel((uint64_t)S * ntables, 1u << 13, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t q = lo; q < hi; ++q) {
            const uint64_t i = q % S;
            const uint64_t* c = edge_counts + q * R;
            const uint32_t* nx = next + i * R;
            double row = 0.0;
            for (uint32_t r = 0; r < R; ++r) row += (double)c[r];  // integers: exact in any order
            const double denom = row + lapS;
            for (uint32_t r = 0; r < R; ++r) {
                double cij = 0.0;
                for (uint32_t r2 = 0; r2 < R; ++r2)
                    if (nx[r2] == nx[r]) cij += (double)c[r2];
                P1[q * R + r] = (cij + laplace) / denom;
            }
        }
    });
    return MVD_OK;
}

int mvd_last_kernel_ms(mvd_ctx* ctx, float* ms) {
    if (!ctx || !ms) return MVD_E_INVALID;
    *ms = ctx->last_ms;
    return MVD_OK;
}

int mvd_copy_stats(mvd_ctx* ctx, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
    if (!ctx) return MVD_E_INVALID;
    if (h2d_bytes) *h2d_bytes = ctx->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = ctx->d2h_bytes;
    return MVD_OK;
}

int mvd_launch_count(mvd_ctx* ctx, uint64_t* launches) {
    if (!ctx || !launches) return MVD_E_INVALID;
    *launches = ctx->launches;
    return MVD_OK;
}

int mvd_set_option(mvd_ctx* ctx, int option, int64_t value) {
    if (!ctx) return MVD_E_INVALID;
    if (option == MVD_OPT_FORCE_GENERIC) {
        ctx->force_generic = value != 0;
        return MVD_OK;
    }
    if (option == MVD_OPT_NO_PAIR) {
        ctx->no_pair = value == 1;
        ctx->force_pair = value == 2;
        return MVD_OK;
    }
    if (option == MVD_OPT_NO_FSM1) {
        ctx->no_fsm1 = value != 0;
        return MVD_OK;
    }
    if (option == MVD_OPT_NO_ANTIPODAL) {
        ctx->no_antipodal = value != 0;
        return MVD_OK;
    }
    if (option == MVD_OPT_ASYNC_DETECT) {
        if (!value) {
            const int rc = drain_async(ctx);
            if (rc != MVD_OK) return rc;
        }
        ctx->async_detect = value != 0;
        return MVD_OK;
    }
    if (option == MVD_OPT_SPLIT_CHUNK) {
        if (value != 0 && value != 256 && value != 512 && value != 1024) return fail(ctx, MVD_E_INVALID, "MVD_OPT_SPLIT_CHUNK takes 0, 256, 512 or 1024");
        ctx->split_chunk = (uint32_t)value;
        return MVD_OK;
    }
    if (option == MVD_OPT_SPLIT_SEQUENTIAL) {
        if (value < 0 || value > 2) return fail(ctx, MVD_E_INVALID, "MVD_OPT_SPLIT_SEQUENTIAL takes 0, 1 or 2");
        ctx->split_sequential = value == 1;
        if (ctx->no_split_classes != (value == 2)) ctx->split_tables_ready = false;    // the float32 rows differ between the modes
        ctx->no_split_classes = value == 2;
        return MVD_OK;
    }
    if (option == MVD_OPT_SPLIT) {
        if (value < 0 || value > 2) return fail(ctx, MVD_E_INVALID, "MVD_OPT_SPLIT takes 0, 1 or 2");
        ctx->split_mode = (int)value;
        return MVD_OK;
    }
    if (option == MVD_OPT_LEARN_WARM) {
        if (value < 0 || value > (1 << 20) || (value & 31)) return fail(ctx, MVD_E_INVALID, "warm-up must be a multiple of 32 in [0, 2^20]");
        ctx->learn_warm = (uint32_t)value;
        ctx->learn_warm_set = true;
        return MVD_OK;
    }
    return fail(ctx, MVD_E_INVALID, "unknown option %d", option);
}

int mvd_last_kernel_kind(mvd_ctx* ctx, int* kind) {
    if (!ctx || !kind) return MVD_E_INVALID;
    *kind = ctx->last_fast;
    return MVD_OK;
}

int mvd_learn_stats(mvd_ctx* ctx, uint32_t* dirty_chunks) {
    if (!ctx || !dirty_chunks) return MVD_E_INVALID;
    *dirty_chunks = ctx->last_dirty;
    return MVD_OK;
}

int mvd_split_stats(mvd_ctx* ctx, uint64_t* subchunks, uint64_t* sequential) {
    if (!ctx) return MVD_E_INVALID;
    if (subchunks) *subchunks = ctx->last_split_sub;
    if (sequential) *sequential = ctx->last_split_seq;
    return MVD_OK;
}

int mvd_int_peak(mvd_ctx* ctx, double* alu_gops, double* alu_fma_gops) {
    if (!ctx) return MVD_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_peak.reserve(64));
    CK(cudaMemsetAsync(ctx->d_peak.p, 0, 64, ctx->stream));
    const int blocks = ctx->prop.multiProcessorCount * 8, iters = 4096;
    double res[2] = {0, 0};
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(ctx->ev0, ctx->stream));
            CK(mvd_launch_int_peak(blocks, ctx->stream, ctx->d_peak.as<uint32_t>(), iters, mode));
            CK(cudaEventRecord(ctx->ev1, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            ctx->launches += 1;
            if (rep > 0) best = std::min(best, ms);
        }
        const double ops = (double)blocks * 256.0 * iters * MVD_PEAK_OPS_PER_ITER;
        res[mode] = ops / (best * 1e-3) * 1e-9;
    }
    if (alu_gops) *alu_gops = res[0];
    if (alu_fma_gops) *alu_fma_gops = res[1];
    return MVD_OK;
}

int mvd_device_info(mvd_ctx* ctx, int* sm_count, int* clock_khz, uint64_t* smem_per_block_optin, char* name, int name_len) {
    if (!ctx) return MVD_E_INVALID;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (clock_khz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
        *clock_khz = khz;
    }
    if (smem_per_block_optin) *smem_per_block_optin = ctx->prop.sharedMemPerBlockOptin;
    if (name && name_len > 0) {
        strncpy(name, ctx->prop.name, (size_t)name_len - 1);
        name[name_len - 1] = 0;
    }
    return MVD_OK;
}

}  // extern "C"
