/*
 * mvd.h -- C ABI of libmvd.so, the B200 (sm_100a) implementation of the reference's
 * Monte-Carlo hybrid-detector hot path.
 *
 * The reference (pure Python) has no FFI; the boundary this library sits behind is the Python
 * function API of its two entry points.  Each entry point below names the reference interface
 * it replaces (file:line into the reference repository); INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only; every function returns MVD_OK (0) or a negative
 * error code and never throws; mvd_last_error() gives the message; the library never frees or
 * retains caller memory beyond the call; all device work of a context is ordered on one CUDA
 * stream; one context per device (one process per GPU under torchrun).
 *
 * Tables are "edge-indexed": entry [i * R + r] belongs to Markov state i (BFS order,
 * viterbi_markov.py:189-192) and received word r (index in itertools.product([0,1], repeat=n),
 * viterbi_markov.py:175 -- first output bit is the MSB), R = 2^n.
 */
#ifndef MVD_H
#define MVD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVD_ABI_VERSION 3
#define MVD_MAX_N 4   /* outputs per step (R = 2^n <= 16)            */
#define MVD_MAX_M 6   /* encoder memory  (2^m <= 64 trellis states)  */
#define MVD_MAX_K 3   /* inputs per step of a code given as tables   */

enum {
    MVD_OK = 0,
    MVD_E_INVALID = -1,       /* bad argument                                             */
    MVD_E_CUDA = -2,          /* CUDA runtime error (no device, launch failure, ...)      */
    MVD_E_UNSUPPORTED = -3,   /* outside the device envelope (k > 3, n > 4, m > 6, ...)   */
    MVD_E_STATE = -4,         /* call order: tables not set                               */
    MVD_E_NOMEM = -5,
    MVD_E_UNKNOWN_STATE = -6  /* a metric vector was not in the state table: the KeyError of
                                 Pd_plotter.py:112 / an 8-bit metric lane overflow          */
};

enum { MVD_SRC_PHILOX = 0, MVD_SRC_BITSTREAM = 1 };
enum { MVD_ENGINE_AUTO = 0, MVD_ENGINE_ACS = 1, MVD_ENGINE_FSM = 2 };

typedef struct mvd_ctx mvd_ctx;

/* Where the info bits and BSC flips of a launch come from.
 * PHILOX: generated on the device, stream spec MVD-PHILOX-2 (mvd/bitsource.py), key = seed.
 * BITSTREAM: read from `bits`, an array of 128-bit words indexed
 *   seg.bits_offset + (sb * (1 + n) + c) * ntrials + (trial - seg.trial_begin)
 * (sb = step / 128, c = 0 info stream, c = 1 + j flips of output j -- k + n words per superblock, inputs first, for a code
 * given as tables --, ntrials = trial_end - trial_begin;
 * bit b of 32-bit lane w of a word is step 128*sb + 32*w + b). */
typedef struct mvd_src {
    int32_t mode;             /* MVD_SRC_*                                              */
    int32_t bits_on_device;   /* BITSTREAM: 1 = `bits` is a device pointer, 0 = host    */
    uint64_t seed;            /* PHILOX key                                             */
    const void* bits;         /* BITSTREAM words                                        */
    uint64_t bits_words;      /* BITSTREAM: number of 128-bit words in `bits`           */
} mvd_src;

/* One hypothesis of one (N, p) sweep point: the body of the trial loop Pd_plotter.py:210-223
 * for `trial_end - trial_begin` iterations, or one learning chain (Pd_plotter.py:149-163). */
typedef struct mvd_segment {
    uint32_t N;                      /* trellis steps per trial / chain length            */
    uint32_t threshold;              /* PHILOX: P(flip) = threshold / 2^32                */
    uint32_t stream;                 /* PHILOX: stream tag (counter word 3)               */
    uint32_t table;                  /* which log-likelihood table set scores this trial  */
    uint32_t enc_taps[MVD_MAX_N];    /* encoder of this hypothesis: bit t of enc_taps[j] =
                                        tap of output j on the input t steps ago          */
    uint32_t decide;                 /* 0: success iff logp1 >  logp_ref (Pd_plotter.py:215)
                                        1: success iff logp1 <= logp_ref (Pd_plotter.py:222) */
    uint32_t random_input;           /* 1 = uniform info bits, 0 = all-zero input         */
    uint64_t trial_begin, trial_end; /* global trial ids [begin, end)                     */
    uint64_t bits_offset;            /* BITSTREAM: first word of this segment             */
} mvd_segment;

int mvd_abi_version(void);

/* Create / destroy a context on CUDA device `device`.  Fails with MVD_E_CUDA when no usable
 * device exists -- there is no CPU fallback. */
int mvd_create(mvd_ctx** out, int device);
int mvd_destroy(mvd_ctx* ctx);
const char* mvd_last_error(const mvd_ctx* ctx);   /* ctx may be NULL: last create error */
int mvd_set_stream(mvd_ctx* ctx, void* cuda_stream);
int mvd_synchronize(mvd_ctx* ctx);

/* Decoder code (always H1's generator, Pd_plotter.py:188): replaces build_trellis
 * (viterbi_markov.py:118-132) + branch_output_and_next_state (:82-106) for k = 1.
 * dec_taps[j] bit t = generator_matrix[j][0][t]. */
int mvd_set_code(mvd_ctx* ctx, int k, int n, int m, const uint32_t* dec_taps);

/* The same for codes GIVEN AS TABLES, any k <= MVD_MAX_K inputs per step: branch_output_and_next_state
 * (viterbi_markov.py:82-106) and build_trellis (:118-132) are generic in k in the reference -- including the register
 * [u_i, s_0, ..] that every input i sees -- and whatever they produce can be handed over as it is:
 *   dec_prev / dec_label [2^m][2^k]: predecessor state and branch label (first output = MSB) of the incoming branches of
 *   every trellis state, in the reference's insertion order (ps ascending, inputs in itertools.product order);
 *   mvd_set_encoders: enc_next / enc_out [nenc][2^m][2^k] for the encoders of the hypotheses, input tuple (u_0 .. u_{k-1})
 *   at index u_0 2^(k-1) + .. + u_{k-1}; a segment's enc_taps[0] is then the INDEX of its encoder.
 * The device path of such a code is the Markov-state walk (MVD_ENGINE_FSM, what MVD_ENGINE_AUTO selects): learning chains,
 * detection trials and traces run on the generic kernels with a table-driven encoder; the info bits of input i come from
 * Philox slot 32 + i (mvd/bitsource.py) or from info stream c = i of the k + n per superblock.  mvd_set_states checks the
 * table against Eq. 4-5 on this trellis (closure).  MVD_ENGINE_ACS, mvd_enumerate_states(_gpu) and mvd_acs_hash need the
 * tap-mask form (k = 1) and return MVD_E_UNSUPPORTED. */
int mvd_set_code_tables(mvd_ctx* ctx, int k, int n, int m, const uint8_t* dec_prev, const uint8_t* dec_label);
int mvd_set_encoders(mvd_ctx* ctx, uint32_t nenc, const uint8_t* enc_next, const uint8_t* enc_out);

/* Markov-state table: replaces the `states` list and the `state_index` dict
 * (viterbi_markov.py:166-195, Pd_plotter.py:136-139).  metrics: S x 2^m bytes, next: S x 2^n. */
int mvd_set_states(mvd_ctx* ctx, uint32_t S, const uint8_t* metrics, const uint32_t* next);

/* Host-side (C++) breadth-first enumeration with the reference's discovery order; an
 * accelerated enumerate_markov_states_allzero (viterbi_markov.py:166-195).  Installs the table
 * like mvd_set_states.  mvd_get_states copies it out (either pointer may be NULL). */
int mvd_enumerate_states(mvd_ctx* ctx, uint32_t max_states, uint32_t* S_out);
int mvd_get_states(mvd_ctx* ctx, uint8_t* metrics, uint32_t* next);

/* The same enumeration on the GPU (csrc/mvd_bfs.cuh): the queue of viterbi_markov.py:183-193 is
 * expanded chunk-parallel, duplicates meet in an open-addressing table in HBM and ties are broken
 * by candidate rank (parent index, received word), so state indices and NEXT are identical to the
 * reference's discovery order.  flags: MVD_BFS_INSTALL = install the table like mvd_set_states
 * (then mvd_get_states works); MVD_BFS_COUNT_ONLY = keep no NEXT table (memories whose state set
 * is too large to use, m = 5, 6: how many states are there?).  chunk_parents: queue entries
 * expanded per pass, 0 = default.  When max_states (<= 1 610 612 736) is exceeded the call returns
 * MVD_E_NOMEM and stats->S is the number enumerated so far (a lower bound), stats->closed = 0. */
enum { MVD_BFS_INSTALL = 1, MVD_BFS_COUNT_ONLY = 2 };
typedef struct mvd_bfs_stats {
    uint32_t S;             /* states enumerated                                      */
    uint32_t frontier;      /* queue position reached (== S when closed)              */
    uint32_t iterations;    /* chunk passes                                           */
    uint32_t launches;      /* kernels launched                                       */
    uint64_t candidates;    /* (state, received word) pairs expanded                  */
    int32_t closed;         /* 1 = the queue ran empty: S is the size of the state set */
    int32_t max_metric;     /* largest relative metric seen                           */
    float ms;               /* device time of the whole enumeration (CUDA events)     */
    uint32_t reserved;
} mvd_bfs_stats;
int mvd_enumerate_states_gpu(mvd_ctx* ctx, uint32_t max_states, uint32_t flags, uint32_t chunk_parents,
                             mvd_bfs_stats* stats);
/* Sizes of the BFS levels the last mvd_enumerate_states_gpu completed (sizes[d] = states first
 * reached after d received words; sizes may be NULL to query *nlevels). */
int mvd_bfs_levels(mvd_ctx* ctx, uint32_t* sizes, uint32_t cap, uint32_t* nlevels);

/* Log-likelihood tables: logP1[t][i*R + r] = log(max(P1_t[i, next[i][r]], 1e-300)) for table
 * set t (one per distinct p; Pd_plotter.py:166-167 then :114-115), logTref likewise for
 * T(p = 1/2) (Pd_plotter.py:193-194).  Host float64, computed by the caller with libm log. */
int mvd_set_loglik(mvd_ctx* ctx, uint32_t ntables, const double* logP1, const double* logTref);

/* Host helper: out[i] = log(max(values[i], 1e-300)) with the C library's log -- the function
 * math.log calls per step at Pd_plotter.py:114-115, so tables built with it carry the reference's own
 * terms.  Runs on the calling thread (no device work); for S 2^n of 10^5..10^6 entries per table. */
int mvd_host_log_table(const double* values, double* out, uint64_t count);
/* P1 edge tables from edge counts, the closed form of Pd_plotter.py:166-167 on the edges (what numpy computes for the
 * dense S x S matrix, gathered at (i, next[i][r])): P1[t][i][r] = (sum of counts[t][i][r'] over the r' with
 * next[i][r'] == next[i][r] + laplace) / (sum_r counts[t][i][r] + laplace * S), float64, one division per entry.
 * Host code on up to 16 threads (as mvd_host_log_table): at S = 150 743 the two numpy statements took 66 ms for seven
 * tables.  next = state indices [S][R] (not premultiplied). */
int mvd_host_p1_edge_tables(const uint64_t* edge_counts, const uint32_t* next, uint32_t S, uint32_t R, uint32_t ntables,
                            double laplace, double* P1);

/* Transition counting: the loop Pd_plotter.py:158-163 for every segment (segment = one or more
 * chains of N steps; only steps t >= burn are counted).  Single-chain segments on the on-device
 * bit source with engine AUTO/FSM are walked chunk-parallel (exact; see csrc/mvd_learn2.cuh).  edge_counts: host, nsegs x S x R,
 * edge_counts[s][i*R + r] = number of counted steps leaving state i on received word r. */
int mvd_learn_counts(mvd_ctx* ctx, const mvd_src* src, const mvd_segment* segs, uint32_t nsegs,
                     uint32_t burn, int engine, uint64_t* edge_counts);

/* Detection trials: Pd_plotter.py:210-223.  tallies: host, nsegs (successes per segment).
 * logp (optional, host): 2 doubles (logp1, logp_ref) per trial, segments concatenated.
 * d_tallies (optional, device uint64[nsegs]): receives a copy of this call's tallies on the
 * context's stream (complete when the call returns), for a device-side allreduce; then `tallies`
 * may be NULL and nothing but the 4-byte error flag crosses PCIe.  Pass NULL otherwise. */
int mvd_detect(mvd_ctx* ctx, const mvd_src* src, const mvd_segment* segs, uint32_t nsegs,
               int engine, uint64_t* tallies, double* logp, void* d_tallies);

/* Trajectory trace of one segment (verification): replaces simulate_markov_sequence(...)["metrics"]
 * (call sites Pd_plotter.py:149,212,219).  state_idx: host, ntrials x (N+1) uint32;
 * metrics (optional): host, ntrials x (N+1) x 2^m bytes (ACS engine: the registers themselves;
 * FSM engine: gathered from the state table). */
int mvd_trace(mvd_ctx* ctx, const mvd_src* src, const mvd_segment* seg, int engine,
              uint32_t* state_idx, uint8_t* metrics);

/* Eq. 4-5 recursion without a state table (memories whose Markov state set cannot be
 * enumerated, m = 5, 6): per trial a 64-bit FNV-1a hash over the metric bytes of D_1..D_N and
 * the final vector D_N.  hashes: host, ntrials; final_metrics (optional): ntrials x 2^m. */
int mvd_acs_hash(mvd_ctx* ctx, const mvd_src* src, const mvd_segment* seg,
                 uint64_t* hashes, uint8_t* final_metrics);

/* Error exponent, Eq. 7: rho[q] = spectral radius of M(u_q), M(u)[i,j] = sum_r P1(i->j,r)^u P2(i->j,r)^(1-u)
 * -- the eigenvalue loop of compute_error_exponent (alpha_exponent.py:155-184) for two Laplace-smoothed
 * joint tensors of the same decoder (learn_transition_tensor, alpha_exponent.py:83-149), given in edge
 * form: lp_h[i*R + r] = log P_h(i -> next[i*R + r], r) and lb_h[i] = log of row i's background entry
 * (lambda / d_i; every (j, r) that is not an edge).  Power iteration on the sparse + rank-one form, one
 * thread block per u; stops when the estimate moves by <= tol * rho or after max_iter products.
 * next: K x R state indices (host); rho: nu doubles; iters (optional): products taken per u.
 * Needs no code / state table in the context. */
int mvd_chernoff_rho(mvd_ctx* ctx, uint32_t K, uint32_t R, const uint32_t* next, const double* lp1, const double* lp2,
                     const double* lb1, const double* lb2, const double* u_vals, uint32_t nu, double tol,
                     uint32_t max_iter, double* rho, uint32_t* iters);

/* The same for dense K x K x R tensors (the objects alpha_exponent.py:155-184 takes):
 * logP_h[(i*K + j)*R + r] = log(clip(P_h[i,j,r], 1e-300, 1)) (alpha_exponent.py:167-168). */
int mvd_chernoff_rho_dense(mvd_ctx* ctx, uint32_t K, uint32_t R, const double* logP1, const double* logP2,
                           const double* u_vals, uint32_t nu, double tol, uint32_t max_iter, double* rho,
                           uint32_t* iters);

/* Parity-template baseline detector (paper section IV): the Monte-Carlo loop of comp_parity.py:165-176.
 * One trial = N info bits -> encode_convolutional (comp_parity.py:65-86: n streams of N + m bits, zero
 * tail) -> BSC -> parity_satisfaction_fraction over t in [max_delay, N + m) (:93-116) -> decide H1 iff the
 * fraction >= gamma (parity_detector, :123-132).  Bit sources as for mvd_detect; in BITSTREAM mode a trial
 * supplies (1 + n) streams of N + m bits in the layout of mvd_src (info bits beyond N are ignored).
 * tallies: host, nsegs (successes per segment); satisfied (optional, host): satisfied positions per trial,
 * segments concatenated.  Needs no code / state table in the context. */
typedef struct mvd_parity_segment {
    uint32_t N;                      /* info bits per trial                                        */
    uint32_t m;                      /* encoder memory = length of the zero tail                   */
    uint32_t n;                      /* output streams                                             */
    uint32_t threshold;              /* PHILOX: P(flip) = threshold / 2^32                         */
    uint32_t stream;                 /* PHILOX: stream tag                                         */
    uint32_t decide;                 /* 0: success iff decided H1 (fraction >= gamma), 1: iff H2   */
    uint32_t enc_taps[MVD_MAX_N];    /* bit t = tap of output j on the input t steps ago           */
    uint32_t tmpl[MVD_MAX_N];        /* bit s of tmpl[j] = template term (j, s): y_j[t - s]        */
    double gamma;                    /* decision threshold (comp_parity.py:161)                    */
    uint64_t trial_begin, trial_end; /* global trial ids [begin, end)                              */
    uint64_t bits_offset;            /* BITSTREAM: first word of this segment                      */
} mvd_parity_segment;
int mvd_parity_detect(mvd_ctx* ctx, const mvd_src* src, const mvd_parity_segment* segs, uint32_t nsegs,
                      uint64_t* tallies, uint32_t* satisfied);

/* Throughput form of the same recursion (n = 2, on-device bits): two trials per thread, all 2^m metric pairs
 * in registers (csrc/mvd_acsp.cuh); only the final vectors D_N come back (ntrials x 2^m bytes), the same bytes
 * mvd_acs_hash returns in final_metrics.  BASELINE config 4: m = 4..6, up to 64 trellis states. */
int mvd_acs_final(mvd_ctx* ctx, const mvd_src* src, const mvd_segment* seg, uint8_t* final_metrics);

/* Timing of the last learn/detect/trace launch on the context's stream (CUDA events), and the
 * number of kernels this library has launched since creation. */
int mvd_last_kernel_ms(mvd_ctx* ctx, float* ms);
int mvd_launch_count(mvd_ctx* ctx, uint64_t* launches);
/* Bytes this context has copied host -> device and device -> host since creation, counted at the
 * cudaMemcpyAsync call sites of the library (tables, segment descriptors, bit streams in; tallies, counts,
 * log-likelihoods, error flags out).  bench.py reports the per-step difference as e2e.h2d/d2h_bytes_per_step. */
int mvd_copy_stats(mvd_ctx* ctx, uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* Kernel time (CUDA events around each launch, summed) and number of the asynchronous detection launches
 * (MVD_OPT_ASYNC_DETECT) drained since the last call of this function; both counters are reset. */
int mvd_async_stats(mvd_ctx* ctx, double* kernel_ms_sum, uint64_t* launches);

/* Options.  MVD_OPT_FORCE_GENERIC (value 0/1): 1 = never take the fast detection kernels
 * (mvd_detect2.cuh), always the generic checked ones -- used by the parity tests to cover both.
 * mvd_last_kernel_kind: 0 = the last launch was a generic kernel, otherwise
 * 1 + lookup (0 direct table, 1 hash table, 2 NEXT-table walk, 3 one-load NEXT-table walk) + 16 * log2(bytes per log-likelihood row entry)
 * + 256 if the two-trials-per-thread kernel ran, + 512 if the tables stayed in global memory (large S);
 * 1024 = chunk-parallel learning chain, 2048 = GPU state enumeration, 4096 = Chernoff spectral radius, 8192 = parity-template trials,
 * 32768 = mvd_acs_final,
 * 16384 = detection trials split along the time axis (mvd_learn_stats then gives the chunks repaired, mvd_split_stats the
 * share of the float64 additions that were re-associated).
 * mvd_learn_stats: chunks of the last chunk-parallel learning call whose speculated start state was
 * wrong and had to be repaired (results are exact either way; this is a performance counter). */
enum { MVD_OPT_FORCE_GENERIC = 1, MVD_OPT_NO_PAIR = 2,     /* NO_PAIR: 1 = one trial per thread, 2 = two per thread
                                                               even for few trials, 0 = automatic              */
       MVD_OPT_LEARN_WARM = 3,     /* warm-up steps of the chunk-parallel chains (default 128 for m <= 3, 128 (m - 1) above) */
       MVD_OPT_NO_FSM1 = 4,        /* 1 = NEXT-table walk with separate log / NEXT tables (two loads per step) */
       MVD_OPT_SPLIT = 5,          /* few long trials (NEXT-table engine, on-device bits) are split along the time axis
                                      (csrc/mvd_split.cuh; identical results): 0 = when it fills the GPU better,
                                      1 = whenever possible, 2 = never.  The warm-up is MVD_OPT_LEARN_WARM's. */
       MVD_OPT_NO_ANTIPODAL = 6,   /* 1 = the two-trials-per-thread m = 2 kernel reads the general branch-metric table even
                                      when every decoder generator has its first and last tap set (the complement-label
                                      short cut; identical results) */
       MVD_OPT_ASYNC_DETECT = 7, /* 1 = mvd_detect calls that ask for device tallies only (tallies == NULL, logp == NULL,
                                      d_tallies != NULL) return as soon as their work is queued on the context's stream;
                                      mvd_synchronize (or any call that reads results, or setting the option back to 0)
                                      waits for them and reports a KeyError of any of them.  At most 64 are kept in
                                      flight.  The segment records and tally words of a sweep stay on the device, so a
                                      loop of sweeps needs no host round trip between them (Pd_plotter.py:210-223
                                      repeated; bench.py's resident leg) */
       MVD_OPT_SPLIT_SEQUENTIAL = 8, /* 1 = the split path adds every log-likelihood term one by one in step order instead of
                                      re-associating the additions inside a binade; 2 = re-association with both sums as
                                      recurrences even when log Tref has <= 3 distinct values (no class counting); identical
                                      results either way: the checks of 0 */
       MVD_OPT_SPLIT_CHUNK = 9 };  /* steps per chunk of the split path: 0 = chosen per call, else 256, 512 or 1024 (identical results) */
int mvd_set_option(mvd_ctx* ctx, int option, int64_t value);
int mvd_last_kernel_kind(mvd_ctx* ctx, int* kind);
int mvd_learn_stats(mvd_ctx* ctx, uint32_t* dirty_chunks);
/* Split path (few long trials, kind 16384; Pd_plotter.py:106-116 for N = 10^4 .. 10^5): the two float64 sums of a trial are formed
 * from partial sums over 128-step sub-chunks wherever the running sum provably stays inside one binade (bit-identical to the
 * step-by-step additions, see csrc/mvd_split.cuh) and term by term elsewhere.  subchunks = sub-chunks of the last split launch,
 * sequential = those added term by term (performance counters). */
int mvd_split_stats(mvd_ctx* ctx, uint64_t* subchunks, uint64_t* sequential);

/* Integer roofline denominators, measured on this device (32-bit lane-ops/s over all SMs):
 * alu_gops     -- dependent-free LOP3 chains: the ALU pipe alone (min/shift/logic/permute issue only there);
 * alu_fma_gops -- alternating LOP3 / IMAD: ALU + FMA pipes together = the warp-instruction issue rate. */
int mvd_int_peak(mvd_ctx* ctx, double* alu_gops, double* alu_fma_gops);

/* Device facts used by the host for sharding and reporting. */
int mvd_device_info(mvd_ctx* ctx, int* sm_count, int* clock_khz, uint64_t* smem_per_block_optin,
                    char* name, int name_len);

#ifdef __cplusplus
}
#endif
#endif /* MVD_H */
