// mvd_types.h -- types and constants shared by the host side (mvd.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mvd.h"

#define MVD_BLOCK 256
#define MVD_EMPTY 0xFFFFFFFFu

enum { MODE_DETECT = 0, MODE_LEARN = 1, MODE_TRACE = 2, MODE_HASH = 3 };

struct DevSeg {
    uint32_t N, threshold, stream, table;
    uint32_t enc_taps[MVD_MAX_N];
    uint32_t decide, random_input, dmin, block_begin;
    unsigned long long trial_begin, trial_end, bits_offset, out_offset;
};

// shared-memory plan of the fast detection kernels (mvd_detect2.cuh)
struct FastPlan {
    uint32_t off_tb, off_bm, off_st, off_ll;   // byte offsets into dynamic shared memory
    uint32_t key_mul, nkeys;           // direct metric-vector -> state table (m <= 2)
    const uint16_t* dstate;            // [nkeys] state index or 0xFFFF
    const uint16_t* dstate2;           // [256] the same through the pair kernel's offset-invariant key (m = 2)
    uint32_t kc[4], kcb;               //   key * 128 = sum kc[s] (128 D[s]) + bias * 128, both 16-bit lanes (kcb carries the bias twice)
    const uint32_t* tcode;             // packed NEXT walk: high word of the double c with log Tref[e] = c * tref_unit
    double tref_unit;
    const uint4* gfsm1;                // [ntables][S*R] packed NEXT-walk entries in global memory (large S)
    const uint32_t* ph_d;              // m = 3 perfect hash (mvd_detect3p.cuh): displacement per bucket [256]
    const uint32_t* ph_t;              //   slot -> state * R, or MVD_EMPTY [ph_slots]
    uint32_t ph_slots;                 //   power of two, 0 = no perfect hash
    uint32_t ph_bshift;                //   bucket = hash >> ph_bshift
    uint32_t ph_c2, ph_c4;             //   multipliers of the second hash (chosen by the host build)
    uint32_t ph_nb;                    //   number of buckets (= entries of ph_d)
    uint32_t ph_slot0;                 //   slot of Markov state 0 (the all-zero vector)
    uint32_t kq[6];                    //   16, 256, 4096, -4369, 1, -4368: multipliers of the offset-invariant key words, handed over as
                                       //   launch-time values so that the key is formed by IMADs on the FMA pipe, not by ALU-pipe shifts
    const double2* ll_slot;            // m = 4: log-likelihood rows indexed by hash SLOT, [ntables][ph_slots * R] (no slot -> state read)
};

struct Params {
    int n, m, R, nstate;
    int k;                          // inputs per step (1 unless the code was given as tables)
    const uint16_t* enc_tab;        // codes given as tables: [nenc][2^m][2^k] next state << 8 | output label, else NULL
    uint32_t S, SR;                 // SR = S * R
    int src_mode;
    uint32_t rk0[10], rk1[10];      // Philox round keys (key + r * Weyl), shared by every thread
    const uint4* bits;
    const DevSeg* segs;
    uint32_t nsegs;
    // state machine / log-likelihood tables (global memory masters)
    const uint32_t* nxt;            // [SR]  next_state * R
    const double2* ll;              // [ntables][SR]  {log P1, log Tref}
    // ACS constants
    const uint32_t* bm;             // [R][2 * NP] branch metrics, 16x2 packed
    // -1 as a launch-time constant: a multiplier the compiler cannot turn back into an ALU-pipe subtraction
    uint32_t fma_km1 = 0xFFFFFFFFu;
    int bm_antipodal;               // every decoder tap mask has bit 0 and bit m set: the labels of a butterfly are X, ~X, ~X, X
    const uint32_t* hkeys;          // [KW][hcap] nibble-packed metric keys
    const uint32_t* hvals;          // [hcap] state * R, or MVD_EMPTY
    uint32_t hcap;                  // power of two
    int tables_in_smem;             // FSM: NX/LL staged in shared memory; ACS: hash + LL staged
    // outputs
    unsigned long long* tallies;
    unsigned long long* tallies2;
    double* logp;
    unsigned long long* counts;     // [nsegs][SR]
    uint32_t burn;
    uint32_t* trace_idx;
    uint8_t* trace_met;
    unsigned long long* hashes;
    uint8_t* final_met;
    int* error_flag;
    FastPlan fp;
};


// ---- fast detection kernels (mvd_detect2.cuh)
#define DET2_BLOCK 512
#define DET2_BIG_BLOCK 768       // one-load NEXT walk whose 4-copy table fits twice per SM: two blocks of 768 threads keep the SM as full as three of 512
#define DET2P_BLOCK 256
#define DET2P_QUEUES ((DET2P_BLOCK / 32) * 1024)   // straggler queues of the pair kernel: 128 items x 8 bytes per warp
#define DET3P_BLOCK 768         // m = 3 pair kernel: one block per SM shares conflict-poor table replicas (mvd_detect3p.cuh)
#define DET3P_QUEUES ((DET3P_BLOCK / 32) * 1024)
#define DET2_MAXSEG 96          // segments per launch: they travel in the kernel parameters (uniform registers)

struct SegBatch {
    DevSeg s[DET2_MAXSEG];
};

enum { LK_DIRECT = 0, LK_HASH = 1, LK_FSM = 2, LK_FSM1 = 3 };

// ---- chunk-parallel learning chains (mvd_learn2.cuh)
#define LEARN_CH 128u       // least steps per chunk (one info-word call); LearnParams.chunk = 128, 256, 512 or 1024
#define LEARN_WARM 128u     // default warm-up steps (LearnParams.warm; multiples of 32)
#define LEARN_BLOCK 128
#define LEARN_HOT_LOG2 10      // per-block table of hot edges in front of the global counts (mvd_learn2.cuh)
#define LEARN_HOT (1u << LEARN_HOT_LOG2)

struct LearnParams {
    uint32_t nchunks;                 // per segment
    uint32_t chunk;                   // steps per chunk (a multiple of LEARN_CH): long chains take longer chunks, the warm-up
                                      // (walked for every chunk, counted for none) weighs less
    uint32_t warm;                    // warm-up steps before a chunk (multiple of 32)
    uint32_t* spec_start;             // [nsegs][nchunks]  state * R at the chunk start (speculated)
    uint32_t* end;                    // [nsegs][nchunks]  state * R at the chunk end
    uint32_t* ndirty;                 // [nsegs]
    int nxt_in_smem;
};

// ---- detection trials split along the time axis (mvd_split.cuh)
#define SPLIT_CH_MAX 1024u                  // most steps per chunk (one walker thread); SplitParams.chunk = 256, 512 or 1024
#define SPLIT_SUB 128u                      // steps per sub-chunk (one prediction / one re-associated partial sum)
#define SPLIT_BLOCK 128
#define SPLIT_RB_HOST 16                    // = SPLIT_RB of mvd_split.cuh: sub-chunk records per batch of the scoring kernel's ring

struct SplitClasses {                        // class mode of the split path (mvd_split.cuh): log Tref has <= 3 distinct non-zero values
    int n;                                  // 0 = no class mode
    double val[3];
};

struct SplitParams {
    SplitClasses cls;
    uint32_t warm;                          // warm-up steps (multiple of 32)
    uint32_t chunk;                         // steps per chunk (multiple of SPLIT_SUB)
    uint32_t nchains;
    unsigned long long nwork;               // (chain, chunk) pairs
    const unsigned long long* work_begin;   // [nsegs + 1] first work item of a segment (items: chunk-major, trial-minor)
    const unsigned long long* edge_begin;   // [nsegs] first 32-bit word of a segment's 16-byte edge groups
    const unsigned long long* sub_begin;    // [nsegs] first sub-chunk record of a segment (records: trial-major, a trial's sub-chunks consecutive)
    uint32_t* spec_start;                   // [nwork] state * R at the chunk start (speculated)
    uint32_t* end;                          // [nwork] state * R at the chunk end
    uint32_t* edges;
    uint32_t* ndirty;                       // [1] chunks repaired (performance counter)
    unsigned long long* nseq;               // [1] sub-chunks whose terms were added one by one (performance counter)
    float2* apx;                            // [sub-chunk records] float32 estimate of the two sums over the sub-chunk
    uint32_t* plan;                         // [sub-chunk records] predicted binades (SPLIT_PLAN_FAST | k1 | k0 << 11) or 0
    double2* res;                           // [sub-chunk records] what the sub-chunk adds to each sum inside the predicted binade
    const float2* apxtab;                   // [ntables][SR] {log P1, log Tref} in float32
    const uint32_t* tietab;                 // [ntables][SR] k1 | k0 << 8: the binade in which each of the two terms is a round-half-even tie (0xFF: none)
    const unsigned long long* tiek;         // [ntables][2] binades (bit k) in which some log P1 / log Tref term of the table is a tie
    const uint32_t* flags;                  // [1] bit 0: some term is positive / not a number (no predictions)
    int sequential;                         // 1 = no predictions: every term added one by one (MVD_OPT_SPLIT_SEQUENTIAL)
    int nxt_in_smem, ll_in_smem;
    int ll_rep_shift;                       // log2 copies of a log-likelihood row in shared memory (0 or 3)
    int apx_rep_shift;                      // log2 copies of a float32 row in the fast walk's shared memory (0 or 4)
    uint32_t walk_apx_offset;               // fast walk: byte offset of the float32 rows in dynamic shared memory
    uint32_t isum_tie_offset;               // split_isum_kernel: byte offset of the tie rows in dynamic shared memory
    uint32_t score_ring_offset;             // split_score_kernel: byte offset of the record ring in dynamic shared memory
    uint32_t chain_block;                   // threads per block of the scoring kernel
    unsigned long long max_trials;          // most trials of any segment (x extent of the scoring grid)
    int edge_bytes;                         // bytes per stored edge index: 1 (S R <= 256), 2 (<= 65 536) or 4
    int fast_walk;                          // 1: split_walk2_kernel (n = 2, warm-up a multiple of 128)
    unsigned long long max_chunks;          // most chunks of any segment
};

// ---- parity-template baseline trials (mvd_parity.cuh)
#define PARITY_BLOCK 256
#define PARITY_MAXSEG 64

struct ParitySeg {
    uint32_t N, m, n, threshold, stream, dmin, decide, max_delay;
    uint32_t enc_taps[MVD_MAX_N];
    uint32_t tmpl[MVD_MAX_N];        // bit s of tmpl[j]: term (j, s) of the template
    double gamma;
    unsigned long long trial_begin, trial_end, bits_offset, out_offset;
    uint32_t seg_index, pad;
};

struct ParityBatch {
    ParitySeg s[PARITY_MAXSEG];
};

#define MVD_PEAK_OPS_PER_ITER 64
