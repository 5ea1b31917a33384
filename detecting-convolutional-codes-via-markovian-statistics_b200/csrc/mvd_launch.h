// mvd_launch.h -- the kernels live in their own translation units (compiled in parallel by build.py);
// mvd.cu reaches them through these launchers.
#pragma once
#include "mvd_types.h"

#include <vector>

cudaError_t mvd_launch_generic(int engine, int mode, bool n2, int m, bool in_smem, dim3 grid, size_t smem, cudaStream_t st,
                               const Params& P);
cudaError_t mvd_launch_generic_acs(int mode, bool n2, int m, dim3 grid, size_t smem, cudaStream_t st, const Params& P);
cudaError_t mvd_launch_generic_fsm(int mode, bool n2, bool in_smem, dim3 grid, size_t smem, cudaStream_t st, const Params& P);
cudaError_t mvd_launch_det2(int lk, int m, int lls, bool gt, bool pair, dim3 grid, unsigned threads, size_t smem,
                            cudaStream_t st, const Params& P, const SegBatch& B);
cudaError_t mvd_launch_det2_acs(int lk, int m, int lls, bool gt, dim3 grid, unsigned threads, size_t smem, cudaStream_t st,
                                const Params& P, const SegBatch& B);
cudaError_t mvd_launch_det2_fsm(int lk, int lls, bool gt, dim3 grid, unsigned threads, size_t smem, cudaStream_t st,
                                const Params& P, const SegBatch& B);
cudaError_t mvd_launch_det2_pair(dim3 grid, unsigned threads, size_t smem, cudaStream_t st, const Params& P, const SegBatch& B);
cudaError_t mvd_launch_det3_pair(int m, dim3 grid, unsigned threads, size_t smem, cudaStream_t st, const Params& P, const SegBatch& B,
                                 bool ds, bool anti);
cudaError_t mvd_launch_slot_rows(const double2* ll, const uint32_t* pht, uint32_t slots, uint32_t SR, uint32_t ntables, double2* out,
                                 cudaStream_t st);
cudaError_t mvd_launch_pack_ll(const double* lp1, const double* ltref, const uint32_t* nxt, const uint32_t* tcode, uint32_t SR,
                               uint32_t ntables, double2* ll, uint4* gfsm1, cudaStream_t st);
cudaError_t mvd_launch_learn(bool smem_tables, size_t lsmem, uint32_t nsegs, cudaStream_t st, const Params& P,
                             const LearnParams& LP);
cudaError_t mvd_launch_int_peak(int blocks, cudaStream_t st, uint32_t* out, int iters, int mode);

// GPU breadth-first enumeration of the Markov states (mvd_tu_bfs.cu)
struct MvdBfsConfig {
    int n = 0, m = 0;
    uint32_t dec_taps[MVD_MAX_N] = {0, 0, 0, 0};
    uint32_t max_states = 0;
    uint32_t chunk_parents = 0;     // queue entries expanded per pass (0 = default 2^22)
    bool keep_next = true;          // false: count only (no NEXT table)
    bool copy_out = true;           // copy metrics / NEXT to the host vectors of the result
};
struct MvdBfsResult {
    uint32_t S = 0, frontier = 0, iterations = 0, launches = 0;
    uint64_t candidates = 0;
    bool closed = false;
    int max_metric = 0;
    float ms = 0.f;
    char error[256] = {0};
    std::vector<uint8_t> metrics;
    std::vector<uint32_t> next;
    std::vector<uint32_t> levels;   // states first reached at BFS depth d (complete levels only)
};
#define MVD_BFS_MAX_STATES 1610612736u   /* 3/4 of the 2^31 slots a 32-bit slot word can address */
int mvd_bfs_run(const MvdBfsConfig& cfg, cudaStream_t st, MvdBfsResult& res);

// Chernoff spectral radius (mvd_tu_chernoff.cu, mvd_chernoff.cuh)
#define CHERNOFF_BLOCK 512

struct ChernoffParams {
    uint32_t K, R, nu, max_iter;
    const uint32_t* nxt;       // [K * R] next state index
    const double* lp1;         // [K * R] log P1 on the edges
    const double* lp2;
    const double* lb1;         // [K] log background of row i under P1
    const double* lb2;
    const double* u_vals;      // [nu]
    double tol;
    double* wd;                // scratch [nu][K * R]   w - bg
    double* bgR;               // scratch [nu][K]       R * bg
    double* xa;                // scratch [nu][K]
    double* xb;                // scratch [nu][K]
    double* rho;               // [nu]
    uint32_t* iters;           // [nu]
};

cudaError_t mvd_launch_chernoff(const ChernoffParams& P, cudaStream_t st);
cudaError_t mvd_launch_chernoff(const ChernoffParams& P, cudaStream_t st);
cudaError_t mvd_launch_chernoff_dense(const ChernoffParams& P, cudaStream_t st);

// parity-template baseline (mvd_tu_parity.cu, mvd_parity.cuh)
cudaError_t mvd_launch_parity(dim3 grid, cudaStream_t st, const Params& P, const ParityBatch& B, uint32_t* satisfied);

// detection trials split along the time axis (mvd_split.cuh, in mvd_tu_learn.cu): walk, repair + predict, partial sums, score --
// four kernels in a row on `st`; the *_bytes are their dynamic shared memory
cudaError_t mvd_launch_split(size_t walk_bytes, size_t isum_bytes, size_t score_bytes, cudaStream_t st, const Params& P, const SplitParams& SP);
// per (table, edge): tie binades, float32 terms, "some term is positive" flag
cudaError_t mvd_launch_split_tables(const double2* ll, uint32_t SR, uint32_t ntables, uint32_t* tie, float2* apx, uint32_t* flags,
                                    unsigned long long* tiek, const SplitClasses& cls, cudaStream_t st);

// Eq. 4-5 at m = 2..6, two trials per thread, final metric vectors only (mvd_tu_acsp.cu)
cudaError_t mvd_launch_acsp(int m, dim3 grid, unsigned threads, cudaStream_t st, const Params& P, const DevSeg& sg,
                            const uint32_t* sel, uint8_t* final_met);
