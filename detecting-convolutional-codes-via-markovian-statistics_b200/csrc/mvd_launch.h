// mvd_launch.h -- the kernels live in their own translation units (compiled in parallel by build.py);
// mvd.cu reaches them through these launchers.
#pragma once
#include "mvd_types.h"

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
cudaError_t mvd_launch_learn(bool smem_tables, size_t lsmem, uint32_t nsegs, cudaStream_t st, const Params& P,
                             const LearnParams& LP);
cudaError_t mvd_launch_int_peak(int blocks, cudaStream_t st, uint32_t* out, int iters, int mode);
