// translation unit: generic NEXT-table kernels (mvd_kernels.cuh) and the integer-peak micro-kernel
#include "mvd_kernels.cuh"
#include "mvd_launch.h"

// ------------------------------------------------------------------------------------------ integer peak
// Dependent chains on 8 independent accumulators per thread.
//   mode 0: LOP3 only -- the ALU pipe alone (LOP3/SHF/PRMT/VIMNMX/IADD3 issue there, 16 lanes/clk/SMSP);
//   mode 1: alternating IMAD (FMA pipe) and LOP3 (ALU pipe) -- both integer-capable pipes, i.e. the
//           issue-rate bound of 1 warp instruction / clk / SMSP.
// OPS_PER_ITER instructions per loop.
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* out, int iters, int mode) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 2654435761u + i + blockIdx.x;
    const uint32_t c = out[0] | 1u;           // runtime values, prevent constant folding
    const uint32_t d = out[2] | 0x10u;
    if (mode == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < MVD_PEAK_OPS_PER_ITER / 8; ++u) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (u & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(d));   // xor3
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(c), "r"(d));        // majority
                }
            }
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < MVD_PEAK_OPS_PER_ITER / 8; ++u) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(d));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(d));
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    if (s == 0x12345678u) out[1] = s;          // practically never true: keeps the chains alive
}

namespace {
template <int MODE, int NOUT, bool SMEM, bool TAB = false>
cudaError_t launch_fsm(dim3 grid, size_t smem, cudaStream_t st, const Params& P) {
    auto kern = fsm_kernel<MODE, NOUT, SMEM, TAB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, MVD_BLOCK, smem, st>>>(P);
    return cudaGetLastError();
}
}  // namespace

cudaError_t mvd_launch_generic_fsm(int mode, bool n2, bool in_smem, dim3 grid, size_t smem, cudaStream_t st, const Params& P) {
    if (P.enc_tab) {                                     // code given as tables (any k)
        if (mode == MODE_DETECT) return in_smem ? launch_fsm<MODE_DETECT, 0, true, true>(grid, smem, st, P) : launch_fsm<MODE_DETECT, 0, false, true>(grid, smem, st, P);
        if (mode == MODE_LEARN) return in_smem ? launch_fsm<MODE_LEARN, 0, true, true>(grid, smem, st, P) : launch_fsm<MODE_LEARN, 0, false, true>(grid, smem, st, P);
        return in_smem ? launch_fsm<MODE_TRACE, 0, true, true>(grid, smem, st, P) : launch_fsm<MODE_TRACE, 0, false, true>(grid, smem, st, P);
    }
    if (mode == MODE_DETECT) {
        if (in_smem) return n2 ? launch_fsm<MODE_DETECT, 2, true>(grid, smem, st, P) : launch_fsm<MODE_DETECT, 0, true>(grid, smem, st, P);
        return n2 ? launch_fsm<MODE_DETECT, 2, false>(grid, smem, st, P) : launch_fsm<MODE_DETECT, 0, false>(grid, smem, st, P);
    }
    if (mode == MODE_LEARN) return in_smem ? launch_fsm<MODE_LEARN, 0, true>(grid, smem, st, P) : launch_fsm<MODE_LEARN, 0, false>(grid, smem, st, P);
    return in_smem ? launch_fsm<MODE_TRACE, 0, true>(grid, smem, st, P) : launch_fsm<MODE_TRACE, 0, false>(grid, smem, st, P);
}

cudaError_t mvd_launch_generic(int engine, int mode, bool n2, int m, bool in_smem, dim3 grid, size_t smem, cudaStream_t st,
                               const Params& P) {
    if (engine == MVD_ENGINE_FSM) return mvd_launch_generic_fsm(mode, n2, in_smem, grid, smem, st, P);
    return mvd_launch_generic_acs(mode, n2, m, grid, smem, st, P);
}

cudaError_t mvd_launch_int_peak(int blocks, cudaStream_t st, uint32_t* out, int iters, int mode) {
    int_peak_kernel<<<blocks, 256, 0, st>>>(out, iters, mode);
    return cudaGetLastError();
}
