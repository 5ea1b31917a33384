// translation unit: Eq. 4-5 at m = 4..6, two trials per thread, final metric vectors only (mvd_acsp.cuh)
#include "mvd_acsp.cuh"
#include "mvd_launch.h"

cudaError_t mvd_launch_acsp(int m, dim3 grid, unsigned threads, cudaStream_t st, const Params& P, const DevSeg& sg,
                            const uint32_t* sel, uint8_t* final_met) {
    AcspSel S;
    for (int i = 0; i < 128; ++i) S.sel[i] = sel[i];
    switch (m) {
        case 2: acsp_kernel<2><<<grid, threads, 0, st>>>(P, sg, S, final_met); break;
        case 3: acsp_kernel<3><<<grid, threads, 0, st>>>(P, sg, S, final_met); break;
        case 4: acsp_kernel<4><<<grid, threads, 0, st>>>(P, sg, S, final_met); break;
        case 5: acsp_kernel<5><<<grid, threads, 0, st>>>(P, sg, S, final_met); break;
        case 6: acsp_kernel<6><<<grid, threads, 0, st>>>(P, sg, S, final_met); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
