// Latency of a dependent DADD / FADD / IADD chain on one warp (cycles per instruction), and with 2 and 4
// independent chains in the same thread.  nvcc -arch=sm_100a -O3 -o dadd_latency dadd_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dadd_chain(double* out, const double* in, int n, long long* cyc, int chains) {
    double a = in[0], b = in[1], c = in[2], d = in[3];
    const double x = in[4];
    long long t0 = clock64();
    if (chains == 1) {
#pragma unroll 16
        for (int i = 0; i < n; ++i) a += x;
    } else if (chains == 2) {
#pragma unroll 16
        for (int i = 0; i < n; ++i) { a += x; b += x; }
    } else {
#pragma unroll 16
        for (int i = 0; i < n; ++i) { a += x; b += x; c += x; d += x; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + b + c + d;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void fadd_chain(float* out, const float* in, int n, long long* cyc) {
    float a = in[0];
    const float x = in[4];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a += x;
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *din, *dout; float *fin, *fout; long long* cyc;
    cudaMalloc(&din, 64); cudaMalloc(&dout, 4096); cudaMalloc(&fin, 64); cudaMalloc(&fout, 4096); cudaMalloc(&cyc, 8);
    double h[5] = {1.0, 2.0, 3.0, 4.0, 1e-9}; float hf[5] = {1.f, 2.f, 3.f, 4.f, 1e-9f};
    cudaMemcpy(din, h, 40, cudaMemcpyHostToDevice); cudaMemcpy(fin, hf, 20, cudaMemcpyHostToDevice);
    const int n = 1 << 16;
    for (int warps = 1; warps <= 8; warps *= 2)
        for (int chains = 1; chains <= 4; chains *= 2) {
            long long c = 0;
            dadd_chain<<<1, 32 * warps>>>(dout, din, n, cyc, chains);
            dadd_chain<<<1, 32 * warps>>>(dout, din, n, cyc, chains);
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("DADD warps/SM=%d chains=%d: %.2f cycles per loop iteration (%.2f per DADD)\n", warps, chains, (double)c / n, (double)c / n / chains);
        }
    long long c = 0;
    fadd_chain<<<1, 32>>>(fout, fin, n, cyc);
    fadd_chain<<<1, 32>>>(fout, fin, n, cyc);
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("FADD 1 warp 1 chain: %.2f cycles per FADD\n", (double)c / n);
    return 0;
}
