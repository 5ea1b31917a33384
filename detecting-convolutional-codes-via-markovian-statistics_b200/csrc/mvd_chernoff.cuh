// mvd_chernoff.cuh -- spectral radius of the Chernoff matrix M(u) of Eq. 7,
// compute_error_exponent of alpha_exponent.py:155-184:
//
//     M(u)[i, j] = sum_r P1(i->j, r)^u * P2(i->j, r)^(1-u),      I_err = min_u -log rho(M(u)).
//
// Both joint tensors come from Laplace-smoothed counts of the same decoder (alpha_exponent.py:
// 83-149), so entry (i, j, r) is either an *edge* (j = NEXT[i][r], probability (c + lambda) / d_i) or
// background (lambda / d_i, the same for every j and r of row i).  M(u) is therefore
//
//     M(u) = R * bg(u) 1^T  +  sparse(i, NEXT[i][r]) (w(u)[i, r] - bg(u)[i]),
//
// with bg(u)[i] = exp(u lb1[i] + (1-u) lb2[i]) and w(u)[i, r] = exp(u lp1[i, r] + (1-u) lp2[i, r]): a
// matrix-vector product costs O(K R) instead of O(K^2 R), which is what lets m = 4 (K = 2.3e5)
// run at all -- the reference's dense K x K x R tensor would need 1.7 TB there.  M(u) is entrywise
// positive, so the Perron root is the limit of sum(M x) / sum(x) under power iteration.
//
// One thread block per grid point u; vectors live in global memory (L2-resident), all reductions
// are fixed-order block reductions, so results are deterministic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mvd_launch.h"

__device__ __forceinline__ double chernoff_block_sum(double v, double* red) {
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
    __syncthreads();                         // red may still be read by the previous call
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double t = 0.0;
    if (wid == 0) {
        t = lane < (CHERNOFF_BLOCK >> 5) ? red[lane] : 0.0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_down_sync(0xFFFFFFFFu, t, d);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// RECOMPUTE = false: w(u) - bg(u) and R bg(u) are built once per u into global scratch and re-read by every product
// (small K: the scratch stays in L2).  RECOMPUTE = true: they are recomputed from the log tables in every product --
// at K > ~10^5 the 401 x K x R x 8-byte scratch (3 GB at K = 232 567) is streamed from DRAM once per product while the
// log tables (64 B per row, shared by all u) stay in L2, and a handful of products suffice there.
template <bool RECOMPUTE>
__global__ void __launch_bounds__(CHERNOFF_BLOCK) chernoff_rho_kernel(const __grid_constant__ ChernoffParams P) {
    __shared__ double red[33];
    const uint32_t ui = blockIdx.x;
    if (ui >= P.nu) return;
    const uint32_t K = P.K, R = P.R;
    const double u = P.u_vals[ui], v = 1.0 - u;
    double* wd = RECOMPUTE ? nullptr : P.wd + (size_t)ui * K * R;
    double* bgR = RECOMPUTE ? nullptr : P.bgR + (size_t)ui * K;
    double* x = P.xa + (size_t)ui * K;
    double* y = P.xb + (size_t)ui * K;
    for (uint32_t i = threadIdx.x; i < K; i += CHERNOFF_BLOCK) {
        if (!RECOMPUTE) {
            const double bg = exp(u * P.lb1[i] + v * P.lb2[i]);
            bgR[i] = bg * (double)R;
            for (uint32_t r = 0; r < R; ++r) {
                const size_t e = (size_t)i * R + r;
                wd[e] = exp(u * P.lp1[e] + v * P.lp2[e]) - bg;
            }
        }
        x[i] = 1.0 / (double)K;
    }
    __syncthreads();
    double rho = 0.0;
    uint32_t it = 0;
    for (; it < P.max_iter; ++it) {
        // sum(x) = 1 by construction
        double part = 0.0;
        for (uint32_t i = threadIdx.x; i < K; i += CHERNOFF_BLOCK) {
            double acc, bg = 0.0;
            if (RECOMPUTE) {
                bg = exp(u * __ldg(P.lb1 + i) + v * __ldg(P.lb2 + i));
                acc = bg * (double)R;
            } else {
                acc = bgR[i];
            }
            for (uint32_t r = 0; r < R; ++r) {
                const size_t e = (size_t)i * R + r;
                const double w = RECOMPUTE ? exp(u * __ldg(P.lp1 + e) + v * __ldg(P.lp2 + e)) - bg : wd[e];
                acc = fma(w, x[P.nxt[e]], acc);
            }
            y[i] = acc;
            part += acc;
        }
        const double s = chernoff_block_sum(part, red);           // = sum(M x) / sum(x)
        const double inv = 1.0 / s;
        for (uint32_t i = threadIdx.x; i < K; i += CHERNOFF_BLOCK) y[i] *= inv;
        __syncthreads();
        double* t = x;
        x = y;
        y = t;
        const bool done = fabs(s - rho) <= P.tol * s;
        rho = s;
        if (done && it > 0) {
            ++it;
            break;
        }
    }
    if (threadIdx.x == 0) {
        P.rho[ui] = rho;
        P.iters[ui] = it;
    }
}

// Dense form, for callers that hold the full K x K x R tensors the reference's API returns
// (alpha_exponent.py:155-184 takes P1_ijr, P2_ijr): M(u) is built once per u into scratch (stored
// transposed so that the row-per-thread product reads coalesced), then the same power iteration.
__global__ void __launch_bounds__(CHERNOFF_BLOCK) chernoff_dense_kernel(const __grid_constant__ ChernoffParams P) {
    __shared__ double red[33];
    const uint32_t ui = blockIdx.x;
    if (ui >= P.nu) return;
    const uint32_t K = P.K, R = P.R;
    const double u = P.u_vals[ui], v = 1.0 - u;
    double* Mt = P.wd + (size_t)ui * K * K;                 // Mt[j * K + i] = M[i][j]
    double* x = P.xa + (size_t)ui * K;
    double* y = P.xb + (size_t)ui * K;
    for (size_t q = threadIdx.x; q < (size_t)K * K; q += CHERNOFF_BLOCK) {
        const size_t i = q / K, j = q % K;
        double acc = 0.0;
        for (uint32_t r = 0; r < R; ++r) acc += exp(u * P.lp1[q * R + r] + v * P.lp2[q * R + r]);   // r ascending, like np.sum(axis=2)
        Mt[j * K + i] = acc;
    }
    for (uint32_t i = threadIdx.x; i < K; i += CHERNOFF_BLOCK) x[i] = 1.0 / (double)K;
    __syncthreads();
    double rho = 0.0;
    uint32_t it = 0;
    for (; it < P.max_iter; ++it) {
        double part = 0.0;
        for (uint32_t i = threadIdx.x; i < K; i += CHERNOFF_BLOCK) {
            double acc = 0.0;
            for (uint32_t j = 0; j < K; ++j) acc = fma(Mt[(size_t)j * K + i], x[j], acc);
            y[i] = acc;
            part += acc;
        }
        const double s = chernoff_block_sum(part, red);
        const double inv = 1.0 / s;
        for (uint32_t i = threadIdx.x; i < K; i += CHERNOFF_BLOCK) y[i] *= inv;
        __syncthreads();
        double* t = x;
        x = y;
        y = t;
        const bool done = fabs(s - rho) <= P.tol * s;
        rho = s;
        if (done && it > 0) {
            ++it;
            break;
        }
    }
    if (threadIdx.x == 0) {
        P.rho[ui] = rho;
        P.iters[ui] = it;
    }
}
