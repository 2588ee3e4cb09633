// tap_grad.cuh -- gradient of one transform level with respect to the filter taps (learnable wavelets).
//
// The reference differentiates through its filters because they are ordinary conv weights
// (src/ptwt/wavelets_learnable.py:167-189 makes them nn.Parameters; src/ptwt/_util.py:129-141 builds the
// conv kernels from them).  For one axis of one level, analysis
//     c_k[i] = sum_m dec_k[m] * x0[2 i + 1 - m]            (x0 = zero extension of the explicitly extended input)
// and synthesis
//     y[n]   = sum_i c_k[i] * rec_k[n + (L - 2) - 2 i]      (n indexes the cropped output)
// give the SAME correlation
//     out_k[t] = sum_{b, i} c_k[b, i] * s[b, 2 i + t + 2 - L],   t = 0 .. L-1,
// with (c, s) = (upstream gradient, input) for analysis (d dec_k[m] = out_k[L - 1 - m]) and (coefficients, upstream
// gradient) for synthesis (d rec_k[t] = out_k[t]).  Accumulation in float64, one atomicAdd per warp and tap.
#pragma once

#include "common.cuh"

namespace wtb {

template <typename T>
__global__ void __launch_bounds__(256) tap_corr_kernel(const T* __restrict__ clo, const T* __restrict__ chi, int64_t cs,
                                                         const T* __restrict__ sig, int64_t ss, int m, int n, int L,
                                                         double* __restrict__ out) {
    const int64_t b = blockIdx.y;
    const T* lo = clo + b * cs;
    const T* hi = chi + b * cs;
    const T* s = sig + b * ss;
    const int off = 2 - L;
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, step = gridDim.x * blockDim.x;
    for (int t = 0; t < L; ++t) {
        double al = 0.0, ah = 0.0;
        for (int i = i0; i < m; i += step) {
            const int idx = 2 * i + t + off;
            if (idx >= 0 && idx < n) {
                const double sv = (double)s[idx];
                al = fma((double)lo[i], sv, al);
                ah = fma((double)hi[i], sv, ah);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            al += __shfl_down_sync(0xffffffffu, al, d);
            ah += __shfl_down_sync(0xffffffffu, ah, d);
        }
        if ((threadIdx.x & 31) == 0) {
            if (al != 0.0) atomicAdd(out + t, al);
            if (ah != 0.0) atomicAdd(out + L + t, ah);
        }
    }
}

template <typename T>
static cudaError_t launch_tap_corr(const T* clo, const T* chi, int64_t cs, const T* sig, int64_t ss, int64_t rows, int m,
                                   int n, int L, double* out, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double) * 2 * (size_t)L, st);
    if (e != cudaSuccess || rows == 0 || m == 0) return e;
    int gx = (m + 255) / 256;
    if (gx > 64) gx = 64;
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
        const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
        tap_corr_kernel<T><<<dim3(gx, (unsigned)nr), 256, 0, st>>>(clo + r0 * cs, chi + r0 * cs, cs, sig + r0 * ss, ss, m, n, L, out);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace wtb
