// generic_axis.cuh -- shape-agnostic single-axis analysis / synthesis kernels.
//
// These are the general path of the filter bank: any filter length (<= WT_MAX_FILT_LEN),
// any extents, any strides, f32 / f64, all five boundary modes.  One thread produces one
// (lo, hi) coefficient pair (analysis) or one reconstructed sample (synthesis); loads go
// through L1, which absorbs the L-fold overlap between neighbouring outputs.  The tiled
// shared-memory kernels in fused2d.cuh etc. replace them on the shapes that matter; this
// file stays the fallback and the in-library cross-check.
//
// Tensor view:  [o1, o2, n, inner]  with inner contiguous (stride 1); the transformed
// axis is `n`.  inner == 1 is the "last axis" case.
#pragma once

#include "common.cuh"

namespace wtb {

template <typename T>
struct AxisFwdParams {
    const T* x;
    T* lo;
    T* hi;
    int64_t n, m, inner, o1, o2;
    int64_t xs_o1, xs_o2, xs_n;
    int64_t ls_o1, ls_o2, ls_m;
    int64_t hs_o1, hs_o2, hs_m;
    int mode, L, padl;
    Taps<T> taps;  // un-flipped dec_lo / dec_hi
};

// c_k[i] = sum_{k<L} dec[L-1-k] * ext(x)[2 i + k - padl]
// (reference: F.pad + conv1d(stride=2) with the flipped filter,
//  src/ptwt/conv_transform.py:136-137, src/ptwt/_util.py:222-228).
template <typename T>
__global__ void __launch_bounds__(256) axis_fwd_kernel(const __grid_constant__ AxisFwdParams<T> p) {
    const int64_t total = p.o1 * p.o2 * p.m * p.inner;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = idx;
        const int64_t j = r % p.inner;
        r /= p.inner;
        const int64_t i = r % p.m;
        r /= p.m;
        const int64_t b2 = r % p.o2;
        const int64_t b1 = r / p.o2;
        const T* __restrict__ xb = p.x + b1 * p.xs_o1 + b2 * p.xs_o2 + j;
        const int64_t base = 2 * i - p.padl;
        T alo = T(0), ahi = T(0);
        if (base >= 0 && base + p.L <= p.n) {
            for (int k = 0; k < p.L; ++k) {
                const T v = __ldg(xb + (base + k) * p.xs_n);
                alo = fma(p.taps.lo[p.L - 1 - k], v, alo);
                ahi = fma(p.taps.hi[p.L - 1 - k], v, ahi);
            }
        } else {
            for (int k = 0; k < p.L; ++k) {
                const int64_t s = ext_index(base + k, p.n, p.mode);
                const T v = s >= 0 ? __ldg(xb + s * p.xs_n) : T(0);
                alo = fma(p.taps.lo[p.L - 1 - k], v, alo);
                ahi = fma(p.taps.hi[p.L - 1 - k], v, ahi);
            }
        }
        p.lo[b1 * p.ls_o1 + b2 * p.ls_o2 + i * p.ls_m + j] = alo;
        p.hi[b1 * p.hs_o1 + b2 * p.hs_o2 + i * p.hs_m + j] = ahi;
    }
}

template <typename T>
struct AxisInvParams {
    const T* lo;
    const T* hi;
    T* y;
    int64_t m;      // coefficient extent along the axis
    int64_t nout;   // samples written along the axis (<= 2(m-1) + L - 2 padl)
    int64_t inner, o1, o2;
    int64_t ls_o1, ls_o2, ls_m;
    int64_t hs_o1, hs_o2, hs_m;
    int64_t ys_o1, ys_o2, ys_n;
    int L, padl;
    Taps<T> taps;  // un-flipped rec_lo / rec_hi
};

// y[t] = sum_i lo[i] rec_lo[t + padl - 2 i] + hi[i] rec_hi[t + padl - 2 i]
// (reference: conv_transpose1d(stride=2) then crop padl on the left,
//  src/ptwt/conv_transform.py:186-199).
template <typename T>
__global__ void __launch_bounds__(256) axis_inv_kernel(const __grid_constant__ AxisInvParams<T> p) {
    const int64_t total = p.o1 * p.o2 * p.nout * p.inner;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = idx;
        const int64_t j = r % p.inner;
        r /= p.inner;
        const int64_t t = r % p.nout;
        r /= p.nout;
        const int64_t b2 = r % p.o2;
        const int64_t b1 = r / p.o2;
        const T* __restrict__ lb = p.lo + b1 * p.ls_o1 + b2 * p.ls_o2 + j;
        const T* __restrict__ hb = p.hi + b1 * p.hs_o1 + b2 * p.hs_o2 + j;
        const int64_t u = t + p.padl;
        int64_t i0 = (u - p.L + 2) >> 1;  // ceil((u - L + 1) / 2), arithmetic shift floors
        if (i0 < 0) i0 = 0;
        int64_t i1 = u >> 1;
        if (i1 > p.m - 1) i1 = p.m - 1;
        T acc = T(0);
        for (int64_t i = i0; i <= i1; ++i) {
            const int k = (int)(u - 2 * i);
            acc = fma(p.taps.lo[k], __ldg(lb + i * p.ls_m), acc);
            acc = fma(p.taps.hi[k], __ldg(hb + i * p.hs_m), acc);
        }
        p.y[b1 * p.ys_o1 + b2 * p.ys_o2 + t * p.ys_n + j] = acc;
    }
}

}  // namespace wtb
