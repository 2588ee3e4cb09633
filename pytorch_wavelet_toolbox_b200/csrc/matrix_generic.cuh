// matrix_generic.cuh -- one level of the boundary-filter matrix FWT (general path).
//
// The reference multiplies by a sparse n x n operator (torch.sparse.mm,
// src/ptwt/matmul_transform.py:422 and :688).  That operator is a stride-2 filter band
// ("sameshift" rows of the convolution matrix, src/ptwt/sparse_math.py:371-377, :503-505)
// whose truncated first / last rows have been replaced by QR-orthogonalised dense rows
// (src/ptwt/sparse_math.py:253-311).  Here the band is applied as a filter and the
// replaced rows as small dense blocks; nothing n x n is ever materialised.
#pragma once

#include "common.cuh"

namespace wtb {

template <typename T>
struct MatFwdParams {
    const T* x;  // [batch, n_in]
    T* lo;       // [batch, n/2]
    T* hi;
    int64_t batch, n, n_in, x_stride, lo_stride, hi_stride;
    int L, shift, odd_mode;
    int nb_top, nb_bot, w_left, w_right;
    const T* lo_left;   // [nb_top + nb_bot, w_left]
    const T* lo_right;  // [nb_top + nb_bot, w_right]
    const T* hi_left;
    const T* hi_right;
    Taps<T> taps;  // un-flipped dec_lo / dec_hi
};

template <typename T>
__device__ __forceinline__ T mat_sample(const T* __restrict__ xb, int64_t c, int64_t n_in, int odd_mode) {
    if (c < n_in) return __ldg(xb + c);
    // the single appended sample of an odd-length level input
    // (F.pad(..., (0, 1), mode), reference matmul_transform.py:381-388, :412-421)
    switch (odd_mode) {
        case WT_MODE_ZERO: return T(0);
        case WT_MODE_REFLECT: return __ldg(xb + (n_in >= 2 ? n_in - 2 : 0));
        case WT_MODE_PERIODIC: return __ldg(xb);
        default: return __ldg(xb + n_in - 1);  // constant, symmetric
    }
}

template <typename T>
__global__ void __launch_bounds__(256) mat_fwd_kernel(const __grid_constant__ MatFwdParams<T> p) {
    const int64_t half = p.n / 2;
    const int64_t total = p.batch * half;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx % half;
        const int64_t b = idx / half;
        const T* __restrict__ xb = p.x + b * p.x_stride;
        T alo = T(0), ahi = T(0);
        if (i < p.nb_top || i >= half - p.nb_bot) {
            // orthogonalised boundary row: dense dot over the two corner windows
            const int64_t r = i < p.nb_top ? i : p.nb_top + (i - (half - p.nb_bot));
            const T* __restrict__ ll = p.lo_left + r * p.w_left;
            const T* __restrict__ hl = p.hi_left + r * p.w_left;
            for (int c = 0; c < p.w_left; ++c) {
                const T v = mat_sample(xb, (int64_t)c, p.n_in, p.odd_mode);
                alo = fma(__ldg(ll + c), v, alo);
                ahi = fma(__ldg(hl + c), v, ahi);
            }
            const T* __restrict__ lr = p.lo_right + r * p.w_right;
            const T* __restrict__ hr = p.hi_right + r * p.w_right;
            const int64_t c0 = p.n - p.w_right;
            for (int c = 0; c < p.w_right; ++c) {
                const T v = mat_sample(xb, c0 + c, p.n_in, p.odd_mode);
                alo = fma(__ldg(lr + c), v, alo);
                ahi = fma(__ldg(hr + c), v, ahi);
            }
        } else {
            const int64_t top = 2 * i + p.shift;  // column hit by tap 0
            for (int m = 0; m < p.L; ++m) {
                const int64_t c = top - m;
                if (c < 0 || c >= p.n) continue;
                const T v = mat_sample(xb, c, p.n_in, p.odd_mode);
                alo = fma(p.taps.lo[m], v, alo);
                ahi = fma(p.taps.hi[m], v, ahi);
            }
        }
        p.lo[b * p.lo_stride + i] = alo;
        p.hi[b * p.hi_stride + i] = ahi;
    }
}

template <typename T>
struct MatInvParams {
    const T* lo;  // [batch, n/2]
    const T* hi;
    T* y;         // [batch, keep]
    int64_t batch, n, keep, lo_stride, hi_stride, y_stride;
    int L, shift;
    int nb_top, nb_bot, w_left, w_right;
    const T* lo_left;
    const T* lo_right;
    const T* hi_left;
    const T* hi_right;
    Taps<T> taps;  // FLIPPED rec_lo / rec_hi (rows of S^T, reference matmul_transform.py:110-116)
};

// y = S [lo; hi] with S^T rows = stride-2 band of the flipped reconstruction filters,
// boundary rows replaced by the dense blocks.
template <typename T>
__global__ void __launch_bounds__(256) mat_inv_kernel(const __grid_constant__ MatInvParams<T> p) {
    const int64_t half = p.n / 2;
    const int64_t total = p.batch * p.keep;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = idx % p.keep;
        const int64_t b = idx / p.keep;
        const T* __restrict__ lb = p.lo + b * p.lo_stride;
        const T* __restrict__ hb = p.hi + b * p.hi_stride;
        T acc = T(0);
        // interior rows: tap m = 2 i + shift - t in [0, L)
        int64_t i0 = (t - p.shift + 1) >> 1;  // ceil((t - shift) / 2)
        int64_t i1 = (t - p.shift + p.L - 1) >> 1;
        if (i0 < p.nb_top) i0 = p.nb_top;
        if (i1 > half - p.nb_bot - 1) i1 = half - p.nb_bot - 1;
        for (int64_t i = i0; i <= i1; ++i) {
            const int m = (int)(2 * i + p.shift - t);
            acc = fma(p.taps.lo[m], __ldg(lb + i), acc);
            acc = fma(p.taps.hi[m], __ldg(hb + i), acc);
        }
        const int nb = p.nb_top + p.nb_bot;
        if (t < p.w_left) {
            for (int r = 0; r < nb; ++r) {
                const int64_t i = r < p.nb_top ? r : half - p.nb_bot + (r - p.nb_top);
                acc = fma(__ldg(p.lo_left + r * p.w_left + t), __ldg(lb + i), acc);
                acc = fma(__ldg(p.hi_left + r * p.w_left + t), __ldg(hb + i), acc);
            }
        }
        const int64_t c0 = p.n - p.w_right;
        if (t >= c0) {
            const int64_t c = t - c0;
            for (int r = 0; r < nb; ++r) {
                const int64_t i = r < p.nb_top ? r : half - p.nb_bot + (r - p.nb_top);
                acc = fma(__ldg(p.lo_right + r * p.w_right + c), __ldg(lb + i), acc);
                acc = fma(__ldg(p.hi_right + r * p.w_right + c), __ldg(hb + i), acc);
            }
        }
        p.y[b * p.y_stride + t] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// Register-blocked synthesis level of the matrix transform.
//
// Interior of  y = S [lo; hi]  (reference src/ptwt/matmul_transform.py:682-699): with even L,
//   y[t] = sum_i rec_lo[t + L/2 - 1 - 2 i] lo[i] + rec_hi[...] hi[i].
// A thread produces OPT consecutive samples starting at t0 = OPT * (global thread index): the
// coefficients it needs are the contiguous window i = t0/2 - C .. t0/2 - C + NCW - 1 of both bands
// (128-bit loads, compile-time tap indices).  Threads whose samples or window touch an
// orthogonalised boundary row / the dense corner blocks evaluate them sample by sample exactly
// like mat_inv_kernel.
// ------------------------------------------------------------------------------------------
template <typename T> struct MatInvCfg;
template <> struct MatInvCfg<float> { static constexpr int OPT = 8, VEC = 4; using V = float4; };
template <> struct MatInvCfg<double> { static constexpr int OPT = 4, VEC = 2; using V = double2; };

template <typename T>
__device__ __forceinline__ T mat_inv_sample(const MatInvParams<T>& p, const T* __restrict__ lb, const T* __restrict__ hb,
                                            int64_t t) {
    const int64_t half = p.n / 2;
    T acc = T(0);
    int64_t i0 = (t - p.shift + 1) >> 1;
    int64_t i1 = (t - p.shift + p.L - 1) >> 1;
    if (i0 < p.nb_top) i0 = p.nb_top;
    if (i1 > half - p.nb_bot - 1) i1 = half - p.nb_bot - 1;
    for (int64_t i = i0; i <= i1; ++i) {
        const int m = (int)(2 * i + p.shift - t);
        acc = fma(p.taps.lo[m], __ldg(lb + i), acc);
        acc = fma(p.taps.hi[m], __ldg(hb + i), acc);
    }
    const int nb = p.nb_top + p.nb_bot;
    if (t < p.w_left) {
        for (int r = 0; r < nb; ++r) {
            const int64_t i = r < p.nb_top ? r : half - p.nb_bot + (r - p.nb_top);
            acc = fma(__ldg(p.lo_left + r * p.w_left + t), __ldg(lb + i), acc);
            acc = fma(__ldg(p.hi_left + r * p.w_left + t), __ldg(hb + i), acc);
        }
    }
    const int64_t c0 = p.n - p.w_right;
    if (t >= c0) {
        const int64_t c = t - c0;
        for (int r = 0; r < nb; ++r) {
            const int64_t i = r < p.nb_top ? r : half - p.nb_bot + (r - p.nb_top);
            acc = fma(__ldg(p.lo_right + r * p.w_right + c), __ldg(lb + i), acc);
            acc = fma(__ldg(p.hi_right + r * p.w_right + c), __ldg(hb + i), acc);
        }
    }
    return acc;
}

template <typename T, int L>
__global__ void __launch_bounds__(256) mat_inv_fast_kernel(const __grid_constant__ MatInvParams<T> p) {
    using Cfg = MatInvCfg<T>;
    using V = typename Cfg::V;
    constexpr int OPT = Cfg::OPT, VEC = Cfg::VEC;
    constexpr int C = ((L / 4) + VEC - 1) / VEC * VEC;                 // window starts at t0/2 - C
    constexpr int NCW0 = C + (OPT + L / 2 - 2) / 2 + 1;
    constexpr int NCV = (NCW0 + VEC - 1) / VEC;
    constexpr int NCW = NCV * VEC;
    const int64_t t0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * OPT;
    if (t0 >= p.keep) return;
    const int64_t b = blockIdx.y;
    const T* __restrict__ lb = p.lo + b * p.lo_stride;
    const T* __restrict__ hb = p.hi + b * p.hi_stride;
    T* __restrict__ yb = p.y + b * p.y_stride;
    const int64_t half = p.n / 2;
    const int64_t ilo = t0 / 2 - C;
    T out[OPT];
    const bool fast = ilo >= p.nb_top && ilo + NCW <= half - p.nb_bot && t0 >= p.w_left && t0 + OPT <= p.n - p.w_right;
    if (fast) {
        T a[NCW], d[NCW];
#pragma unroll
        for (int q = 0; q < NCV; ++q) {
            const V u = __ldg(reinterpret_cast<const V*>(lb + ilo) + q);
            const V w = __ldg(reinterpret_cast<const V*>(hb + ilo) + q);
            const T* up = reinterpret_cast<const T*>(&u);
            const T* wp = reinterpret_cast<const T*>(&w);
#pragma unroll
            for (int e = 0; e < VEC; ++e) { a[VEC * q + e] = up[e]; d[VEC * q + e] = wp[e]; }
        }
#pragma unroll
        for (int g = 0; g < OPT; ++g) {
            T acc = T(0);
#pragma unroll
            for (int w = 0; w < NCW; ++w) {
                // rec index k = t + L/2 - 1 - 2 i with t = t0 + g, i = t0/2 - C + w; taps are stored flipped
                const int k = g + L / 2 - 1 + 2 * C - 2 * w;
                if (k >= 0 && k < L) {
                    acc = fma(p.taps.lo[L - 1 - k], a[w], acc);
                    acc = fma(p.taps.hi[L - 1 - k], d[w], acc);
                }
            }
            out[g] = acc;
        }
    } else {
        for (int g = 0; g < OPT; ++g) out[g] = (t0 + g < p.keep) ? mat_inv_sample(p, lb, hb, t0 + g) : T(0);
    }
    if (t0 + OPT <= p.keep) {
#pragma unroll
        for (int q = 0; q < OPT / VEC; ++q) {
            V v;
            T* vp = reinterpret_cast<T*>(&v);
#pragma unroll
            for (int e = 0; e < VEC; ++e) vp[e] = out[VEC * q + e];
            reinterpret_cast<V*>(yb + t0)[q] = v;
        }
    } else {
        for (int g = 0; g < OPT && t0 + g < p.keep; ++g) yb[t0 + g] = out[g];
    }
}

template <typename T>
static bool launch_mat_inv_fast(const MatInvParams<T>& p, cudaStream_t st, cudaError_t* err) {
    using Cfg = MatInvCfg<T>;
    constexpr int OPT = Cfg::OPT, VEC = Cfg::VEC;
    *err = cudaSuccess;
    const int L = p.L;
    if ((L & 1) || L < 2 || L > 16 || p.batch > 65535 || p.keep <= 0) return false;
    if (p.shift != L / 2) return false;
    if (((uintptr_t)p.lo & 15) || ((uintptr_t)p.hi & 15) || ((uintptr_t)p.y & 15)) return false;
    if ((p.lo_stride % VEC) || (p.hi_stride % VEC) || (p.y_stride % VEC)) return false;
    const int64_t nblk = (p.keep + (int64_t)OPT * 256 - 1) / ((int64_t)OPT * 256);
    if (nblk > 0x7fffffff) return false;
    dim3 grid((unsigned)nblk, (unsigned)p.batch);
#define WTB_MI(LL) case LL: mat_inv_fast_kernel<T, LL><<<grid, 256, 0, st>>>(p); break;
    switch (L) {
        WTB_MI(2) WTB_MI(4) WTB_MI(6) WTB_MI(8) WTB_MI(10) WTB_MI(12) WTB_MI(14) WTB_MI(16)
        default: return false;
    }
#undef WTB_MI
    *err = cudaGetLastError();
    return true;
}

// ------------------------------------------------------------------------------------------
// One level of the operator along an arbitrary axis of a [outer, n, inner] tensor (inner contiguous):
// the separable 2-D / 3-D boundary-wavelet transforms apply the 1-D operator along each axis
// (reference src/ptwt/matmul_transform_2.py:514-531, matmul_transform_3.py:255-262).  Consecutive
// threads walk the contiguous inner index, so every access is coalesced whatever the axis.
// Output layout = the reference's "A x, then split": low-pass rows 0 .. n/2-1, high-pass n/2 .. n-1
// along the transformed axis of one [outer, n, inner] buffer.
// ------------------------------------------------------------------------------------------
template <typename T>
struct MatAxisParams {
    const T* x;       // analysis: [outer, n_in, inner]; synthesis: [outer, n, inner] (lo | hi along the axis)
    T* y;             // analysis: [outer, n, inner] (lo | hi);  synthesis: [outer, keep, inner]
    int64_t outer, inner, n, n_in, keep;
    int64_t x_os, x_as, y_os, y_as;   // outer / axis strides in elements
    int L, shift, odd_mode;
    int nb_top, nb_bot, w_left, w_right;
    const T* lo_left;
    const T* lo_right;
    const T* hi_left;
    const T* hi_right;
    Taps<T> taps;
};

template <typename T>
__device__ __forceinline__ T mat_axis_sample(const T* __restrict__ xc, int64_t s, int64_t as, int64_t n_in, int odd_mode) {
    if (s < n_in) return __ldg(xc + s * as);
    switch (odd_mode) {   // the single appended sample of an odd extent (F.pad(..., (0, 1)) along this axis)
        case WT_MODE_ZERO: return T(0);
        case WT_MODE_REFLECT: return __ldg(xc + (n_in >= 2 ? n_in - 2 : 0) * as);
        case WT_MODE_PERIODIC: return __ldg(xc);
        default: return __ldg(xc + (n_in - 1) * as);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) mat_axis_fwd_kernel(const __grid_constant__ MatAxisParams<T> p) {
    const int64_t half = p.n / 2;
    const int64_t total = p.outer * half * p.inner;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx % p.inner;
        const int64_t r = idx / p.inner;
        const int64_t i = r % half, o = r / half;
        const T* __restrict__ xc = p.x + o * p.x_os + c;
        T alo = T(0), ahi = T(0);
        if (i < p.nb_top || i >= half - p.nb_bot) {
            const int64_t rr = i < p.nb_top ? i : p.nb_top + (i - (half - p.nb_bot));
            for (int q = 0; q < p.w_left; ++q) {
                const T v = mat_axis_sample(xc, (int64_t)q, p.x_as, p.n_in, p.odd_mode);
                alo = fma(__ldg(p.lo_left + rr * p.w_left + q), v, alo);
                ahi = fma(__ldg(p.hi_left + rr * p.w_left + q), v, ahi);
            }
            const int64_t c0 = p.n - p.w_right;
            for (int q = 0; q < p.w_right; ++q) {
                const T v = mat_axis_sample(xc, c0 + q, p.x_as, p.n_in, p.odd_mode);
                alo = fma(__ldg(p.lo_right + rr * p.w_right + q), v, alo);
                ahi = fma(__ldg(p.hi_right + rr * p.w_right + q), v, ahi);
            }
        } else {
            const int64_t top = 2 * i + p.shift;
            for (int m = 0; m < p.L; ++m) {
                const int64_t s = top - m;
                if (s < 0 || s >= p.n) continue;
                const T v = mat_axis_sample(xc, s, p.x_as, p.n_in, p.odd_mode);
                alo = fma(p.taps.lo[m], v, alo);
                ahi = fma(p.taps.hi[m], v, ahi);
            }
        }
        T* __restrict__ yc = p.y + o * p.y_os + c;
        yc[i * p.y_as] = alo;
        yc[(half + i) * p.y_as] = ahi;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) mat_axis_inv_kernel(const __grid_constant__ MatAxisParams<T> p) {
    const int64_t half = p.n / 2;
    const int64_t total = p.outer * p.keep * p.inner;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx % p.inner;
        const int64_t r = idx / p.inner;
        const int64_t t = r % p.keep, o = r / p.keep;
        const T* __restrict__ lc = p.x + o * p.x_os + c;          // low-pass rows
        const T* __restrict__ hc = lc + half * p.x_as;            // high-pass rows
        T acc = T(0);
        int64_t i0 = (t - p.shift + 1) >> 1;
        int64_t i1 = (t - p.shift + p.L - 1) >> 1;
        if (i0 < p.nb_top) i0 = p.nb_top;
        if (i1 > half - p.nb_bot - 1) i1 = half - p.nb_bot - 1;
        for (int64_t i = i0; i <= i1; ++i) {
            const int m = (int)(2 * i + p.shift - t);
            acc = fma(p.taps.lo[m], __ldg(lc + i * p.x_as), acc);
            acc = fma(p.taps.hi[m], __ldg(hc + i * p.x_as), acc);
        }
        const int nb = p.nb_top + p.nb_bot;
        if (t < p.w_left) {
            for (int rr = 0; rr < nb; ++rr) {
                const int64_t i = rr < p.nb_top ? rr : half - p.nb_bot + (rr - p.nb_top);
                acc = fma(__ldg(p.lo_left + rr * p.w_left + t), __ldg(lc + i * p.x_as), acc);
                acc = fma(__ldg(p.hi_left + rr * p.w_left + t), __ldg(hc + i * p.x_as), acc);
            }
        }
        const int64_t c0 = p.n - p.w_right;
        if (t >= c0) {
            const int64_t q = t - c0;
            for (int rr = 0; rr < nb; ++rr) {
                const int64_t i = rr < p.nb_top ? rr : half - p.nb_bot + (rr - p.nb_top);
                acc = fma(__ldg(p.lo_right + rr * p.w_right + q), __ldg(lc + i * p.x_as), acc);
                acc = fma(__ldg(p.hi_right + rr * p.w_right + q), __ldg(hc + i * p.x_as), acc);
            }
        }
        p.y[o * p.y_os + t * p.y_as + c] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// Register-blocked versions of the two per-axis kernels for a STRIDED axis (inner > 1): a thread owns one
// column c of the inner index and R = 4 consecutive outputs along the axis, loads the 2 R + L - 2 samples
// (analysis) or the window of both bands (synthesis) it needs once -- every load is a coalesced row
// segment across the warp -- and applies the taps with compile-time indices.  Groups that touch an
// orthogonalised boundary row, a corner block or the appended odd sample fall back to the per-output
// evaluation of the kernels above.  blockDim = (64 columns, 4 groups); grid.z strides over `outer`.
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void mat_axis_fwd_one(const MatAxisParams<T>& p, const T* __restrict__ xc, int64_t i, T& alo, T& ahi) {
    const int64_t half = p.n / 2;
    alo = T(0); ahi = T(0);
    if (i < p.nb_top || i >= half - p.nb_bot) {
        const int64_t rr = i < p.nb_top ? i : p.nb_top + (i - (half - p.nb_bot));
        for (int q = 0; q < p.w_left; ++q) {
            const T v = mat_axis_sample(xc, (int64_t)q, p.x_as, p.n_in, p.odd_mode);
            alo = fma(__ldg(p.lo_left + rr * p.w_left + q), v, alo);
            ahi = fma(__ldg(p.hi_left + rr * p.w_left + q), v, ahi);
        }
        const int64_t c0 = p.n - p.w_right;
        for (int q = 0; q < p.w_right; ++q) {
            const T v = mat_axis_sample(xc, c0 + q, p.x_as, p.n_in, p.odd_mode);
            alo = fma(__ldg(p.lo_right + rr * p.w_right + q), v, alo);
            ahi = fma(__ldg(p.hi_right + rr * p.w_right + q), v, ahi);
        }
    } else {
        const int64_t top = 2 * i + p.shift;
        for (int m = 0; m < p.L; ++m) {
            const int64_t s = top - m;
            if (s < 0 || s >= p.n) continue;
            const T v = mat_axis_sample(xc, s, p.x_as, p.n_in, p.odd_mode);
            alo = fma(p.taps.lo[m], v, alo);
            ahi = fma(p.taps.hi[m], v, ahi);
        }
    }
}

template <typename T, int L>
__global__ void __launch_bounds__(256) mat_axis_fwd_blk_kernel(const __grid_constant__ MatAxisParams<T> p) {
    constexpr int R = 4, NS = 2 * R + L - 2;
    const int64_t c = (int64_t)blockIdx.x * 64 + threadIdx.x;
    const int64_t half = p.n / 2;
    const int64_t i0 = ((int64_t)blockIdx.y * 4 + threadIdx.y) * R;
    if (c >= p.inner || i0 >= half) return;
    const int shift = L / 2;
    const int64_t s0 = 2 * i0 + shift - (L - 1);                 // first sample of the group's window
    const bool fast = i0 >= p.nb_top && i0 + R <= half - p.nb_bot && s0 >= 0 && s0 + NS <= p.n_in;
    for (int64_t o = blockIdx.z; o < p.outer; o += gridDim.z) {
        const T* __restrict__ xc = p.x + o * p.x_os + c;
        T* __restrict__ yc = p.y + o * p.y_os + c;
        if (fast) {
            T v[NS];
#pragma unroll
            for (int q = 0; q < NS; ++q) v[q] = __ldg(xc + (s0 + q) * p.x_as);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                T alo = T(0), ahi = T(0);
#pragma unroll
                for (int m = 0; m < L; ++m) {
                    alo = fma(p.taps.lo[m], v[2 * r + (L - 1) - m], alo);
                    ahi = fma(p.taps.hi[m], v[2 * r + (L - 1) - m], ahi);
                }
                yc[(i0 + r) * p.y_as] = alo;
                yc[(half + i0 + r) * p.y_as] = ahi;
            }
        } else {
            for (int r = 0; r < R && i0 + r < half; ++r) {
                T alo, ahi;
                mat_axis_fwd_one(p, xc, i0 + r, alo, ahi);
                yc[(i0 + r) * p.y_as] = alo;
                yc[(half + i0 + r) * p.y_as] = ahi;
            }
        }
    }
}

template <typename T>
__device__ __forceinline__ T mat_axis_inv_one(const MatAxisParams<T>& p, const T* __restrict__ lc, const T* __restrict__ hc, int64_t t) {
    const int64_t half = p.n / 2;
    T acc = T(0);
    int64_t i0 = (t - p.shift + 1) >> 1;
    int64_t i1 = (t - p.shift + p.L - 1) >> 1;
    if (i0 < p.nb_top) i0 = p.nb_top;
    if (i1 > half - p.nb_bot - 1) i1 = half - p.nb_bot - 1;
    for (int64_t i = i0; i <= i1; ++i) {
        const int m = (int)(2 * i + p.shift - t);
        acc = fma(p.taps.lo[m], __ldg(lc + i * p.x_as), acc);
        acc = fma(p.taps.hi[m], __ldg(hc + i * p.x_as), acc);
    }
    const int nb = p.nb_top + p.nb_bot;
    if (t < p.w_left) {
        for (int rr = 0; rr < nb; ++rr) {
            const int64_t i = rr < p.nb_top ? rr : half - p.nb_bot + (rr - p.nb_top);
            acc = fma(__ldg(p.lo_left + rr * p.w_left + t), __ldg(lc + i * p.x_as), acc);
            acc = fma(__ldg(p.hi_left + rr * p.w_left + t), __ldg(hc + i * p.x_as), acc);
        }
    }
    const int64_t c0 = p.n - p.w_right;
    if (t >= c0) {
        const int64_t q = t - c0;
        for (int rr = 0; rr < nb; ++rr) {
            const int64_t i = rr < p.nb_top ? rr : half - p.nb_bot + (rr - p.nb_top);
            acc = fma(__ldg(p.lo_right + rr * p.w_right + q), __ldg(lc + i * p.x_as), acc);
            acc = fma(__ldg(p.hi_right + rr * p.w_right + q), __ldg(hc + i * p.x_as), acc);
        }
    }
    return acc;
}

template <typename T, int L>
__global__ void __launch_bounds__(256) mat_axis_inv_blk_kernel(const __grid_constant__ MatAxisParams<T> p) {
    constexpr int R = 4, H = L / 2, C = L / 4;
    constexpr int NCW = C + (R + H - 2) / 2 + 1;                  // coefficients per band in the window of R samples
    const int64_t c = (int64_t)blockIdx.x * 64 + threadIdx.x;
    const int64_t half = p.n / 2;
    const int64_t t0 = ((int64_t)blockIdx.y * 4 + threadIdx.y) * R;
    if (c >= p.inner || t0 >= p.keep) return;
    const int64_t ilo = t0 / 2 - C;
    const bool fast = ilo >= p.nb_top && ilo + NCW <= half - p.nb_bot && t0 >= p.w_left && t0 + R <= p.n - p.w_right &&
                      t0 + R <= p.keep;
    for (int64_t o = blockIdx.z; o < p.outer; o += gridDim.z) {
        const T* __restrict__ lc = p.x + o * p.x_os + c;
        const T* __restrict__ hc = lc + half * p.x_as;
        T* __restrict__ yc = p.y + o * p.y_os + c;
        if (fast) {
            T a[NCW], d[NCW];
#pragma unroll
            for (int w = 0; w < NCW; ++w) { a[w] = __ldg(lc + (ilo + w) * p.x_as); d[w] = __ldg(hc + (ilo + w) * p.x_as); }
#pragma unroll
            for (int e = 0; e < R; ++e) {
                T acc = T(0);
#pragma unroll
                for (int w = 0; w < NCW; ++w) {
                    const int kk = e + H - 1 + 2 * C - 2 * w;     // rec index; the taps are stored flipped
                    if (kk >= 0 && kk < L) {
                        acc = fma(p.taps.lo[L - 1 - kk], a[w], acc);
                        acc = fma(p.taps.hi[L - 1 - kk], d[w], acc);
                    }
                }
                yc[(t0 + e) * p.y_as] = acc;
            }
        } else {
            for (int e = 0; e < R && t0 + e < p.keep; ++e) yc[(t0 + e) * p.y_as] = mat_axis_inv_one(p, lc, hc, t0 + e);
        }
    }
}

template <typename T>
static bool launch_mat_axis_blk(const MatAxisParams<T>& p, bool inverse, cudaStream_t st, cudaError_t* err) {
    *err = cudaSuccess;
    const int L = p.L;
    if ((L & 1) || L < 2 || L > 16 || p.shift != L / 2 || p.inner < 2) return false;
    const int64_t groups = ((inverse ? p.keep : p.n / 2) + 3) / 4;
    const int64_t gx = (p.inner + 63) / 64, gy = (groups + 3) / 4;
    if (gx > 0x7fffffff || gy > 65535) return false;
    dim3 block(64, 4), grid((unsigned)gx, (unsigned)gy, (unsigned)(p.outer < 65535 ? p.outer : 65535));
#define WTB_MAB(LL)                                                                   \
    case LL:                                                                          \
        if (inverse) mat_axis_inv_blk_kernel<T, LL><<<grid, block, 0, st>>>(p);       \
        else mat_axis_fwd_blk_kernel<T, LL><<<grid, block, 0, st>>>(p);               \
        break;
    switch (L) {
        WTB_MAB(2) WTB_MAB(4) WTB_MAB(6) WTB_MAB(8) WTB_MAB(10) WTB_MAB(12) WTB_MAB(14) WTB_MAB(16)
        default: return false;
    }
#undef WTB_MAB
    *err = cudaGetLastError();
    return true;
}

}  // namespace wtb
