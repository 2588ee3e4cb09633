// axis1d_fast.cuh -- one analysis level along a contiguous axis, register-blocked (f32 / f64).
//
// Serves the 1-D padded transform (wavedec, src/ptwt/conv_transform.py:133-141) and the interior band
// of the boundary-filter matrix transform (MatrixWavedec, src/ptwt/matmul_transform.py:409-425).  A
// thread produces OPT consecutive (lo, hi) pairs from one contiguous window of 2*OPT + L - 2 samples
// that it loads with 128-bit loads (neighbouring threads' windows overlap by L - 2 samples: L1 hits),
// and writes them with 128-bit stores.  Threads whose window touches the signal boundary (or, for
// the matrix transform, an orthogonalised boundary row) take a per-sample slow path that evaluates
// the boundary extension / the dense boundary blocks exactly like the general kernels.
#pragma once

#include "common.cuh"

namespace wtb {

template <typename T>
struct Fast1dParams {
    const T* x;
    T* lo;
    T* hi;
    int64_t xs, ls, hs;     // batch strides (elements)
    int n;                  // logical input length the level acts on
    int n_in;               // samples physically present (matrix transform with an odd level input: n - 1)
    int m;                  // outputs per signal
    int base;               // output i reads samples 2 i + base + k, k = 0 .. L-1, with taps f[k]
    int mode;               // conv: boundary mode; matrix: padding mode of the single appended sample
    int nb_top, nb_bot, w_left, w_right;
    const T* lo_left;
    const T* lo_right;
    const T* hi_left;
    const T* hi_right;
    T flo[16], fhi[16];     // taps in window order (flipped decomposition filters)
};

template <typename T> struct Fast1dCfg;
template <> struct Fast1dCfg<float> { static constexpr int OPT = 8, VEC = 4; using V = float4; };
template <> struct Fast1dCfg<double> { static constexpr int OPT = 4, VEC = 2; using V = double2; };

template <typename T, bool MATRIX>
__device__ __forceinline__ T fast1d_sample(const Fast1dParams<T>& p, const T* __restrict__ xb, int j) {
    if (MATRIX) {
        if (j < 0 || j >= p.n) return T(0);                 // truncated band row
        if (j < p.n_in) return __ldg(xb + j);
        switch (p.mode) {                                   // the one appended sample of an odd-length input
            case WT_MODE_ZERO: return T(0);
            case WT_MODE_REFLECT: return __ldg(xb + (p.n_in >= 2 ? p.n_in - 2 : 0));
            case WT_MODE_PERIODIC: return __ldg(xb);
            default: return __ldg(xb + p.n_in - 1);
        }
    } else {
        const int s = ext_index32(j, p.n, p.mode);
        return s >= 0 ? __ldg(xb + s) : T(0);
    }
}

template <typename T, int L, int OFF, bool MATRIX>
__global__ void __launch_bounds__(256) axis1d_fast_kernel(const __grid_constant__ Fast1dParams<T> p) {
    using Cfg = Fast1dCfg<T>;
    using V = typename Cfg::V;
    constexpr int OPT = Cfg::OPT, VEC = Cfg::VEC;
    constexpr int NV = 2 * OPT + L - 2;                     // samples in a thread's window
    constexpr int NVV = (OFF + NV + VEC - 1) / VEC;         // vector loads covering it
    const int i0 = (blockIdx.x * 256 + threadIdx.x) * OPT;
    if (i0 >= p.m) return;
    const int b = blockIdx.y;
    const T* __restrict__ xb = p.x + (int64_t)b * p.xs;
    T* __restrict__ lob = p.lo + (int64_t)b * p.ls;
    T* __restrict__ hib = p.hi + (int64_t)b * p.hs;
    const int j0 = 2 * i0 + p.base;                         // first sample of the window
    const int a0 = j0 - OFF;                                // 16-byte aligned start
    const int half = p.m;
    bool fast = (a0 >= 0) && (a0 + NVV * VEC <= p.n_in) && (i0 + OPT <= p.m);
    if (MATRIX) fast = fast && (i0 >= p.nb_top) && (i0 + OPT <= half - p.nb_bot);
    T lo[OPT], hi[OPT];
    if (fast) {
        T v[NVV * VEC];
#pragma unroll
        for (int q = 0; q < NVV; ++q) {
            const V t = __ldg(reinterpret_cast<const V*>(xb + a0) + q);
            if (VEC == 4) {
                const float4 f = *reinterpret_cast<const float4*>(&t);
                v[VEC * q] = f.x; v[VEC * q + 1] = f.y; v[VEC * q + (VEC > 2 ? 2 : 0)] = f.z; v[VEC * q + (VEC > 2 ? 3 : 1)] = f.w;
            } else {
                const double2 f = *reinterpret_cast<const double2*>(&t);
                v[VEC * q] = f.x; v[VEC * q + 1] = f.y;
            }
        }
#pragma unroll
        for (int g = 0; g < OPT; ++g) {
            T a = T(0), h = T(0);
#pragma unroll
            for (int k = 0; k < L; ++k) {
                a = fma(p.flo[k], v[OFF + 2 * g + k], a);
                h = fma(p.fhi[k], v[OFF + 2 * g + k], h);
            }
            lo[g] = a; hi[g] = h;
        }
    } else {
        for (int g = 0; g < OPT; ++g) {
            const int i = i0 + g;
            T a = T(0), h = T(0);
            if (i < p.m) {
                if (MATRIX && (i < p.nb_top || i >= half - p.nb_bot)) {
                    const int r = i < p.nb_top ? i : p.nb_top + (i - (half - p.nb_bot));
                    for (int c = 0; c < p.w_left; ++c) {
                        const T s = fast1d_sample<T, true>(p, xb, c);
                        a = fma(__ldg(p.lo_left + r * p.w_left + c), s, a);
                        h = fma(__ldg(p.hi_left + r * p.w_left + c), s, h);
                    }
                    for (int c = 0; c < p.w_right; ++c) {
                        const T s = fast1d_sample<T, true>(p, xb, p.n - p.w_right + c);
                        a = fma(__ldg(p.lo_right + r * p.w_right + c), s, a);
                        h = fma(__ldg(p.hi_right + r * p.w_right + c), s, h);
                    }
                } else {
                    for (int k = 0; k < L; ++k) {
                        const T s = fast1d_sample<T, MATRIX>(p, xb, 2 * i + p.base + k);
                        a = fma(p.flo[k], s, a);
                        h = fma(p.fhi[k], s, h);
                    }
                }
            }
            lo[g] = a; hi[g] = h;
        }
    }
    if (i0 + OPT <= p.m) {
#pragma unroll
        for (int q = 0; q < OPT / VEC; ++q) {
            if (VEC == 4) {
                reinterpret_cast<float4*>(lob + i0)[q] = make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
                reinterpret_cast<float4*>(hib + i0)[q] = make_float4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
            } else {
                reinterpret_cast<double2*>(lob + i0)[q] = make_double2(lo[2 * q], lo[2 * q + 1]);
                reinterpret_cast<double2*>(hib + i0)[q] = make_double2(hi[2 * q], hi[2 * q + 1]);
            }
        }
    } else {
        for (int g = 0; g < OPT && i0 + g < p.m; ++g) { lob[i0 + g] = lo[g]; hib[i0 + g] = hi[g]; }
    }
}

// Launch helper: returns false when the fast kernel does not apply (alignment, filter length).
template <typename T, bool MATRIX>
static bool launch_axis1d_fast(Fast1dParams<T>& p, int L, int64_t batch, cudaStream_t st, cudaError_t* err) {
    using Cfg = Fast1dCfg<T>;
    constexpr int VEC = Cfg::VEC, OPT = Cfg::OPT;
    *err = cudaSuccess;
    if ((L & 1) || L < 2 || L > 16 || batch > 65535 || p.m <= 0) return false;
    if (((uintptr_t)p.x & 15) || ((uintptr_t)p.lo & 15) || ((uintptr_t)p.hi & 15)) return false;
    if ((p.xs % VEC) || (p.ls % VEC) || (p.hs % VEC)) return false;
    const int off = ((p.base % VEC) + VEC) % VEC;
    dim3 grid((p.m + OPT * 256 - 1) / (OPT * 256), (unsigned)batch);
#define WTB_F1D(LL, OO)                                                       \
    if constexpr (OO < VEC) {                                                 \
        if (L == LL && off == OO) {                                           \
            axis1d_fast_kernel<T, LL, OO, MATRIX><<<grid, 256, 0, st>>>(p);   \
            *err = cudaGetLastError();                                        \
            return true;                                                      \
        }                                                                     \
    }
#define WTB_F1D_L(LL) WTB_F1D(LL, 0) WTB_F1D(LL, 1) WTB_F1D(LL, 2) WTB_F1D(LL, 3)
    WTB_F1D_L(2) WTB_F1D_L(4) WTB_F1D_L(6) WTB_F1D_L(8) WTB_F1D_L(10) WTB_F1D_L(12) WTB_F1D_L(14) WTB_F1D_L(16)
#undef WTB_F1D_L
#undef WTB_F1D
    return false;
}


// ---- synthesis along a contiguous axis (waverec, src/ptwt/conv_transform.py:184-199) -------------------
//   y[t] = sum_i lo[i] rec_lo[t + L-2 - 2i] + hi[i] rec_hi[t + L-2 - 2i]
// A thread writes OPT consecutive samples from OPT/2 + L/2 - 1 coefficients of each band.
template <typename T>
struct Fast1dInvParams {
    const T* lo;
    const T* hi;
    T* y;
    int64_t ls, hs, ys;
    int m;        // coefficients per band
    int nout;     // samples written (<= 2 m - L + 2)
    T rlo[16], rhi[16];
};

template <typename T, int L>
__global__ void __launch_bounds__(256) axis1d_inv_fast_kernel(const __grid_constant__ Fast1dInvParams<T> p) {
    using Cfg = Fast1dCfg<T>;
    using V = typename Cfg::V;
    constexpr int OPT = Cfg::OPT, VEC = Cfg::VEC, HALF = L / 2;
    constexpr int NC = OPT / 2 + HALF - 1;                   // coefficients per band in a thread's window
    constexpr int NCV = (NC + VEC - 1) / VEC;
    const int t0 = (blockIdx.x * 256 + threadIdx.x) * OPT;
    if (t0 >= p.nout) return;
    const int b = blockIdx.y;
    const T* __restrict__ lb = p.lo + (int64_t)b * p.ls;
    const T* __restrict__ hb = p.hi + (int64_t)b * p.hs;
    T* __restrict__ yb = p.y + (int64_t)b * p.ys;
    const int i0 = t0 / 2;
    T out[OPT];
    if (i0 + NCV * VEC <= p.m) {
        T a[NCV * VEC], d[NCV * VEC];
#pragma unroll
        for (int q = 0; q < NCV; ++q) {
            const V u = __ldg(reinterpret_cast<const V*>(lb + i0) + q);
            const V w = __ldg(reinterpret_cast<const V*>(hb + i0) + q);
            if (VEC == 4) {
                const float4 f = *reinterpret_cast<const float4*>(&u);
                const float4 g = *reinterpret_cast<const float4*>(&w);
                a[VEC * q] = f.x; a[VEC * q + 1] = f.y; a[VEC * q + (VEC > 2 ? 2 : 0)] = f.z; a[VEC * q + (VEC > 2 ? 3 : 1)] = f.w;
                d[VEC * q] = g.x; d[VEC * q + 1] = g.y; d[VEC * q + (VEC > 2 ? 2 : 0)] = g.z; d[VEC * q + (VEC > 2 ? 3 : 1)] = g.w;
            } else {
                const double2 f = *reinterpret_cast<const double2*>(&u);
                const double2 g = *reinterpret_cast<const double2*>(&w);
                a[VEC * q] = f.x; a[VEC * q + 1] = f.y; d[VEC * q] = g.x; d[VEC * q + 1] = g.y;
            }
        }
#pragma unroll
        for (int s = 0; s < OPT / 2; ++s) {
            T e0 = T(0), e1 = T(0);
#pragma unroll
            for (int j = 0; j < HALF; ++j) {
                e0 = fma(p.rlo[L - 2 - 2 * j], a[s + j], e0);
                e0 = fma(p.rhi[L - 2 - 2 * j], d[s + j], e0);
                e1 = fma(p.rlo[L - 1 - 2 * j], a[s + j], e1);
                e1 = fma(p.rhi[L - 1 - 2 * j], d[s + j], e1);
            }
            out[2 * s] = e0; out[2 * s + 1] = e1;
        }
    } else {
        for (int g = 0; g < OPT; ++g) {
            const int u = t0 + g + L - 2;
            T acc = T(0);
            for (int k = (u & 1); k < L; k += 2) {
                const int i = (u - k) >> 1;
                if (i >= 0 && i < p.m) {
                    acc = fma(p.rlo[k], __ldg(lb + i), acc);
                    acc = fma(p.rhi[k], __ldg(hb + i), acc);
                }
            }
            out[g] = acc;
        }
    }
    if (t0 + OPT <= p.nout) {
#pragma unroll
        for (int q = 0; q < OPT / VEC; ++q) {
            if (VEC == 4) reinterpret_cast<float4*>(yb + t0)[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
            else reinterpret_cast<double2*>(yb + t0)[q] = make_double2(out[2 * q], out[2 * q + 1]);
        }
    } else {
        for (int g = 0; g < OPT && t0 + g < p.nout; ++g) yb[t0 + g] = out[g];
    }
}

template <typename T>
static bool launch_axis1d_inv_fast(Fast1dInvParams<T>& p, int L, int64_t batch, cudaStream_t st, cudaError_t* err) {
    using Cfg = Fast1dCfg<T>;
    constexpr int VEC = Cfg::VEC, OPT = Cfg::OPT;
    *err = cudaSuccess;
    if ((L & 1) || L < 2 || L > 16 || batch > 65535 || p.nout <= 0) return false;
    if (((uintptr_t)p.lo & 15) || ((uintptr_t)p.hi & 15) || ((uintptr_t)p.y & 15)) return false;
    if ((p.ls % VEC) || (p.hs % VEC) || (p.ys % VEC)) return false;
    dim3 grid((p.nout + OPT * 256 - 1) / (OPT * 256), (unsigned)batch);
#define WTB_I1D(LL) case LL: axis1d_inv_fast_kernel<T, LL><<<grid, 256, 0, st>>>(p); break;
    switch (L) {
        WTB_I1D(2) WTB_I1D(4) WTB_I1D(6) WTB_I1D(8) WTB_I1D(10) WTB_I1D(12) WTB_I1D(14) WTB_I1D(16)
        default: return false;
    }
#undef WTB_I1D
    *err = cudaGetLastError();
    return true;
}

}  // namespace wtb
