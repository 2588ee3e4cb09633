// axis1d_fused.cuh -- several levels of the padded 1-D analysis transform (wavedec) in ONE kernel.
//
// The reference runs  F.pad -> conv1d(stride 2) -> split  once per level
// (src/ptwt/conv_transform.py:133-141), so every approximation vector cA_l makes a round trip through
// memory and a 10-level transform is 10 (x3) kernel launches.  Here a CTA takes a chunk of the signal
// through up to CONVF_MAXK levels in shared memory: only the detail coefficients and the last
// approximation of the group leave the SM -- the compulsory traffic.
//
//   * level j output i reads the extended samples 2 i - (L-2) .. 2 i + 1 of level j-1 (pad L-2 left,
//     reference _get_pad, src/ptwt/_util.py:198-228); a CTA owns TK outputs of the last fused level and,
//     going backwards, the sample ranges of the finer levels they depend on (halo L-2 on the left only);
//   * samples are kept de-interleaved (even / odd polyphase arrays): an output pair reads CONTIGUOUS runs
//     of both arrays with 64/128-bit shared loads and compile-time tap offsets (same scheme as
//     matrix_fused.cuh, whose window merely starts L/2 - 1 earlier instead of L - 2);
//   * boundary extension (zero / constant / reflect / symmetric) is evaluated in-kernel by the CTAs at the
//     two ends: the mirrored index always falls inside the staged range because every fused level is at
//     least 2 L long (host check).  "periodic" wraps to the far end of the signal, which a chunk does not
//     hold: that mode keeps the per-level kernels.
#pragma once

#include "matrix_fused.cuh"

namespace wtb {

constexpr int CONVF_MAXK = 6;

template <typename T>
struct ConvFusedParams {
    const T* x;                  // [batch, n[0]]
    int64_t x_stride;
    T* hi[CONVF_MAXK];           // detail of fused level j (1-based j -> index j-1), [batch, n[j]]
    int64_t hi_stride[CONVF_MAXK];
    T* lo;                       // approximation of the last fused level
    int64_t lo_stride;
    int k;                       // fused levels
    int n[CONVF_MAXK + 1];       // n[0] = input length, n[j] = (n[j-1] + L - 1) / 2
    int mode;
    int tk;                      // outputs of the last fused level per CTA (multiple of 4)
    int cap0;                    // capacity (samples) of the level-0 staging arrays
    T flo[16], fhi[16];          // taps in window order: out[i] = sum_k f[k] ext(a)[2i - (L-2) + k]
};

template <typename T, int L>
__global__ void __launch_bounds__(256) conv1d_fused_kernel(const __grid_constant__ ConvFusedParams<T> p) {
    using V2 = typename Vec2Of<T>::type;
    constexpr int HL = L - 2;                     // window of output i starts at sample 2 i - HL
    constexpr int DELTA = HL & 1;                 // = 0 for even L
    constexpr int PQ = ((HL + 1) / 2) & 1;        // parity of the first polyphase index (ranges start at multiples of 4)
    constexpr int NE = (L + 2 + DELTA + 1) / 2;   // polyphase entries covering the window of an output pair
    constexpr int NEV = (NE + PQ + 1) / 2 * 2;    // rounded to whole 2-element vectors

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* bufA = reinterpret_cast<T*>(smem_raw);     // even | odd arrays of the current level input
    const int capA = p.cap0 / 2 + 8;
    T* bufB = bufA + 2 * capA;

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int K = p.k;

    // ranges: [rlo[j], rhi[j]) = level-j outputs this CTA computes (j >= 1) / level-0 samples it stages (j = 0)
    int rlo[CONVF_MAXK + 1], rhi[CONVF_MAXK + 1];
    rlo[K] = blockIdx.x * p.tk;
    rhi[K] = min(rlo[K] + p.tk, p.n[K]);
    if (rlo[K] >= p.n[K]) return;
#pragma unroll
    for (int j = CONVF_MAXK; j >= 1; --j) {
        if (j > K) continue;
        const int nprev = p.n[j - 1];
        int lo = 2 * rlo[j] - HL, hi = 2 * rhi[j];
        if (lo < 0) hi = max(hi, min(L, nprev));                       // mirrored samples of the left extension
        if (hi > nprev) { lo = min(lo, max(nprev - L, 0)); hi = nprev; }   // ... and of the right one
        lo = max(lo, 0) & ~3;
        rlo[j - 1] = lo;
        rhi[j - 1] = hi;
    }

    // stage the level-0 samples, de-interleaved
    {
        const T* __restrict__ xb = p.x + (int64_t)b * p.x_stride;
        const int s0 = rlo[0], cnt = rhi[0] - rlo[0];
        T* ev = bufA;
        T* od = bufA + capA;
        for (int q = tid; 2 * q < cnt; q += 256) {
            const int s = s0 + 2 * q;
            if (s + 1 < rhi[0]) {
                const V2 v = __ldg(reinterpret_cast<const V2*>(xb + s));   // s is even and the row start is aligned
                ev[q] = v.x; od[q] = v.y;
            } else {
                ev[q] = __ldg(xb + s); od[q] = T(0);
            }
        }
    }
    __syncthreads();

    T* cur = bufA;
    int cur_cap = capA;
    T* nxt = bufB;
#pragma unroll 1
    for (int j = 1; j <= K; ++j) {
        const int m = p.n[j];                     // outputs of level j
        const int nprev = p.n[j - 1];
        const int in0 = rlo[j - 1];               // sample index of polyphase entry 0
        const int in1 = rhi[j - 1];
        const int nxt_cap = ((rhi[j] - rlo[j]) / 2 + 9) & ~1;
        const T* ev = cur;
        const T* od = cur + cur_cap;
        T* nev = nxt;
        T* nod = nxt + nxt_cap;
        const int own0 = (blockIdx.x * p.tk) << (K - j);
        const int own1 = min(((blockIdx.x + 1) * p.tk) << (K - j), m);
        T* __restrict__ hib = p.hi[j - 1] + (int64_t)b * p.hi_stride[j - 1];
        T* __restrict__ lob = p.lo + (int64_t)b * p.lo_stride;
        const int npairs = (rhi[j] - rlo[j] + 1) / 2;
        for (int pr = tid; pr < npairs; pr += 256) {
            const int i = rlo[j] + 2 * pr;        // even output index
            T alo[2] = {T(0), T(0)}, ahi[2] = {T(0), T(0)};
            const int q0 = ((2 * i - HL - in0) >> 1) - PQ;
            const bool interior = (2 * i - HL >= in0) && (2 * i + 3 < in1) && (2 * i + 3 < nprev) && (q0 >= 0) && (i + 1 < m);
            if (interior) {
                T e[NEV], o[NEV];
#pragma unroll
                for (int v = 0; v < NEV / 2; ++v) {
                    const V2 a = *reinterpret_cast<const V2*>(ev + q0 + 2 * v);
                    const V2 c = *reinterpret_cast<const V2*>(od + q0 + 2 * v);
                    e[2 * v] = a.x; e[2 * v + 1] = a.y; o[2 * v] = c.x; o[2 * v + 1] = c.y;
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
#pragma unroll
                    for (int k = 0; k < L; ++k) {
                        const int d = 2 * r + k + DELTA + 2 * PQ;    // offset from sample in0 + 2 q0
                        const T s = (d & 1) ? o[d >> 1] : e[d >> 1];
                        alo[r] = fma(p.flo[k], s, alo[r]);
                        ahi[r] = fma(p.fhi[k], s, ahi[r]);
                    }
                }
            } else {
                for (int r = 0; r < 2; ++r) {
                    const int ii = i + r;
                    if (ii >= m) continue;
                    for (int k = 0; k < L; ++k) {
                        const int s = ext_index32(2 * ii - HL + k, nprev, p.mode);   // < 0: zero extension
                        if (s < 0) continue;
                        const int d = s - in0;
                        const T v = (d & 1) ? od[d >> 1] : ev[d >> 1];
                        alo[r] = fma(p.flo[k], v, alo[r]);
                        ahi[r] = fma(p.fhi[k], v, ahi[r]);
                    }
                }
            }
            // approximation -> next level (de-interleaved), detail -> HBM (owned range only)
            const int rel = (i - rlo[j]) >> 1;
            if (j < K) { nev[rel] = alo[0]; nod[rel] = alo[1]; }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int ii = i + r;
                if (ii >= own0 && ii < own1) {
                    hib[ii] = ahi[r];
                    if (j == K) lob[ii] = alo[r];
                }
            }
        }
        __syncthreads();
        T* t = const_cast<T*>(cur); cur = nxt; nxt = t;
        cur_cap = nxt_cap;
    }
}

// Host: one fused group of k levels.  n[0..k] are the lengths (n[0] = group input).  Returns false when the
// group is not eligible (the caller then uses the per-level kernels).
template <typename T>
static bool launch_conv1d_fused(int L, int k, const int64_t* n, int mode, const T* x, int64_t xs, int64_t batch,
                                void* const* hi_out, const int64_t* hi_stride, T* lo_out, int64_t lo_stride, const T* flo,
                                const T* fhi, cudaStream_t st, cudaError_t* err) {
    *err = cudaSuccess;
    if ((L & 1) || L < 2 || L > 16 || k < 2 || k > CONVF_MAXK || batch > 65535 || batch < 1) return false;
    if (mode == WT_MODE_PERIODIC) return false;
    if (((uintptr_t)x & 15) || (xs & 3) || n[0] >= (int64_t(1) << 30)) return false;
    ConvFusedParams<T> p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.x_stride = xs; p.k = k; p.mode = mode;
    for (int j = 0; j <= k; ++j) p.n[j] = (int)n[j];
    for (int j = 0; j < k; ++j) {
        if (n[j] < 2 * L) return false;                               // mirrored indices stay inside the staged range
        if (n[j + 1] != (n[j] + L - 1) / 2) return false;
        p.hi[j] = (T*)hi_out[j]; p.hi_stride[j] = hi_stride[j];
    }
    p.lo = lo_out; p.lo_stride = lo_stride;
    for (int q = 0; q < L; ++q) { p.flo[q] = flo[q]; p.fhi[q] = fhi[q]; }
    const int nk = p.n[k];
    int chunk0 = sizeof(T) == 8 ? 8192 : 16384;                      // level-0 samples per CTA
    if (knob_is_set(K_CONVF_CHUNK)) { const int v = (int)knob_val(K_CONVF_CHUNK, 0); if (v >= 64 && v <= 32768) chunk0 = v; }
    int tk = chunk0 >> k;
    if (tk < 4) tk = 4;
    tk = (tk + 3) & ~3;
    if (tk > nk) tk = (nk + 3) & ~3;
    p.tk = tk;
    int cap0 = (tk << k) + ((L + 6) << k) + 64;
    if (cap0 > p.n[0] + 16) cap0 = (p.n[0] + 16 + 3) & ~3;
    cap0 = (cap0 + 3) & ~3;
    p.cap0 = cap0;
    const size_t smem = (size_t)(2 * (cap0 / 2 + 8) + 2 * (cap0 / 4 + 16 + L)) * sizeof(T);
    if (smem > 200 * 1024) return false;
    dim3 grid((nk + tk - 1) / tk, (unsigned)batch);
#define WTB_CF(LL)                                                                                              \
    case LL: {                                                                                                  \
        cudaError_t e = ensure_dyn_smem(conv1d_fused_kernel<T, LL>, smem > 200 * 1024 ? smem : 200 * 1024);          \
        if (e != cudaSuccess) { *err = e; return true; }                                                        \
        conv1d_fused_kernel<T, LL><<<grid, 256, smem, st>>>(p);                                                 \
        break;                                                                                                  \
    }
    switch (L) {
        WTB_CF(2) WTB_CF(4) WTB_CF(6) WTB_CF(8) WTB_CF(10) WTB_CF(12) WTB_CF(14) WTB_CF(16)
        default: return false;
    }
#undef WTB_CF
    *err = cudaGetLastError();
    return true;
}

}  // namespace wtb
