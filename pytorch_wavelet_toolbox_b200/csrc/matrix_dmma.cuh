// matrix_dmma.cuh -- the boundary-filter matrix FWT (float64) on the FP64 TENSOR CORES.
//
// north_star: "tensor cores are used only for the MatrixWavedec path where the boundary-filter sparse matmul is
// reformulated as a banded dense contraction".  The level operator A_n of the reference
// (torch.sparse.mm(A_level, .), src/ptwt/matmul_transform.py:409-425) is block-Toeplitz away from its corner blocks:
// every output pair (lo[i], hi[i]) is the same L-tap window sliding by two samples.  Four consecutive outputs of both
// bands (8 rows) read one window of L + 6 samples, so a tile of 8 such groups is the dense product
//
//     D[8 x 8] = A[8 x (L+6)] * B[(L+6) x 8],   A[(band, s)][u] = f_band[u - 2 s],   B[u][g] = x[2 i0 + 8 g - HL + u]
//
// evaluated with mma.sync.aligned.m8n8k4.f64 (DMMA; tcgen05 has no f64 kind): about (L+6)/4 instructions of 256 FMAs
// for 64 outputs.  60 % of those FMAs multiply structural zeros of A -- the price of the dense form -- but the kernel
// is bound by instruction issue, not by the FP64 pipe (profiles/r02_matfwd_stream_ncu_summary.txt: DFMA 25 % of 214 M
// warp instructions, fp64 pipe 22 % busy), and the DMMA form needs ~0.3 warp instructions per output instead of 1.5.
//
// Everything around the contraction is the streaming cascade of matrix_fused.cuh: a CTA takes consecutive chunks of one
// row through K levels, level inputs live in shared memory (interleaved, as the B fragment wants them), the next chunk
// arrives by cp.async meanwhile, details go to HBM, the corner blocks (dense orthogonalised boundary rows) are applied
// by scalar code in the CTAs at the two ends.
#pragma once

#include "matrix_fused.cuh"

namespace wtb {

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int L, int NT, bool PERM>
__global__ void __launch_bounds__(NT) mat_fwd_dmma_kernel(const __grid_constant__ MatFusedParams<double> p) {
    constexpr int HL = L / 2 - 1, HR = L / 2;
    // The contraction index u (window sample of a group of 4 outputs) is split as u = E * k + e: lane k of a fragment
    // column holds E CONSECUTIVE samples (E even, window start shifted by SH to an even sample), so the B fragments of
    // one tile are E / 2 aligned 128-bit shared loads per lane (conflict-free for E = 6) instead of one 64-bit load per
    // k-step with 4-way bank conflicts; the A fragment is permuted the same way.
    // PERM = false: u = 4 e + k (E = ceil((L + 6) / 4) k-steps, one 64-bit load each, 4-way conflicts).
    constexpr int SH = PERM ? (HL & 1) : 0;
    constexpr int E = PERM ? ((L + 6 + SH + 3) / 4 + 1) / 2 * 2 : (L + 6 + 3) / 4;
    constexpr int KS = PERM ? E : 1;              // stride of the lane index k in the window
    constexpr int ES = PERM ? 1 : 4;              // stride of the k-step e in the window
    constexpr int NW = NT / 32;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* raw0 = reinterpret_cast<double*>(smem_raw);   // level-0 samples of the current / next chunk (two buffers)
    double* raw1 = raw0 + p.cap0;
    double* levA = raw1 + p.cap0;                          // approximations, alternating
    double* levB = levA + (p.cap0 / 2 + 16);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int K = p.k;
    const double* __restrict__ xb = p.x + (int64_t)b * p.x_stride;

    __shared__ int s_r[2][2][MATF_MAXK + 1];
    const int nchunks = (p.n[K] + p.tk - 1) / p.tk;
    const int c_first = blockIdx.x * p.cpc, c_last = min(c_first + p.cpc, nchunks);
    if (c_first >= nchunks) return;
    auto ranges = [&](int chunk, int set) {
        int lo_j = chunk * p.tk, hi_j = min(lo_j + p.tk, p.n[K]);
        s_r[set][0][K] = lo_j; s_r[set][1][K] = hi_j;
        for (int j = K; j >= 1; --j) {
            const int half = p.n[j];
            int lo = 2 * lo_j - HL, hi = 2 * (hi_j - 1) + HR + 1;
            if (lo_j < p.nb_top[j - 1]) lo = 0, hi = max(hi, p.w_left[j - 1]);
            if (hi_j > half - p.nb_bot[j - 1]) hi = p.n[j - 1], lo = min(lo, p.n[j - 1] - p.w_right[j - 1]);
            lo = max(lo, 0) & ~3;
            hi = min(hi, p.n[j - 1]);
            s_r[set][0][j - 1] = lo_j = lo;
            s_r[set][1][j - 1] = hi_j = hi;
        }
    };
    auto prefetch = [&](double* dst, int s0, int s1) {
        const int cnt = s1 - s0, nv = cnt / 2;
        for (int q = tid; q < nv; q += NT) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(dst + 2 * q);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(xb + s0 + 2 * q) : "memory");
        }
        if ((cnt & 1) && tid == 0) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(dst + cnt - 1);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(xb + s1 - 1) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // A fragment (row-major 8 x 4 per k-step e): this lane holds A[m = lane / 4][u = E * (lane % 4) + e],
    // row m = 4 * band + s  ->  f_band[u - SH - 2 s]
    double afrag[E];
    {
        const int m = lane >> 2, band = m >> 2, s = m & 3;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int kk = KS * (lane & 3) + ES * e - SH - 2 * s;
            afrag[e] = (kk >= 0 && kk < L) ? (band ? p.fhi[kk] : p.flo[kk]) : 0.0;
        }
    }
    const int frag_n = lane >> 2, frag_k = lane & 3;      // B fragment: B[k = lane % 4][n = lane / 4]
    const int out_band = lane >> 4, out_s = (lane >> 2) & 3;
    const int out_g0 = 2 * (lane & 3);                     // D fragment: columns (groups) 2 (lane % 4), + 1

    if (tid == 0) ranges(c_first, 0);
    __syncthreads();
    prefetch(raw0, s_r[0][0][0], s_r[0][1][0]);

    int set = 0;
    for (int chunk = c_first; chunk < c_last; ++chunk, set ^= 1) {
        const int* rlo = s_r[set][0];
        const int* rhi = s_r[set][1];
        double* in = set ? raw1 : raw0;
        if (tid == 0 && chunk + 1 < c_last) ranges(chunk + 1, set ^ 1);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();                              // this chunk's samples landed; the next chunk's ranges are visible
        if (chunk + 1 < c_last) prefetch(set ? raw0 : raw1, s_r[set ^ 1][0][0], s_r[set ^ 1][1][0]);

        double* nxt = levA;
#pragma unroll 1
        for (int j = 1; j <= K; ++j) {
            const int half = p.n[j], nprev = p.n[j - 1];
            const int in0 = rlo[j - 1], in_cnt = rhi[j - 1] - rlo[j - 1];
            const int o0 = rlo[j], o1 = rhi[j];
            const int own0 = (chunk * p.tk) << (K - j), own1 = min(((chunk + 1) * p.tk) << (K - j), half);
            const int nbt = p.nb_top[j - 1], nbb = p.nb_bot[j - 1];
            double* __restrict__ hib = p.hi[j - 1] + (int64_t)b * p.hi_stride[j - 1];
            double* __restrict__ lob = p.lo + (int64_t)b * p.lo_stride;
            const bool last = j == K;

            // ---- interior outputs: tiles of 32 output positions x 2 bands, one warp per tile ----------------------
            const int ntiles = (o1 - o0 + 31) / 32;
            const int lo_all = max(nbt, o0), hi_all = min(half - nbb, o1);
            // Tiles [t_lo, t_hi) are "fast": all 32 outputs are interior ones of this chunk and every sample of their
            // windows is staged (the conditions are linear in the tile index, so the range is computed once per level).
            int t_lo, t_hi;
            {
                const int need0 = max(lo_all - o0, (HL + SH + in0 + 1) / 2 - o0);          // first admissible i0 - o0
                t_lo = need0 > 0 ? (need0 + 31) / 32 : 0;
                const int lim_out = (hi_all - o0) / 32;                                     // i0 + 32 <= hi_all
                const int lim_in = in_cnt - 56 - 4 * E + HL + SH + in0 - 2 * o0;            // 2 (i0 - o0) <= lim_in
                t_hi = min(lim_out, lim_in >= 0 ? lim_in / 64 + 1 : 0);
                t_hi = max(min(t_hi, ntiles), t_lo);
            }
            auto generic_tile = [&](const int t) {
                const int i0 = o0 + 32 * t;
                const int rel0 = 2 * i0 + 8 * frag_n - HL - SH + KS * frag_k - in0;
                double c0 = 0.0, c1 = 0.0;
                const int ia = i0 + 4 * out_g0 + out_s, ib = ia + 4;
                // samples clamped into the staged range (only outputs that the corner-block code below overwrites, or
                // that lie beyond o1, can touch the clamp), stores checked one by one
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int rel = min(max(rel0 + ES * e, 0), in_cnt - 1);
                    dmma_m8n8k4(c0, c1, afrag[e], in[rel]);
                }
                if (out_band == 0) {
                    if (!last) {
                        if (ia < o1) nxt[ia - o0] = c0;
                        if (ib < o1) nxt[ib - o0] = c1;
                    } else {
                        if (ia >= own0 && ia < own1 && ia >= nbt && ia < half - nbb) lob[ia] = c0;
                        if (ib >= own0 && ib < own1 && ib >= nbt && ib < half - nbb) lob[ib] = c1;
                    }
                } else {
                    if (ia >= own0 && ia < own1 && ia >= nbt && ia < half - nbb) hib[ia] = c0;
                    if (ib >= own0 && ib < own1 && ib >= nbt && ib < half - nbb) hib[ib] = c1;
                }
            };
            for (int t = warp; t < t_lo; t += NW) generic_tile(t);
            for (int t = t_hi + warp; t < ntiles; t += NW) generic_tile(t);
            {
                // fast tiles: running pointers, no range checks besides the owned range of the global stores
                const int tf = t_lo + warp;
                int ia = o0 + 32 * tf + 4 * out_g0 + out_s;
                const double* src = in + (2 * (o0 + 32 * tf) + 8 * frag_n - HL - SH + KS * frag_k - in0);
                double* dsm = nxt + (ia - o0);
                double* dgl = (out_band ? hib : lob) + ia;
                const bool to_smem = out_band == 0 && !last;
                for (int t = tf; t < t_hi; t += NW) {
                    double c0 = 0.0, c1 = 0.0;
                    if constexpr (PERM) {
#pragma unroll
                        for (int e = 0; e < E; e += 2) {
                            const double2 v = *reinterpret_cast<const double2*>(src + e);
                            dmma_m8n8k4(c0, c1, afrag[e], v.x);
                            dmma_m8n8k4(c0, c1, afrag[e + 1], v.y);
                        }
                    } else {
                        double v[E];
#pragma unroll
                        for (int e = 0; e < E; ++e) v[e] = src[4 * e];
#pragma unroll
                        for (int e = 0; e < E; ++e) dmma_m8n8k4(c0, c1, afrag[e], v[e]);
                    }
                    if (to_smem) {
                        dsm[0] = c0; dsm[4] = c1;
                    } else {
                        if (ia >= own0 && ia < own1) dgl[0] = c0;
                        if (ia + 4 >= own0 && ia + 4 < own1) dgl[4] = c1;
                    }
                    src += 64 * NW; dsm += 32 * NW; dgl += 32 * NW; ia += 32 * NW;
                }
            }
            __syncthreads();
            // ---- corner blocks: the dense orthogonalised boundary rows (the CTAs at the two ends of the row) -------
            if (o0 < nbt || o1 > half - nbb) {
                // outputs [o0, min(o1, nbt)) and [max(o0, half - nbb), o1), both bands
                const int nt_ = max(min(o1, nbt) - o0, 0);
                const int b0 = max(o0, half - nbb), nb_ = max(o1 - b0, 0);
                for (int q = tid; q < 2 * (nt_ + nb_); q += NT) {
                    const int band = q & 1, r = q >> 1;
                    const int ii = r < nt_ ? o0 + r : b0 + (r - nt_);
                    const bool top = ii < nbt;
                    const int rr = top ? ii : nbt + (ii - (half - nbb));
                    const int w = top ? p.w_left[j - 1] : p.w_right[j - 1];
                    const int s0 = top ? 0 : nprev - w;
                    const double* __restrict__ blk = (top ? (band ? p.hi_left[j - 1] : p.lo_left[j - 1])
                                                          : (band ? p.hi_right[j - 1] : p.lo_right[j - 1])) + rr * w;
                    double acc = 0.0;
                    for (int c = 0; c < w; ++c) acc = fma(__ldg(blk + c), in[s0 + c - in0], acc);
                    if (band == 0) {
                        if (!last) nxt[ii - o0] = acc;
                        else if (ii >= own0 && ii < own1) lob[ii] = acc;
                    } else if (ii >= own0 && ii < own1) {
                        hib[ii] = acc;
                    }
                }
                __syncthreads();
            }
            in = nxt;
            nxt = (nxt == levA) ? levB : levA;
        }
    }
}

// Host: launch one fused group of k levels on the FP64 tensor cores; false = not applicable (caller falls back).
static bool launch_mat_fwd_dmma(int L, int k, const int64_t* n, const int32_t* nbt, const int32_t* nbb, const int32_t* wl,
                                const int32_t* wr, const double* const* blk_ptrs, const double* x, int64_t xs, int64_t batch,
                                void* const* hi_out, const int64_t* hi_stride, double* lo_out, int64_t lo_stride,
                                const Taps<double>& taps, cudaStream_t st, cudaError_t* err) {
    *err = cudaSuccess;
    if ((L & 1) || L < 2 || L > 16 || k < 1 || k > MATF_MAXK || batch > 65535) return false;
    if (((uintptr_t)x & 15) || (xs & 1) || n[0] >= (int64_t(1) << 30)) return false;
    MatFusedParams<double> p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.x_stride = xs; p.k = k;
    p.n[0] = (int)n[0];
    for (int j = 0; j < k; ++j) {
        if (n[j] & 1) return false;
        p.n[j + 1] = (int)(n[j] / 2);
        if (j + 1 < k && n[j + 1] != n[j] / 2) return false;
        p.hi[j] = (double*)hi_out[j]; p.hi_stride[j] = hi_stride[j];
        p.nb_top[j] = nbt[j]; p.nb_bot[j] = nbb[j]; p.w_left[j] = wl[j]; p.w_right[j] = wr[j];
        p.lo_left[j] = blk_ptrs[4 * j]; p.lo_right[j] = blk_ptrs[4 * j + 1];
        p.hi_left[j] = blk_ptrs[4 * j + 2]; p.hi_right[j] = blk_ptrs[4 * j + 3];
        if (nbt[j] + nbb[j] > p.n[j + 1]) return false;
    }
    p.lo = lo_out; p.lo_stride = lo_stride;
    for (int q = 0; q < L; ++q) { p.flo[q] = taps.lo[L - 1 - q]; p.fhi[q] = taps.hi[L - 1 - q]; }
    const int nk = p.n[k];
    int chunk0 = 2048;
    if (knob_is_set(K_MATF_CHUNK)) { const int v = (int)knob_val(K_MATF_CHUNK, 0); if (v >= 64 && v <= 16384) chunk0 = v; }
    if (n[0] <= 8192 && n[0] > chunk0) chunk0 = (int)n[0];
    int tk = chunk0 >> k;
    if (tk < 4) tk = 4;
    tk = (tk + 3) & ~3;
    if (tk > nk) tk = (nk + 3) & ~3;
    p.tk = tk;
    int cap0 = (tk << k) + ((L + 6) << k) + 64;
    if (cap0 > p.n[0] + 16) cap0 = (p.n[0] + 16 + 3) & ~3;
    cap0 = (cap0 + 3) & ~3;
    p.cap0 = cap0;
    const size_t smem = (size_t)(2 * cap0 + (cap0 / 2 + 16) + (cap0 / 4 + 16)) * sizeof(double);
    if (smem > 200 * 1024) return false;
    const int nchunks = (nk + tk - 1) / tk;
    int cpc = (int)knob_val(K_MATF_CPC, 8);
    if (cpc < 1) cpc = 1;
    while (cpc > 1 && (int64_t)((nchunks + cpc - 1) / cpc) * batch < 4 * 148) cpc /= 2;
    p.cpc = cpc;
    dim3 grid((nchunks + cpc - 1) / cpc, (unsigned)batch);
    const int nt = knob_val(K_MATF_NT, 128) == 256 ? 256 : 128;
    const bool perm = knob_on(K_DMMA_PERM);
#define WTB_MD_LAUNCH(LL, NTT, PP)                                                                     \
    {                                                                                                  \
        cudaError_t e = ensure_dyn_smem(mat_fwd_dmma_kernel<LL, NTT, PP>, 200 * 1024);                 \
        if (e != cudaSuccess) { *err = e; return true; }                                               \
        mat_fwd_dmma_kernel<LL, NTT, PP><<<grid, NTT, smem, st>>>(p);                                  \
    }
#define WTB_MD(LL)                                                                                     \
    case LL:                                                                                           \
        if (nt == 128) { if (perm) WTB_MD_LAUNCH(LL, 128, true) else WTB_MD_LAUNCH(LL, 128, false) }      \
        else { if (perm) WTB_MD_LAUNCH(LL, 256, true) else WTB_MD_LAUNCH(LL, 256, false) }                \
        break;
    switch (L) {
        WTB_MD(2) WTB_MD(4) WTB_MD(6) WTB_MD(8) WTB_MD(10) WTB_MD(12) WTB_MD(14) WTB_MD(16)
        default: return false;
    }
#undef WTB_MD
#undef WTB_MD_LAUNCH
    *err = cudaGetLastError();
    return true;
}

}  // namespace wtb
