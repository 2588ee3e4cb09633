// fused2d_wpair.cuh -- TWO consecutive analysis levels of the 2-D transform in one kernel, built from
// independent warps (float32, even filter length <= 8).
//
// Replaces two turns of the reference's level loop (src/ptwt/conv_transform_2.py:142-149:
// F.pad -> conv2d(4 x [L x L], stride 2) -> split, with res_ll fed back) without writing the first
// level's approximation band to HBM: with one launch per level that band makes a round trip
// (+25 % traffic at the first level, 1.33x over the whole pyramid).
//
// Structure (no CTA-wide barrier anywhere; a CTA is ONE warp):
//   * a warp owns a strip of 128 level-1 columns (lane <-> 4 adjacent columns) and a segment of rows
//     and marches down it; input rows arrive by TMA (cp.async.bulk.tensor, 8-byte elements so that
//     the 264-sample rows fit one box) in groups of 4 rows into a private 3-stage ring, completion on
//     the warp's own mbarriers; the warp that consumed a stage re-arms it;
//   * row pass: each lane slides the L taps over its 16-sample window (4 LDS.128 per row);
//   * column pass WITHOUT a shared-memory ring: every pair of input rows is scattered into the L/2
//     output rows it contributes to, held as FFMA2 accumulators in registers
//     (acc[i] += dec[2(i-k)+1] * row[2k] + dec[2(i-k)] * row[2k+1]); one output row completes per
//     pair and goes straight to HBM (three detail bands, 128-bit stores, 512 contiguous bytes/warp);
//   * the completed approximation row goes to an 8-row ring in shared memory (the only exchange
//     between lanes); level 2 reads it back with lane <-> 2 level-2 columns, same row pass, same
//     scatter accumulators, and stores its four bands;
//   * boundary extension: out-of-range input samples are patched into the staged tile from the
//     extension source (all modes but periodic; zero fill is TMA's out-of-bounds fill); the level-2
//     extension of the approximation band is served from the ring (rows) and by patching the ring
//     rows of the edge strips (columns), so level 2 sees exactly ext(cA1) like the reference.
//
// Horizontal halo: a strip computes HL1 extra approximation columns on its left (6 % redundant
// arithmetic, the re-read input columns hit L2); vertical halo: a segment restarts both levels
// (3 (L/2-1) row pairs + L-2 approximation rows).
//
// Algorithmic bytes: 4 B * (H*W read + 3*Mh1*Mw1 + 4*Mh2*Mw2 written).
#pragma once

#include "fused2d.cuh"

namespace wtb {

constexpr int WPAIR_MAXSEG = 48;

struct WPairParams {
    const float* x;            // level input [batch, H, W]
    int64_t x_bs, x_rs;
    float* d1;                 // first-level detail bands: k = 1, 2, 3 at d1 + (k-1) * d1_band
    int64_t d1_bs, d1_rs, d1_band;
    float* o2[4];              // second-level bands k = 0 (approximation), 1, 2, 3
    int64_t o2_bs[4], o2_rs[4];
    int H, W, Mh1, Mw1, Mh2, Mw2;
    int mode, batch0;
    int nseg;
    int seg_start[WPAIR_MAXSEG + 1];   // second-level row ranges [seg_start[i], seg_start[i+1]), longest first
    float2 pl[4], ph[4];       // row pass: {dec[L-1-2q], dec[L-2-2q]}
    float2 vl[8], vh[8];       // column pass: {dec[m], dec[m]}
};

template <int L>
struct WPairGeom {
    static constexpr int HALO = L - 2, NA = L / 2;
    static constexpr int HAL = (HALO + 3) / 4 * 4;          // left halo of the staged tile, 16-byte aligned
    static constexpr int OFF1 = HAL - HALO;
    static constexpr int HL1 = HAL;                         // approximation columns left of the owned ones
    static constexpr int OFF2 = HL1 - HALO;
    static constexpr int TW1 = 128;                         // level-1 columns per strip (4 per lane)
    static constexpr int TW2 = (TW1 - HL1) / 2;             // level-2 columns owned by a strip
    static constexpr int NV1 = OFF1 + L + 6, NV1_4 = (NV1 + 3) / 4;
    static constexpr int TILE_W = 8 * 31 + 4 * NV1_4;       // staged input columns
    static constexpr int NV2 = OFF2 + L + 2, NV2_4 = (NV2 + 3) / 4;
    static constexpr int RP = 4 * 31 + 4 * NV2_4;           // pitch of the approximation ring
    static constexpr int RING = 8;                          // >= L rows (boundary sources stay resident)
    static constexpr int ROWS = 4, NSTG = 3;
    static constexpr int STAGE_BYTES = ROWS * TILE_W * 4;
    static constexpr int STAGE_STRIDE = (STAGE_BYTES + 127) / 128 * 128;
    static constexpr int SMEM = NSTG * STAGE_STRIDE + (RING + 1) * RP * 4 + 64;
    static_assert(L % 2 == 0 && L >= 2 && L <= 8, "wpair kernel: even filter length <= 8");
    static_assert(TILE_W % 4 == 0 && TILE_W / 2 <= 256, "tile row must fit one TMA box of 8-byte elements");
    static_assert(RING >= L, "ring too small for the boundary sources");
};

// lo[c], hi[c] for NC consecutive outputs from the register window w (see row_filter8)
template <int L, int NC, int OFF, int NW>
__device__ __forceinline__ void wp_rowfilt(const float (&w)[NW], const float2* __restrict__ pl,
                                           const float2* __restrict__ ph, float (&lo)[NC], float (&hi)[NC]) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        float2 a = make_float2(0.f, 0.f), h = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < L / 2; ++q) {
            const float2 x = make_float2(w[2 * c + 2 * q + OFF], w[2 * c + 2 * q + OFF + 1]);
            a = ffma2(pl[q], x, a);
            h = ffma2(ph[q], x, h);
        }
        lo[c] = a.x + a.y;
        hi[c] = h.x + h.y;
    }
}

// Scatter one pair of row-filtered rows (index 0: even row 2k, 1: odd row 2k+1) into the L/2 output rows
// k .. k+L/2-1 it contributes to.  Slot s completes with this pair, slot (s + L/2 - 1) % (L/2) starts.
// acc[slot][band][column pair]; band = 0: lo_H lo_W, 1: lo_H hi_W, 2: hi_H lo_W, 3: hi_H hi_W.
// s must be a compile-time constant after unrolling (the accumulators live in registers).
template <int L, int NCP>
__device__ __forceinline__ void wp_scatter(float2 (&acc)[L / 2][4][NCP], const float2 (&lo)[2][NCP],
                                           const float2 (&hi)[2][NCP], const WPairParams& p, const int s) {
    constexpr int NA = L / 2;
#pragma unroll
    for (int d = 0; d < NA; ++d) {
        const int slot = (s + d) % NA;
        const float2 tl0 = p.vl[2 * d + 1], tl1 = p.vl[2 * d], th0 = p.vh[2 * d + 1], th1 = p.vh[2 * d];
#pragma unroll
        for (int cp = 0; cp < NCP; ++cp) {
            const float2 z = make_float2(0.f, 0.f);
            const float2 a0 = d == NA - 1 ? z : acc[slot][0][cp], a1 = d == NA - 1 ? z : acc[slot][1][cp];
            const float2 a2 = d == NA - 1 ? z : acc[slot][2][cp], a3 = d == NA - 1 ? z : acc[slot][3][cp];
            acc[slot][0][cp] = ffma2(tl1, lo[1][cp], ffma2(tl0, lo[0][cp], a0));
            acc[slot][1][cp] = ffma2(tl1, hi[1][cp], ffma2(tl0, hi[0][cp], a1));
            acc[slot][2][cp] = ffma2(th1, lo[1][cp], ffma2(th0, lo[0][cp], a2));
            acc[slot][3][cp] = ffma2(th1, hi[1][cp], ffma2(th0, hi[0][cp], a3));
        }
    }
}

// Row pass + scatter of one level-1 step: rows tr, tr + TILE_W of the staged tile (lane window)
template <int L>
__device__ __forceinline__ void wp_l1_compute(const float* __restrict__ tr, float2 (&acc)[L / 2][4][2],
                                              const WPairParams& p, const int s) {
    using Gm = WPairGeom<L>;
    float2 lo[2][2], hi[2][2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float w[4 * Gm::NV1_4];
#pragma unroll
        for (int q = 0; q < Gm::NV1_4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(tr + r * Gm::TILE_W + 4 * q);
            w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
        }
        float l4[4], h4[4];
        wp_rowfilt<L, 4, Gm::OFF1>(w, p.pl, p.ph, l4, h4);
        lo[r][0] = make_float2(l4[0], l4[1]); lo[r][1] = make_float2(l4[2], l4[3]);
        hi[r][0] = make_float2(h4[0], h4[1]); hi[r][1] = make_float2(h4[2], h4[3]);
    }
    wp_scatter<L, 2>(acc, lo, hi, p, s);
}

// Row pass + scatter of one level-2 step: approximation rows rA (even), rB (odd) of the ring (lane window)
template <int L>
__device__ __forceinline__ void wp_l2_compute(const float* __restrict__ rA, const float* __restrict__ rB,
                                              float2 (&acc)[L / 2][4][1], const WPairParams& p, const int s) {
    using Gm = WPairGeom<L>;
    float2 lo[2][1], hi[2][1];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const float* src = r ? rB : rA;
        float w[4 * Gm::NV2_4];
#pragma unroll
        for (int q = 0; q < Gm::NV2_4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(src + 4 * q);
            w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
        }
        float l2[2], h2[2];
        wp_rowfilt<L, 2, Gm::OFF2>(w, p.pl, p.ph, l2, h2);
        lo[r][0] = make_float2(l2[0], l2[1]);
        hi[r][0] = make_float2(h2[0], h2[1]);
    }
    wp_scatter<L, 1>(acc, lo, hi, p, s);
}

// logical slot s becomes physical slot 0 (the fast loop uses compile-time slots starting from 0)
template <int NA, int NCP>
__device__ __forceinline__ void wp_rotate(float2 (&acc)[NA][4][NCP], const int s) {
    if (NA == 1 || s == 0) return;
    float2 t[NA][4][NCP];
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < NCP; ++c) t[i][k][c] = acc[i][k][c];
#pragma unroll
    for (int r = 1; r < NA; ++r)
        if (s == r) {
#pragma unroll
            for (int i = 0; i < NA; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int c = 0; c < NCP; ++c) acc[i][k][c] = t[(i + r) % NA][k][c];
        }
}

template <int L>
__global__ void __launch_bounds__(32, 12)
fwd2d_wpair_kernel(const __grid_constant__ WPairParams p, const __grid_constant__ CUtensorMap tmap) {
    using Gm = WPairGeom<L>;
    constexpr int HALO = Gm::HALO, NA = Gm::NA, HAL = Gm::HAL, HL1 = Gm::HL1, TW1 = Gm::TW1, TW2 = Gm::TW2;
    constexpr int TILE_W = Gm::TILE_W, RP = Gm::RP, RING = Gm::RING, ROWS = Gm::ROWS, NSTG = Gm::NSTG;
    constexpr int STG_F = Gm::STAGE_STRIDE / 4;
    constexpr int FB = NA;                             // groups per fast block: 2 NA level-1 steps, NA level-2 steps

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_tile = reinterpret_cast<float*>(smem_raw);                              // [NSTG][ROWS][TILE_W]
    float* s_ring = reinterpret_cast<float*>(smem_raw + NSTG * Gm::STAGE_STRIDE);    // [RING + 1][RP], last row = 0
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + (RING + 1) * RP);

    const int lane = threadIdx.x;
    const int b = p.batch0 + blockIdx.y;

    // ---- strip -------------------------------------------------------------------------------------
    const int X0n = blockIdx.x * TW2;                 // nominal first owned level-2 column
    int X0 = X0n;
    {   // every extension source of the approximation columns level 2 reads must lie inside the strip:
        // the right-most strip is shifted left when needed (it still stores its nominal columns only)
        const int lim = (p.Mw1 - L + HL1) / 2;
        if (X0 > lim) X0 = max(lim & ~1, 0);
    }
    const int cA0 = 2 * X0 - HL1;                     // first approximation column computed (multiple of 4)
    const int c_in0 = 2 * cA0 - HAL;                  // first staged input column (multiple of 4)
    const int own1_lo = 2 * X0n, own1_hi = min(2 * (X0n + TW2), p.Mw1);
    const int own2_lo = X0n, own2_hi = min(X0n + TW2, p.Mw2);

    // ---- segment (long segments come first in the grid, the short ones fill the tail) ---------------
    const int Y0 = p.seg_start[blockIdx.z], Y1 = p.seg_start[blockIdx.z + 1];
    // approximation rows [a_start, a_end) are computed here; the bottom extension sources (the last RING rows)
    // are included even when the segment is short
    const int a_start = max(0, min(2 * (Y0 - NA + 1), p.Mh1 - RING));
    const int a_end = min(p.Mh1, 2 * Y1);
    const int n1 = a_end - a_start + NA - 1;          // level-1 steps (pairs of input rows)
    const int ngroups = (n1 + 1) / 2;                 // TMA groups of 4 input rows
    const int r_in0 = 2 * a_start - HALO;             // first staged input row
    const int n2 = Y1 - Y0 + NA - 1;                  // level-2 steps (pairs of approximation rows)

    if (lane == 0) {
        tma_prefetch_desc(&tmap);
#pragma unroll
        for (int s = 0; s < NSTG; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    for (int i = lane; i < RP; i += 32) s_ring[RING * RP + i] = 0.f;
    __syncwarp();
    if (lane == 0) {
        for (int s = 0; s < NSTG && s < ngroups; ++s) {
            mbar_expect_tx(&bars[s], (uint32_t)Gm::STAGE_BYTES);
            tma_load_3d(s_tile + s * STG_F, &tmap, &bars[s], c_in0 / 2, r_in0 + s * ROWS, b);
        }
    }

    const float* __restrict__ xb = p.x + (int64_t)b * p.x_bs;
    const bool need_patch = (p.mode != WT_MODE_ZERO) || (p.W & 1);
    const bool edge_in = need_patch && (c_in0 < 0 || c_in0 + TILE_W > p.W);   // warp-uniform
    const bool edge_a = cA0 < 0 || cA0 + TW1 > p.Mw1;                         // warp-uniform

    // level-1 stores: this lane's 4 columns, running pointer = row `produced` of band 1
    const int col1 = cA0 + 4 * lane;
    const bool store1 = col1 >= own1_lo && col1 < own1_hi;
    float* pd1 = p.d1 + (int64_t)b * p.d1_bs + (int64_t)a_start * p.d1_rs + col1;
    // level-2 stores: this lane's 2 columns, running pointers = row K2 of the four bands
    const int col2 = X0 + 2 * lane;
    const bool store2 = col2 >= own2_lo && col2 < own2_hi;
    float* po2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) po2[k] = p.o2[k] + (int64_t)b * p.o2_bs[k] + (int64_t)Y0 * p.o2_rs[k] + col2;
    // ring patch table of the edge strips: source index (within the ring row) of each out-of-range column
    int psrc[4] = {-2, -2, -2, -2};
    if (edge_a) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = col1 + j;
            if (col < 0 || col >= p.Mw1) {
                const int s = ext_index32(col, p.Mw1, p.mode);
                psrc[j] = (s >= cA0 && s < cA0 + TW1) ? s - cA0 : -1;
            }
        }
    }

    float2 acc1[NA][4][2];
    float2 acc2[NA][4][1];
#pragma unroll
    for (int s = 0; s < NA; ++s)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc1[s][k][0] = make_float2(0.f, 0.f); acc1[s][k][1] = make_float2(0.f, 0.f);
            acc2[s][k][0] = make_float2(0.f, 0.f);
        }

    int s1 = 0, s2 = 0, j2 = 0, produced = a_start;
    int stage = 0;
    uint32_t par = 0;
    float* const ring_lane = s_ring + 4 * lane;
    const float* const tile_lane = s_tile + 8 * lane;

    // columns of the staged tile outside the image: the L samples next to each border are all any valid output reads
    auto patch_cols = [&](float* tile, const int rbase) {
        const int nl = c_in0 < 0 ? -c_in0 : 0;
        const int l0 = max(nl - L, 0);
        const int r0 = min(max(p.W - c_in0, 0), TILE_W), r1 = min(r0 + L, TILE_W);
        const int wl = nl - l0, wb = wl + (r1 - r0);
        for (int idx = lane; idx < ROWS * wb; idx += 32) {
            const int rr = idx / wb, q = idx - rr * wb;
            const int t = q < wl ? l0 + q : r0 + (q - wl);
            const int sc = ext_index32(c_in0 + t, p.W, p.mode);
            tile[rr * TILE_W + t] = sc >= 0 ? __ldg(xb + (int64_t)(rbase + rr) * p.x_rs + sc) : 0.f;
        }
        __syncwarp();
    };
    // ring write + detail stores of the level-1 row completed in slot S (compile-time), then the column extension
    // of the approximation row in the edge strips
#define WTB_WP_L1_OUT(S, STORE)                                                                                         \
    {                                                                                                                   \
        float* rrow = ring_lane + (produced & (RING - 1)) * RP;                                                          \
        *reinterpret_cast<float4*>(rrow) = make_float4(acc1[S][0][0].x, acc1[S][0][0].y, acc1[S][0][1].x, acc1[S][0][1].y); \
        if (STORE) {                                                                                                    \
            *reinterpret_cast<float4*>(pd1) = make_float4(acc1[S][1][0].x, acc1[S][1][0].y, acc1[S][1][1].x, acc1[S][1][1].y); \
            *reinterpret_cast<float4*>(pd1 + p.d1_band) =                                                               \
                make_float4(acc1[S][2][0].x, acc1[S][2][0].y, acc1[S][2][1].x, acc1[S][2][1].y);                        \
            *reinterpret_cast<float4*>(pd1 + 2 * p.d1_band) =                                                           \
                make_float4(acc1[S][3][0].x, acc1[S][3][0].y, acc1[S][3][1].x, acc1[S][3][1].y);                        \
        }                                                                                                               \
        pd1 += p.d1_rs;                                                                                                 \
        if (edge_a) {                                                                                                   \
            __syncwarp();                                                                                               \
            float* row0 = rrow - 4 * lane;                                                                              \
            _Pragma("unroll") for (int jj = 0; jj < 4; ++jj)                                                            \
                if (psrc[jj] != -2) rrow[jj] = psrc[jj] >= 0 ? row0[psrc[jj]] : 0.f;                                    \
        }                                                                                                               \
        ++produced;                                                                                                     \
    }
#define WTB_WP_L2_OUT(S, STORE)                                                                                         \
    {                                                                                                                   \
        if (STORE) {                                                                                                    \
            *reinterpret_cast<float2*>(po2[0]) = acc2[S][0][0]; *reinterpret_cast<float2*>(po2[1]) = acc2[S][1][0];     \
            *reinterpret_cast<float2*>(po2[2]) = acc2[S][2][0]; *reinterpret_cast<float2*>(po2[3]) = acc2[S][3][0];     \
        }                                                                                                               \
        _Pragma("unroll") for (int k = 0; k < 4; ++k) po2[k] += p.o2_rs[k];                                             \
    }

    int g = 0;
    while (g < ngroups) {
        // ================= fast blocks: FB groups of the steady state, every slot a compile-time constant ==========
        {
            const int rb = r_in0 + g * ROWS;
            const int K2n = Y0 - (NA - 1) + j2;                // next level-2 pair
            bool fast = rb >= 0 && rb + FB * ROWS <= p.H &&            // staged rows inside the image
                        2 * g >= NA - 1 && 2 * (g + FB) <= n1 &&       // every step completes a row
                        produced >= 2 * Y0 &&                          // every completed row is stored
                        j2 >= NA - 1 && j2 + FB <= n2 &&               // every level-2 step completes a row
                        K2n >= 0 && 2 * (K2n + FB) <= p.Mh1 &&         // level-2 sources inside the band ...
                        (2 * K2n + 1 == produced || 2 * K2n + 1 == produced + 1);   // ... one pair per group
            if (fast) {
                wp_rotate<NA, 2>(acc1, s1);
                wp_rotate<NA, 1>(acc2, s2);
                s1 = 0; s2 = 0;
                int l2row = 2 * K2n;
                do {
#pragma unroll
                    for (int u = 0; u < FB; ++u) {
                        float* tile = s_tile + stage * STG_F;
                        mbar_wait(&bars[stage], par);
                        if (edge_in) patch_cols(tile, r_in0 + (g + u) * ROWS);
                        const float* tl = tile_lane + stage * STG_F;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            wp_l1_compute<L>(tl + (2 * h) * TILE_W, acc1, p, (2 * u + h) % NA);
                            WTB_WP_L1_OUT((2 * u + h) % NA, store1)
                        }
                        __syncwarp();
                        if (lane == 0 && g + u + NSTG < ngroups) {
                            fence_proxy_async();
                            mbar_expect_tx(&bars[stage], (uint32_t)Gm::STAGE_BYTES);
                            tma_load_3d(tile, &tmap, &bars[stage], c_in0 / 2, r_in0 + (g + u + NSTG) * ROWS, b);
                        }
                        if (++stage == NSTG) { stage = 0; par ^= 1u; }
                        wp_l2_compute<L>(ring_lane + (l2row & (RING - 1)) * RP, ring_lane + ((l2row + 1) & (RING - 1)) * RP,
                                         acc2, p, u);
                        WTB_WP_L2_OUT(u, store2)
                        l2row += 2;
                        __syncwarp();
                    }
                    g += FB;
                    j2 += FB;
                } while (r_in0 + (g + FB) * ROWS <= p.H && 2 * (g + FB) <= n1 && j2 + FB <= n2 &&
                         l2row + 2 * FB <= p.Mh1);
                continue;
            }
        }
        // ================= generic group: prologue / epilogue / borders ================================================
        {
            float* tile = s_tile + stage * STG_F;
            mbar_wait(&bars[stage], par);
            const int rbase = r_in0 + g * ROWS;
            if (need_patch) {
                const bool rows_oob = rbase < 0 || rbase + ROWS > p.H;
                if (rows_oob) {
                    for (int idx = lane; idx < ROWS * TILE_W; idx += 32) {
                        const int rr = idx / TILE_W, t = idx - rr * TILE_W;
                        const int r = rbase + rr, c = c_in0 + t;
                        if (r < 0 || r >= p.H || c < 0 || c >= p.W) {
                            const int sr = ext_index32(r, p.H, p.mode), sc = ext_index32(c, p.W, p.mode);
                            tile[idx] = (sr >= 0 && sc >= 0) ? __ldg(xb + (int64_t)sr * p.x_rs + sc) : 0.f;
                        }
                    }
                    __syncwarp();
                } else if (edge_in) {
                    patch_cols(tile, rbase);
                }
            }
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * g + h;
                if (j < n1) {
                    const float* tr = tile + (2 * h) * TILE_W + 8 * lane;
                    const bool valid = j >= NA - 1;
                    const bool st = store1 && produced >= 2 * Y0;
                    switch (s1) {
                        case 0: wp_l1_compute<L>(tr, acc1, p, 0); if (valid) WTB_WP_L1_OUT(0, st) break;
                        case 1: if constexpr (NA > 1) { wp_l1_compute<L>(tr, acc1, p, (NA > 1 ? 1 : 0)); if (valid) WTB_WP_L1_OUT((NA > 1 ? 1 : 0), st) } break;
                        case 2: if constexpr (NA > 2) { wp_l1_compute<L>(tr, acc1, p, (NA > 2 ? 2 : 0)); if (valid) WTB_WP_L1_OUT((NA > 2 ? 2 : 0), st) } break;
                        default: if constexpr (NA > 3) { wp_l1_compute<L>(tr, acc1, p, (NA > 3 ? 3 : 0)); if (valid) WTB_WP_L1_OUT((NA > 3 ? 3 : 0), st) } break;
                    }
                    s1 = s1 + 1 == NA ? 0 : s1 + 1;
                }
            }
            __syncwarp();   // tile consumed by every lane; ring rows visible
            if (lane == 0 && g + NSTG < ngroups) {
                fence_proxy_async();
                mbar_expect_tx(&bars[stage], (uint32_t)Gm::STAGE_BYTES);
                tma_load_3d(tile, &tmap, &bars[stage], c_in0 / 2, r_in0 + (g + NSTG) * ROWS, b);
            }
            if (++stage == NSTG) { stage = 0; par ^= 1u; }
            // level 2: every pair of approximation rows whose sources exist
            while (j2 < n2) {
                const int K2 = Y0 - (NA - 1) + j2;                     // output row this pair completes
                int sA = 2 * K2, sB = 2 * K2 + 1;
                if (sA < 0 || sB >= p.Mh1) { sA = ext_index32(sA, p.Mh1, p.mode); sB = ext_index32(sB, p.Mh1, p.mode); }
                if (max(sA, sB) >= produced) break;
                const float* rA = ring_lane + (sA < 0 ? RING : (sA & (RING - 1))) * RP;
                const float* rB = ring_lane + (sB < 0 ? RING : (sB & (RING - 1))) * RP;
                const bool valid = j2 >= NA - 1;
                switch (s2) {
                    case 0: wp_l2_compute<L>(rA, rB, acc2, p, 0); if (valid) WTB_WP_L2_OUT(0, store2) break;
                    case 1: if constexpr (NA > 1) { wp_l2_compute<L>(rA, rB, acc2, p, (NA > 1 ? 1 : 0)); if (valid) WTB_WP_L2_OUT((NA > 1 ? 1 : 0), store2) } break;
                    case 2: if constexpr (NA > 2) { wp_l2_compute<L>(rA, rB, acc2, p, (NA > 2 ? 2 : 0)); if (valid) WTB_WP_L2_OUT((NA > 2 ? 2 : 0), store2) } break;
                    default: if constexpr (NA > 3) { wp_l2_compute<L>(rA, rB, acc2, p, (NA > 3 ? 3 : 0)); if (valid) WTB_WP_L2_OUT((NA > 3 ? 3 : 0), store2) } break;
                }
                s2 = s2 + 1 == NA ? 0 : s2 + 1;
                ++j2;
            }
            __syncwarp();   // level-2 reads of the ring precede the next group's writes
            ++g;
        }
    }
#undef WTB_WP_L1_OUT
#undef WTB_WP_L2_OUT
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static bool make_tmap_3d_pairs(CUtensorMap* map, const float* base, int64_t B, int64_t H, int64_t W, int64_t bs, int64_t rs,
                               int box_w_pairs, int box_h) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    if (((uintptr_t)base & 15) || (rs & 3) || (bs & 3)) return false;
    if ((W & 1) && rs <= W) return false;             // the odd sample's partner must be addressable
    if (box_w_pairs > 256 || box_h > 256 || ((box_w_pairs * 8) & 15)) return false;
    cuuint64_t dims[3] = {(cuuint64_t)((W + 1) / 2), (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)(rs * 4), (cuuint64_t)(bs * 4)};
    cuuint32_t box[3] = {(cuuint32_t)box_w_pairs, (cuuint32_t)box_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (B == 1) strides[1] = (cuuint64_t)H * strides[0];
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Returns true when the two levels were launched here (*err carries the launch status).
template <int L>
static bool launch_fwd2d_wpair_t(const float* x, int64_t B, int H, int W, int64_t x_bs, int64_t x_rs, const wt_level& l1,
                                 const wt_level& l2, int mode, const Taps<float>& taps, cudaStream_t st,
                                 uint64_t* launches, cudaError_t* err) {
    using Gm = WPairGeom<L>;
    WPairParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.x_bs = x_bs; p.x_rs = x_rs; p.H = H; p.W = W;
    p.Mh1 = (int)l1.dims[0]; p.Mw1 = (int)l1.dims[1]; p.Mh2 = (int)l2.dims[0]; p.Mw2 = (int)l2.dims[1];
    if (mode == WT_MODE_PERIODIC) return false;
    if (p.Mh1 < 2 * Gm::RING || p.Mw1 < 4 * L || p.Mh2 < L || p.Mw2 < L) return false;
    // level-1 details: three bands with common strides, 128-bit stores
    p.d1 = (float*)l1.details; p.d1_bs = l1.details_batch_stride; p.d1_rs = l1.strides[0]; p.d1_band = l1.band_stride;
    if (l1.strides[1] != 1 || ((uintptr_t)p.d1 & 15) || (p.d1_bs & 3) || (p.d1_rs & 3) || (p.d1_band & 3) ||
        p.d1_rs < (p.Mw1 + 3) / 4 * 4)
        return false;
    // level-2 bands: 64-bit stores
    if (l2.strides[1] != 1 || l2.approx_strides[1] != 1) return false;
    p.o2[0] = (float*)l2.approx; p.o2_bs[0] = l2.approx_batch_stride; p.o2_rs[0] = l2.approx_strides[0];
    for (int k = 1; k < 4; ++k) {
        p.o2[k] = (float*)l2.details + (int64_t)(k - 1) * l2.band_stride;
        p.o2_bs[k] = l2.details_batch_stride; p.o2_rs[k] = l2.strides[0];
    }
    for (int k = 0; k < 4; ++k)
        if (((uintptr_t)p.o2[k] & 7) || (p.o2_bs[k] & 1) || (p.o2_rs[k] & 1) || p.o2_rs[k] < (p.Mw2 + 1) / 2 * 2) return false;
    // strips: the (possibly shifted) last strip must still reach the last level-2 column
    const int nstrip = (p.Mw2 + Gm::TW2 - 1) / Gm::TW2;
    {
        int X0 = (nstrip - 1) * Gm::TW2;
        const int lim = (p.Mw1 - L + Gm::HL1) / 2;
        if (X0 > lim) X0 = lim & ~1;
        if (X0 < 0 || X0 + Gm::TW2 < p.Mw2) return false;
        if (nstrip > 1 && X0 < (nstrip - 2) * Gm::TW2) return false;
    }
    p.mode = mode;
    for (int q = 0; q < L / 2; ++q) {
        p.pl[q] = make_float2(taps.lo[L - 1 - 2 * q], taps.lo[L - 2 - 2 * q]);
        p.ph[q] = make_float2(taps.hi[L - 1 - 2 * q], taps.hi[L - 2 - 2 * q]);
    }
    for (int m = 0; m < L; ++m) {
        p.vl[m] = make_float2(taps.lo[m], taps.lo[m]);
        p.vh[m] = make_float2(taps.hi[m], taps.hi[m]);
    }
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (!make_tmap_3d_pairs(&tmap, x, B, H, W, x_bs, x_rs, Gm::TILE_W / 2, Gm::ROWS)) return false;
    // segments: long ones (restart overhead ~3 %) first, then geometrically shorter ones that fill the tail of the
    // grid (the segment index is the slowest grid dimension, so the short tasks are dispatched last)
    {
        const int forced = (int)knob_val(K_WPAIR_SEG, 0);
        int big = forced > 0 ? forced : 144;
        // enough tasks to fill the machine a few times
        const int64_t want = 3 * 148 * 12;
        while (forced <= 0 && big > 32 && (int64_t)((p.Mh2 + big - 1) / big) * nstrip * B < want) big -= 16;
        int y = 0, n = 0;
        p.seg_start[0] = 0;
        while (y < p.Mh2 && n < WPAIR_MAXSEG - 1) {
            const int left = p.Mh2 - y;
            int sz = big;
            if (forced <= 0 && left <= 3 * big) sz = left / 3 > 16 ? (left + 2) / 3 : (left > 24 ? 16 : left);
            if (sz > left || left - sz < 8) sz = left;
            y += sz;
            p.seg_start[++n] = y;
        }
        if (y < p.Mh2) p.seg_start[n] = p.Mh2;   // table full: the last segment takes the rest
        p.nseg = n;
    }
    const int nseg = p.nseg;
    auto kern = fwd2d_wpair_kernel<L>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gm::SMEM); });
    if (attr_err != cudaSuccess) { *err = attr_err; return true; }
    *err = cudaSuccess;
    for (int64_t b0 = 0; b0 < B; b0 += 65535) {
        p.batch0 = (int)b0;
        const int nb = (int)((B - b0) < 65535 ? (B - b0) : 65535);
        dim3 grid(nstrip, nb, nseg);
        kern<<<grid, 32, Gm::SMEM, st>>>(p, tmap);
        ++*launches;
        *err = cudaGetLastError();
        if (*err != cudaSuccess) return true;
    }
    return true;
}

template <typename T>
static bool try_wpair(const T*, int64_t, int, int, int64_t, int64_t, const wt_level&, const wt_level&, int, int,
                      const Taps<T>&, cudaStream_t, uint64_t*, cudaError_t*) {
    return false;
}
template <>
bool try_wpair<float>(const float* x, int64_t B, int H, int W, int64_t x_bs, int64_t x_rs, const wt_level& l1,
                      const wt_level& l2, int L, int mode, const Taps<float>& taps, cudaStream_t st,
                      uint64_t* launches, cudaError_t* err) {
    if (knob_on(K_NO_WPAIR) || knob_on(K_DISABLE_FUSED)) return false;
    switch (L) {
        case 2: return launch_fwd2d_wpair_t<2>(x, B, H, W, x_bs, x_rs, l1, l2, mode, taps, st, launches, err);
        case 4: return launch_fwd2d_wpair_t<4>(x, B, H, W, x_bs, x_rs, l1, l2, mode, taps, st, launches, err);
        case 6: return launch_fwd2d_wpair_t<6>(x, B, H, W, x_bs, x_rs, l1, l2, mode, taps, st, launches, err);
        case 8: return launch_fwd2d_wpair_t<8>(x, B, H, W, x_bs, x_rs, l1, l2, mode, taps, st, launches, err);
        default: return false;
    }
}

}  // namespace wtb
