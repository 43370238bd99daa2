// msda_bwd_win.cuh -- whole-clip backward with per-block pre-aggregation of grad_value (encoder form).
//
// msda_bwd_kernel is limited by the rate at which reductions can LEAVE an SM (5.3 cycles per 128-byte row for
// red.global.add.v4.f32, benchmarks/micro/smem_accumulate.cu), and in the encoder most of those rows are duplicates:
// the 64 pixel-queries of an 8 x 8 tile sample the same few hundred rows of the coarser levels of every frame.
// Shared-memory FLOAT atomics are CAS loops on sm_100a (14 cycles per row), but shared-memory INTEGER atomics are
// native (ATOMS.ADD) and, with the four words of a lane rotated by its group index so that the four rows of a warp
// instruction fall into different bank octets, cost 1.2 cycles per row.  So:
//
//   * a thread block owns one head and one 8 x 8 pixel tile of one pyramid level (query i == pixel i: encoder
//     self-attention, deformable_transformer.py:184-198) and walks the tile's taps one sampled FRAME at a time;
//   * for every level it keeps a WINDOW of 32-bit fixed-point accumulators in shared memory, placed around the tile's
//     footprint in that level (tile rectangle scaled to the level, grown by `margin` pixels, clipped to the map);
//     levels are served coarsest first until `budget_rows` rows are used -- the finest level usually gets none and
//     keeps scattering straight to global memory, it has the fewest duplicates anyway;
//   * a tap whose four corners lie inside its level's window is accumulated there (16 ATOMS.ADD per lane instead of
//     four 16-byte global reductions); any other tap takes the direct path of msda_bwd_kernel.  Window placement is a
//     performance heuristic only: results do not depend on it beyond the fixed-point rounding below;
//   * after each frame the windows are flushed: rows that received anything are converted back to float, sent to
//     grad_value with ONE vector reduction per row, and cleared.
//
// Fixed point.  One contribution is (attention weight x bilinear weight) x grad_out[channel], at most A x G with
// A = max|attn| (absmax pre-pass over the weight tensors) and G = max|grad_out| over this block's queries and head.
// It is scaled so that A x G maps to 2^kFixBits / (queries x points per slot) and converted with round-to-nearest; a
// window word receives at most queries x points contributions between two flushes, so |sum| < 2^kFixBits: no
// overflow.  With 64 queries x 4 points the quantum is A x G x 2^-22; the sum of a few hundred roundings stays
// 1 - 2e-5 of max|grad_value| at the DeVIS shape (tests: < 1e-4, the north-star bound).  grad_sampling_loc and
// grad_attn_weight do not go through the windows and are bit-identical to msda_bwd_kernel's.
#pragma once
#include "msda_bwd.cuh"

#ifndef DEVIS_BWDW_MIN_BLOCKS
#define DEVIS_BWDW_MIN_BLOCKS 3
#endif

namespace devis {

constexpr int kWinTile = 8;          // queries form kWinTile x kWinTile pixel tiles
constexpr int kWinThreads = 256;     // 32 lane groups; each walks (tile pixels / 32) queries one after the other
constexpr int kFixBits = 30;

struct WinArgs {
    BwdArgs<ClipTable> b;
    int tiles_x[kMaxLevels];     // tiles per row of level l
    int tile_start[kMaxLevels];  // first tile index of level l
    int n_tiles;
    int margin;                  // window margin around the tile footprint, in pixels of the sampled level
    int budget_rows;             // shared-memory window rows per block
    const unsigned *aw_max_bits; // bits of max|attn weight| over both segments (non-negative float)
};

// window of one level: rows [y_lo, y_lo + wh) x columns [x_lo, x_lo + ww) of that level's map, stored row-major from
// window row `base`; ww == 0: the level has no window
struct LevelWindow {
    int x_lo, y_lo, ww, wh, base;
};

// adds round(c * gr[i]) to the lane's i-th word of the window row at shared address `row` (lo[i]: the lane's byte
// offsets inside a row, rotated by its group index so that the four rows of a warp instruction use different banks)
__device__ __forceinline__ void win_add4(unsigned row, const unsigned (&lo)[4], float c, const float (&gr)[4])
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int v = __float_as_int(fmaf(c, gr[i], 12582912.f)) - 0x4B400000;
        asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(row + lo[i]), "r"(v) : "memory");
    }
}

template <bool BF16>
__global__ void __launch_bounds__(kWinThreads, DEVIS_BWDW_MIN_BLOCKS) msda_bwdw_kernel(const WinArgs a)
{
    constexpr int LPG = 8;
    using X = TapExchange<LPG>;
    extern __shared__ int4 s_slot[];
    const BwdArgs<ClipTable> &b = a.b;
    const int outer = blockIdx.y;
    const int L = b.src.L;
    build_slots(s_slot, b.src, b.d, outer, b.n_slots_total);
    int *s_lw = reinterpret_cast<int *>(s_slot + b.n_slots_total);                       // LevelWindow[L] as 5 ints
    unsigned *s_misc = reinterpret_cast<unsigned *>(s_lw + 5 * kMaxLevels);              // [0] bits of max|grad_out|
    float *xbuf_all = reinterpret_cast<float *>(s_misc + 4);
    const int warp = threadIdx.x >> 5;
    float *xbuf = xbuf_all + warp * (2 * X::kWordsPerWarpBuf);
    unsigned *wrec_all = reinterpret_cast<unsigned *>(xbuf_all + (kWinThreads / 32) * 2 * X::kWordsPerWarpBuf);
    unsigned *wrec = wrec_all + warp * 64;                                               // [parity][tap j][group]
    int *win = reinterpret_cast<int *>(wrec_all + (kWinThreads / 32) * 64);              // [budget_rows][32]

    const int M = b.d.M, Lq = b.d.Lq;
    const int j = threadIdx.x % LPG;
    const int g = (threadIdx.x & 31) / LPG;
    const int grp = threadIdx.x / LPG;
    constexpr int NG = kWinThreads / LPG;
    const int tile = blockIdx.x / M, m = blockIdx.x - tile * M;

    // which level and which tile of it
    int lq = 0;
    while (lq + 1 < L && tile >= a.tile_start[lq + 1]) ++lq;
    const int tl = tile - a.tile_start[lq];
    const int ty0 = (tl / a.tiles_x[lq]) * kWinTile, tx0 = (tl - (tl / a.tiles_x[lq]) * a.tiles_x[lq]) * kWinTile;
    const int Hq = b.src.H[lq], Wq = b.src.W[lq];
    const int th = min(kWinTile, Hq - ty0), tw = min(kWinTile, Wq - tx0);

    // windows: the tile rectangle mapped into every level, coarsest level first while the budget lasts
    if (threadIdx.x == 0) {
        int used = 0;
        for (int l = L - 1; l >= 0; --l) {
            const int H = b.src.H[l], W = b.src.W[l];
            // pixel centres of the tile's first / last column and row, in level-l pixel coordinates
            const float fx0 = ((float)tx0 + 0.5f) / (float)Wq * (float)W - 0.5f;
            const float fx1 = ((float)(tx0 + tw - 1) + 0.5f) / (float)Wq * (float)W - 0.5f;
            const float fy0 = ((float)ty0 + 0.5f) / (float)Hq * (float)H - 0.5f;
            const float fy1 = ((float)(ty0 + th - 1) + 0.5f) / (float)Hq * (float)H - 0.5f;
            int x_lo = max((int)floorf(fx0) - a.margin, 0), x_hi = min((int)floorf(fx1) + 1 + a.margin, W - 1);
            int y_lo = max((int)floorf(fy0) - a.margin, 0), y_hi = min((int)floorf(fy1) + 1 + a.margin, H - 1);
            int ww = x_hi - x_lo + 1, wh = y_hi - y_lo + 1;
            if (ww > 255 || ww < 1 || wh < 1 || used + ww * wh > a.budget_rows || used + ww * wh > 0xffff) ww = wh = 0;
            s_lw[5 * l + 0] = x_lo;
            s_lw[5 * l + 1] = y_lo;
            s_lw[5 * l + 2] = ww;
            s_lw[5 * l + 3] = wh;
            s_lw[5 * l + 4] = used;
            used += ww * wh;
        }
        s_misc[0] = 0u;
        s_misc[1] = (unsigned)used;
    }
    __syncthreads();
    const int used_rows = (int)s_misc[1];
    for (int i = threadIdx.x; i < used_rows * 32; i += kWinThreads) win[i] = 0;

    // queries of this group: tile pixels grp, grp + NG, ...
    const int n_pix = th * tw;
    const int QPC = (kWinTile * kWinTile + NG - 1) / NG;
    float gmax = 0.f;
    for (int i = 0; i < QPC; ++i) {
        const int p = grp + i * NG;
        if (p < n_pix) {
            const int q = b.src.lsi[lq] + (ty0 + p / tw) * Wq + tx0 + (p - (p / tw) * tw);
            const size_t row = ((size_t)outer * Lq + q) * M + m;
            const float4 v = BF16 ? ldg_bf16x4(reinterpret_cast<const uint2 *>(b.grad_out) + row * LPG + j)
                                  : ldg_f4(reinterpret_cast<const float4 *>(b.grad_out) + row * LPG + j);
            gmax = fmaxf(gmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    // NaN compares false everywhere: a block whose grad_out holds NaN/Inf must not use the fixed-point path at all
    const bool gbad = !(gmax <= 3.0e38f);
    if ((threadIdx.x & 31) == 0) atomicMax(&s_misc[0], gbad ? 0x7f800000u : __float_as_uint(gmax));
    __syncthreads();

    // fixed-point scale: contribution <= A*G  ->  at most 2^kFixBits / (queries * points) after scaling
    const float G = __uint_as_float(s_misc[0]);
    const float A = __uint_as_float(a.aw_max_bits[0]);
    int pmax = b.seg[0].P;
    if (b.n_seg > 1) pmax = max(pmax, b.seg[1].P);
    int cbits = 0;
    while ((1 << cbits) < kWinTile * kWinTile * pmax) ++cbits;
    const int room = min(22, kFixBits - cbits);                 // the magic-number conversion is exact up to 2^22
    const float bound = A * G;
    const bool fix_ok = room >= 8 && bound > 1.0e-30f && bound < 1.0e30f && used_rows > 0;
    // total scale 2^room / (A*G) = go_scale * kappa: grad_out is scaled by a power of two (exact), the per-corner
    // factor by kappa in (0.5, 1]
    int be = 0;
    const float bf = fix_ok ? frexpf(bound, &be) : 0.5f;       // bound = bf * 2^be, bf in [0.5, 1)
    const float go_scale = fix_ok ? ldexpf(1.f, room - be + 1) : 0.f;
    const float kappa = fix_ok ? 0.5f / bf : 0.f;
    const float inv_scale = fix_ok ? ldexpf(bound, -room) : 0.f;

    constexpr unsigned kQuadBytes = BF16 ? 8u : 16u;
    const unsigned rowbytes = (unsigned)(M * LPG) * kQuadBytes;
    const char *vbase = reinterpret_cast<const char *>(b.value) + (size_t)(m * LPG + j) * kQuadBytes;
    asm volatile("" : "+l"(vbase));
    char *gvb = b.grad_value ? reinterpret_cast<char *>(b.grad_value) + (size_t)(m * LPG + j) * 16u : nullptr;
    constexpr unsigned kGvShift = BF16 ? 1u : 0u;
    const bool use_win = fix_ok && gvb != nullptr;
    // this lane's word inside a window row for its i-th component: block (i + g) & 3, word j
    const unsigned wbase = (unsigned)__cvta_generic_to_shared(win);
    unsigned lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) lo[i] = (unsigned)((((i + g) & 3) << 3) + j) * 4u;

    int slot_base = 0, parity = 0;
    for (int sg = 0; sg < b.n_seg; ++sg) {
        const int P = b.seg[sg].P, K = b.seg[sg].n_slots * P;
        const int KF = L * P;                                   // taps of one sampled frame
        const float *loc = reinterpret_cast<const float *>(b.seg[sg].loc);
        const float *aw = reinterpret_cast<const float *>(b.seg[sg].aw);
        float *gloc = reinterpret_cast<float *>(b.seg[sg].grad_loc);
        float *gaw = reinterpret_cast<float *>(b.seg[sg].grad_aw);
        for (int kf = 0; kf < K; kf += KF) {
            for (int qi = 0; qi < QPC; ++qi) {
                const int p = grp + qi * NG;
                const bool qlive = p < n_pix;
                const int q = qlive ? b.src.lsi[lq] + (ty0 + p / tw) * Wq + tx0 + (p - (p / tw) * tw) : 0;
                const size_t row = ((size_t)outer * Lq + q) * M + m;
                float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
                if (qlive)
                    gg = BF16 ? ldg_bf16x4(reinterpret_cast<const uint2 *>(b.grad_out) + row * LPG + j)
                              : ldg_f4(reinterpret_cast<const float4 *>(b.grad_out) + row * LPG + j);
                // rotated, scaled copy for the window path: gr[i] = component (i + g) & 3
                float gr[4];
                {
                    const float gk = go_scale * kappa;
                    const float c0 = gg.x * gk, c1 = gg.y * gk, c2 = gg.z * gk, c3 = gg.w * gk;
                    gr[0] = g == 0 ? c0 : g == 1 ? c1 : g == 2 ? c2 : c3;
                    gr[1] = g == 0 ? c1 : g == 1 ? c2 : g == 2 ? c3 : c0;
                    gr[2] = g == 0 ? c2 : g == 1 ? c3 : g == 2 ? c0 : c1;
                    gr[3] = g == 0 ? c3 : g == 1 ? c0 : g == 2 ? c1 : c2;
                }
                for (int k0 = kf; k0 < kf + KF; k0 += LPG) {
                    const int k = k0 + j;
                    const bool klive = k < kf + KF;
                    const int slot = slot_base + (klive ? k / P : 0);
                    const int4 sl = s_slot[slot];
                    const bool live = klive && qlive;
                    float2 xy = make_float2(0.f, 0.f);
                    float w = 0.f;
                    if (live) {
                        xy = __ldg(reinterpret_cast<const float2 *>(loc + row * K * 2) + k);
                        w = __ldg(aw + row * K + k);
                    }
                    const TapGeom t = tap_geometry(xy.x, xy.y, sl, live);
                    float *buf = xbuf + parity * X::kWordsPerWarpBuf;
                    unsigned *wb = wrec + parity * 32;
                    parity ^= 1;
                    X::publish(buf, j, g, t, w, rowbytes);
                    {   // window record of this tap: TL window row | dx << 16 | (dy * ww) << 17, or ~0
                        unsigned rec = 0xffffffffu;
                        if (use_win && t.ok != 0u) {
                            const int l = (slot - slot_base) % L;
                            const int x_lo = s_lw[5 * l], y_lo = s_lw[5 * l + 1], ww = s_lw[5 * l + 2], wh = s_lw[5 * l + 3];
                            const int dx = t.rTR - t.rTL;                        // 0 / 1 column
                            const int dy = (t.rBL != t.rTL) ? 1 : 0;             // 0 / 1 row
                            const int wx = t.x0c - x_lo, wy = t.y0c - y_lo;
                            if (ww > 0 && wx >= 0 && wx + dx < ww && wy >= 0 && wy + dy < wh)
                                rec = (unsigned)(s_lw[5 * l + 4] + wy * ww + wx) | ((unsigned)dx << 16) |
                                      ((unsigned)(dy * ww) << 17);
                        }
                        wb[j * 4 + g] = rec;
                    }
                    __syncwarp();

                    float dsum[LPG][4];
#pragma unroll
                    for (int jj = 0; jj < LPG; ++jj) {
                        uint4 off;
                        float4 c;
                        X::fetch(buf, jj, g, off, c);
                        const unsigned rec = wb[jj * 4 + g];
                        float4 v00, v01, v10, v11;
                        if (BF16) {
                            v00 = ldg_bf16x4(reinterpret_cast<const uint2 *>(vbase + off.x));
                            v01 = ldg_bf16x4(reinterpret_cast<const uint2 *>(vbase + off.y));
                            v10 = ldg_bf16x4(reinterpret_cast<const uint2 *>(vbase + off.z));
                            v11 = ldg_bf16x4(reinterpret_cast<const uint2 *>(vbase + off.w));
                        } else {
                            v00 = ldg_f4(reinterpret_cast<const float4 *>(vbase + off.x));
                            v01 = ldg_f4(reinterpret_cast<const float4 *>(vbase + off.y));
                            v10 = ldg_f4(reinterpret_cast<const float4 *>(vbase + off.z));
                            v11 = ldg_f4(reinterpret_cast<const float4 *>(vbase + off.w));
                        }
                        dsum[jj][0] = fmaf(v00.w, gg.w, fmaf(v00.z, gg.z, fmaf(v00.y, gg.y, v00.x * gg.x)));
                        dsum[jj][1] = fmaf(v01.w, gg.w, fmaf(v01.z, gg.z, fmaf(v01.y, gg.y, v01.x * gg.x)));
                        dsum[jj][2] = fmaf(v10.w, gg.w, fmaf(v10.z, gg.z, fmaf(v10.y, gg.y, v10.x * gg.x)));
                        dsum[jj][3] = fmaf(v11.w, gg.w, fmaf(v11.z, gg.z, fmaf(v11.y, gg.y, v11.x * gg.x)));
                        if (rec != 0xffffffffu) {
                            // 4 corners x 4 components: value = round(c * kappa * go * go_scale) by the 1.5 * 2^23 trick;
                            // a corner outside the map has c == 0 and adds 0 to its (clamped, in-window) row
                            const unsigned aTL = wbase + (rec & 0xffffu) * 128u;
                            const unsigned dxb = ((rec >> 16) & 1u) * 128u, dyb = (rec >> 17) * 128u;
                            win_add4(aTL, lo, c.x, gr);
                            win_add4(aTL + dxb, lo, c.y, gr);
                            win_add4(aTL + dyb, lo, c.z, gr);
                            win_add4(aTL + dyb + dxb, lo, c.w, gr);
                        } else if (gvb) {
                            if (c.x != 0.f) red_add_f4(reinterpret_cast<float *>(gvb + ((size_t)off.x << kGvShift)), c.x * gg.x, c.x * gg.y, c.x * gg.z, c.x * gg.w);
                            if (c.y != 0.f) red_add_f4(reinterpret_cast<float *>(gvb + ((size_t)off.y << kGvShift)), c.y * gg.x, c.y * gg.y, c.y * gg.z, c.y * gg.w);
                            if (c.z != 0.f) red_add_f4(reinterpret_cast<float *>(gvb + ((size_t)off.z << kGvShift)), c.z * gg.x, c.z * gg.y, c.z * gg.z, c.z * gg.w);
                            if (c.w != 0.f) red_add_f4(reinterpret_cast<float *>(gvb + ((size_t)off.w << kGvShift)), c.w * gg.x, c.w * gg.y, c.w * gg.z, c.w * gg.w);
                        }
                    }

                    float Ac[4];
                    reduce_scatter_taps<LPG>(dsum, j, Ac);
                    if (live) {
                        const bool hit = t.ok != 0u;
                        const float hh = (t.ok & 1u) ? t.hh : 0.f, lh = (t.ok & 2u) ? t.lh : 0.f;
                        const float hw = (t.ok & 4u) ? t.hw : 0.f, lw = (t.ok & 8u) ? t.lw : 0.f;
                        const float l_in = (t.ok & 4u) ? 1.f : 0.f, r_in = (t.ok & 8u) ? 1.f : 0.f;
                        const float t_in = (t.ok & 1u) ? 1.f : 0.f, b_in = (t.ok & 2u) ? 1.f : 0.f;
                        const float val = hh * (hw * Ac[0] + lw * Ac[1]) + lh * (hw * Ac[2] + lw * Ac[3]);
                        const float gx = hh * (r_in * Ac[1] - l_in * Ac[0]) + lh * (r_in * Ac[3] - l_in * Ac[2]);
                        const float gy = hw * (b_in * Ac[2] - t_in * Ac[0]) + lw * (b_in * Ac[3] - t_in * Ac[1]);
                        gaw[row * K + k] = hit ? val : 0.f;
                        reinterpret_cast<float2 *>(gloc + row * K * 2)[k] =
                            hit ? make_float2((float)sl.y * gx * w, (float)sl.x * gy * w) : make_float2(0.f, 0.f);
                    }
                }
            }
            // ---- flush the windows of the frame just sampled -------------------------------------------------
            if (use_win) {
                __syncthreads();
                const int fslot = slot_base + (kf / KF) * L;         // first slot of this frame
                for (int l = 0; l < L; ++l) {
                    const int ww = s_lw[5 * l + 2], wh = s_lw[5 * l + 3];
                    if (ww == 0) continue;
                    const int x_lo = s_lw[5 * l], y_lo = s_lw[5 * l + 1], base = s_lw[5 * l + 4];
                    const int4 sl = s_slot[fslot + l];
                    // a warp takes window rows wy = warp, warp + 8, ...; its 4 groups take 4 neighbouring columns
                    for (int wy = warp; wy < wh; wy += kWinThreads / 32) {
                        const size_t vrow0 = (size_t)sl.z + (size_t)(y_lo + wy) * sl.y + x_lo;
                        for (int wx0 = 0; wx0 < ww; wx0 += 4) {
                            const int wx = wx0 + g;
                            const bool valid = wx < ww;
                            const unsigned arow = wbase + (unsigned)(base + wy * ww + (valid ? wx : 0)) * 128u;
                            int tq[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(tq[i]) : "r"(arow + lo[i]));
                                if (!valid) tq[i] = 0;
                            }
                            const bool nz = (tq[0] | tq[1] | tq[2] | tq[3]) != 0;
                            const unsigned any = __ballot_sync(0xffffffffu, nz);
                            if ((any >> (g * 8)) & 0xffu) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) asm volatile("st.shared.s32 [%0], %1;" ::"r"(arow + lo[i]), "r"(0) : "memory");
                                // undo the rotation: component c sits in tq[(c - g) & 3]
                                const int c0 = g == 0 ? tq[0] : g == 1 ? tq[3] : g == 2 ? tq[2] : tq[1];
                                const int c1 = g == 0 ? tq[1] : g == 1 ? tq[0] : g == 2 ? tq[3] : tq[2];
                                const int c2 = g == 0 ? tq[2] : g == 1 ? tq[1] : g == 2 ? tq[0] : tq[3];
                                const int c3 = g == 0 ? tq[3] : g == 1 ? tq[2] : g == 2 ? tq[1] : tq[0];
                                red_add_f4(reinterpret_cast<float *>(gvb + (vrow0 + wx) * (size_t)(M * LPG) * 16u),
                                           (float)c0 * inv_scale, (float)c1 * inv_scale, (float)c2 * inv_scale,
                                           (float)c3 * inv_scale);
                            }
                        }
                    }
                }
                __syncthreads();
            }
        }
        slot_base += b.seg[sg].n_slots;
    }
}

}  // namespace devis
