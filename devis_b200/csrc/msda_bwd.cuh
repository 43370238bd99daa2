// msda_bwd.cuh -- backward of multi-scale deformable attention, grouped-lane kernel for sm_100a.
//
// Replaces ms_deformable_col2im_gpu_kernel_shm_blocksize_aware_reduce_v1<32>
// (cuda/ms_deform_im2col_cuda.cuh:301-403) and ms_deform_attn_col2im_bilinear (:87-159), where a
// one-warp block owns one (query, head), every lane redoes the tap geometry, each tap costs two
// __syncthreads() plus a serial 32-element reduction on thread 0, and grad_value is scattered with
// 4 scalar atomics per (tap, channel).
//
// Here (same work split as msda_fwd.cuh: LPG lanes x 4 channels per (query, head), tap geometry
// prepared by one lane per tap and broadcast by shuffles):
//   * per tap each lane forms the 4 corner dot products  A_c = sum_ch grad_out[ch] * value_c[ch]
//     over its own 4 channels only (16 FMA); everything else about grad_attn / grad_loc is linear in
//     those four numbers, so it is done AFTER the channel reduction, once per tap, by the lane that
//     prepared the tap:
//         grad_attn = hh*hw*A00 + hh*lw*A01 + lh*hw*A10 + lh*lw*A11
//         grad_x    = W * attn * (hh*(A01-A00) + lh*(A11-A10))
//         grad_y    = H * attn * (hw*(A10-A00) + lw*(A11-A01))
//     (algebraically the sums the reference accumulates per channel at :123-158);
//   * the channel reduction is a reduce-scatter butterfly across the group (LPG taps x 4 numbers in,
//     tap j's 4 totals land on lane j): 3.5 shuffles per tap for LPG=8 instead of 12 for four
//     xor-reductions, no shared memory, no barriers;
//   * grad_value: one 16-byte vector reduction per lane and corner
//     (red.global.add.v4.f32, SASS REDG.E.ADD.F32x4) instead of 4 scalar atomics, skipped for
//     corners with zero weight.  Shared-memory float atomics are CAS loops on sm_100a
//     (ATOMS.CAST.SPIN), which is why the accumulation is not staged in shared memory.
//   * grad_sampling_loc / grad_attn_weight are written exactly once per tap with plain coalesced
//     stores -- no zero-fill pass, no atomics (out-of-range taps write 0 like the reference's zeros()).
#pragma once
#include "msda_common.cuh"

namespace devis {

template <class SlotSrc>
struct BwdArgs {
    const void *value;
    const void *grad_out;
    float *grad_value;  // fp32 accumulation target (also for bf16 value); nullptr = skip
    Segment seg[2];
    int n_seg;
    int n_slots_total;
    SlotSrc src;
    OpDims d;
    const int *q_perm;
};

// d[t][c]: partial sums of tap t, corner c held by this lane.  On return r[c] is the group total of
// corner c for tap j (j = this lane's index in the group).
template <int LPG>
__device__ __forceinline__ void reduce_scatter_taps(float (&d)[LPG][4], int j, float (&r)[4])
{
#pragma unroll
    for (int half = LPG / 2; half >= 1; half >>= 1) {
        const bool upper = (j & half) != 0;
#pragma unroll
        for (int t = 0; t < half; ++t) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float send = upper ? d[t][c] : d[t + half][c];
                const float keep = upper ? d[t + half][c] : d[t][c];
                d[t][c] = keep + __shfl_xor_sync(0xffffffffu, send, half, LPG);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) r[c] = d[0][c];
}

template <bool BF16, int LPG, int QPG, class SlotSrc>
__global__ void __launch_bounds__(256) msda_bwd_kernel(const BwdArgs<SlotSrc> a)
{
    extern __shared__ int4 s_slot[];
    const int outer = blockIdx.y;
    build_slots(s_slot, a.src, a.d, outer, a.n_slots_total);

    const int M = a.d.M, Lq = a.d.Lq;
    const int j = threadIdx.x % LPG;
    const int grp = threadIdx.x / LPG;
    const int QC = blockDim.x / LPG;
    const int qchunk = blockIdx.x / M, m = blockIdx.x - qchunk * M;

    int q[QPG];
    bool qlive[QPG];
    float4 go[QPG];
#pragma unroll
    for (int i = 0; i < QPG; ++i) {
        const int qi = (qchunk * QPG + i) * QC + grp;
        qlive[i] = qi < Lq;
        q[i] = qlive[i] ? (a.q_perm ? a.q_perm[qi] : qi) : 0;
        go[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (qlive[i]) {
            const size_t row = ((size_t)outer * Lq + q[i]) * M + m;
            go[i] = BF16 ? ldg_bf16x4(reinterpret_cast<const uint2 *>(a.grad_out) + row * LPG + j)
                         : ldg_f4(reinterpret_cast<const float4 *>(a.grad_out) + row * LPG + j);
        }
    }

    const unsigned ps = (unsigned)(M * LPG);
    const float4 *vb32 = reinterpret_cast<const float4 *>(a.value) + m * LPG + j;
    const uint2 *vb16 = reinterpret_cast<const uint2 *>(a.value) + m * LPG + j;
    float *gvb = a.grad_value ? a.grad_value + 4 * (m * LPG + j) : nullptr;

    int slot_base = 0;
    for (int sg = 0; sg < a.n_seg; ++sg) {
        const int P = a.seg[sg].P, K = a.seg[sg].n_slots * P;
        const float *loc = reinterpret_cast<const float *>(a.seg[sg].loc);
        const float *aw = reinterpret_cast<const float *>(a.seg[sg].aw);
        float *gloc = reinterpret_cast<float *>(a.seg[sg].grad_loc);
        float *gaw = reinterpret_cast<float *>(a.seg[sg].grad_aw);
        for (int k0 = 0; k0 < K; k0 += LPG) {
            const int k = k0 + j;
            const bool klive = k < K;
            const int4 sl = s_slot[slot_base + (klive ? k / P : 0)];
#pragma unroll
            for (int i = 0; i < QPG; ++i) {
                const size_t row = ((size_t)outer * Lq + q[i]) * M + m;
                const bool live = klive && qlive[i];
                float2 xy = make_float2(0.f, 0.f);
                float w = 0.f;
                if (live) {
                    xy = __ldg(reinterpret_cast<const float2 *>(loc + row * K * 2) + k);
                    w = __ldg(aw + row * K + k);
                }
                const TapGeom g = tap_geometry(xy.x, xy.y, sl, live);
                const float b00 = g.hh * g.hw, b01 = g.hh * g.lw, b10 = g.lh * g.hw, b11 = g.lh * g.lw;
                const float w00 = (g.ok & 1u) ? w * b00 : 0.f;
                const float w01 = (g.ok & 2u) ? w * b01 : 0.f;
                const float w10 = (g.ok & 4u) ? w * b10 : 0.f;
                const float w11 = (g.ok & 8u) ? w * b11 : 0.f;
                const unsigned rT = (unsigned)g.rowT | ((unsigned)g.dcol << 31);
                const unsigned rB = (unsigned)g.rowB;

                float dsum[LPG][4];
#pragma unroll
                for (int jj = 0; jj < LPG; ++jj) {
                    const unsigned t = __shfl_sync(0xffffffffu, rT, jj, LPG);
                    const unsigned b = __shfl_sync(0xffffffffu, rB, jj, LPG);
                    const float c00 = __shfl_sync(0xffffffffu, w00, jj, LPG);
                    const float c01 = __shfl_sync(0xffffffffu, w01, jj, LPG);
                    const float c10 = __shfl_sync(0xffffffffu, w10, jj, LPG);
                    const float c11 = __shfl_sync(0xffffffffu, w11, jj, LPG);
                    const unsigned dc = (t >> 31) ? ps : 0u;
                    const size_t oT = (size_t)(t & 0x7fffffffu) * ps, oB = (size_t)b * ps;
                    float4 v00, v01, v10, v11;
                    if (BF16) {
                        v00 = ldg_bf16x4(vb16 + oT);
                        v01 = ldg_bf16x4(vb16 + oT + dc);
                        v10 = ldg_bf16x4(vb16 + oB);
                        v11 = ldg_bf16x4(vb16 + oB + dc);
                    } else {
                        v00 = ldg_f4(vb32 + oT);
                        v01 = ldg_f4(vb32 + oT + dc);
                        v10 = ldg_f4(vb32 + oB);
                        v11 = ldg_f4(vb32 + oB + dc);
                    }
                    const float4 gg = go[i];
                    dsum[jj][0] = v00.x * gg.x + v00.y * gg.y + v00.z * gg.z + v00.w * gg.w;
                    dsum[jj][1] = v01.x * gg.x + v01.y * gg.y + v01.z * gg.z + v01.w * gg.w;
                    dsum[jj][2] = v10.x * gg.x + v10.y * gg.y + v10.z * gg.z + v10.w * gg.w;
                    dsum[jj][3] = v11.x * gg.x + v11.y * gg.y + v11.z * gg.z + v11.w * gg.w;
                    if (gvb) {
                        if (c00 != 0.f) red_add_f4(gvb + 4 * oT, c00 * gg.x, c00 * gg.y, c00 * gg.z, c00 * gg.w);
                        if (c01 != 0.f) red_add_f4(gvb + 4 * (oT + dc), c01 * gg.x, c01 * gg.y, c01 * gg.z, c01 * gg.w);
                        if (c10 != 0.f) red_add_f4(gvb + 4 * oB, c10 * gg.x, c10 * gg.y, c10 * gg.z, c10 * gg.w);
                        if (c11 != 0.f) red_add_f4(gvb + 4 * (oB + dc), c11 * gg.x, c11 * gg.y, c11 * gg.z, c11 * gg.w);
                    }
                }

                float A[4];
                reduce_scatter_taps<LPG>(dsum, j, A);
                if (live) {
                    const float a00 = (g.ok & 1u) ? A[0] : 0.f, a01 = (g.ok & 2u) ? A[1] : 0.f;
                    const float a10 = (g.ok & 4u) ? A[2] : 0.f, a11 = (g.ok & 8u) ? A[3] : 0.f;
                    const float gx = g.hh * (a01 - a00) + g.lh * (a11 - a10);
                    const float gy = g.hw * (a10 - a00) + g.lw * (a11 - a01);
                    const bool hit = g.ok != 0u;  // out-of-range taps write the reference's zero fill
                    gaw[row * K + k] = hit ? b00 * a00 + b01 * a01 + b10 * a10 + b11 * a11 : 0.f;
                    reinterpret_cast<float2 *>(gloc + row * K * 2)[k] =
                        hit ? make_float2((float)sl.y * gx * w, (float)sl.x * gy * w) : make_float2(0.f, 0.f);
                }
            }
        }
        slot_base += a.seg[sg].n_slots;
    }
}

}  // namespace devis
