// msda_fwd.cuh -- forward of multi-scale deformable attention, grouped-lane kernel for sm_100a.
//
// Replaces ms_deformable_im2col_gpu_kernel (cuda/ms_deform_im2col_cuda.cuh:237-299), which maps one
// thread to one output channel and therefore recomputes each tap's floor/weights D times and gathers
// with 4-byte loads.  Here:
//   * a GROUP of LPG lanes owns one (query, head); each lane owns 4 consecutive channels, so a value
//     row of one head (D = 4*LPG channels) is fetched by the group as one 16*LPG-byte line
//     (128 B for D=32 fp32): 4 corner loads per tap are 4 fully used L1 wavefronts per warp;
//   * tap geometry is computed ONCE per tap: lane j of the group reads location / weight of tap
//     k0+j (coalesced over the group), does the floor / range test / bilinear weights with the
//     attention weight folded in, and the group then walks the LPG taps broadcasting the 6 words
//     (2 row indices, 4 weights) with width-LPG shuffles;
//   * a CTA works on ONE head and QC*QPG queries, and walks slots in order, so that its working set
//     at any time is one (frame, level, head) neighbourhood of value -- the L1/L2 locality that the
//     whole-clip op relies on; QPG queries per group are interleaved (accumulators in registers).
//
// The same kernel serves the per-call op (SlotSrc = DeviceLevels, one segment) and the whole-clip
// temporal op (SlotSrc = ClipTable, two segments, value indexed through the frame table).
#pragma once
#include "msda_common.cuh"

namespace devis {

template <class SlotSrc>
struct FwdArgs {
    const void *value;
    void *out;
    Segment seg[2];
    int n_seg;
    int n_slots_total;
    SlotSrc src;
    OpDims d;
    const int *q_perm;  // optional query visiting order (length Lq), nullptr = identity
};

template <bool BF16, int LPG, int QPG, class SlotSrc>
__global__ void __launch_bounds__(256) msda_fwd_kernel(const FwdArgs<SlotSrc> a)
{
    extern __shared__ int4 s_slot[];
    const int outer = blockIdx.y;
    build_slots(s_slot, a.src, a.d, outer, a.n_slots_total);

    const int M = a.d.M, Lq = a.d.Lq;
    const int j = threadIdx.x % LPG;          // channel-quad owned by this lane == tap it prepares
    const int grp = threadIdx.x / LPG;
    const int QC = blockDim.x / LPG;          // groups (queries in flight) per CTA
    const int qchunk = blockIdx.x / M, m = blockIdx.x - qchunk * M;

    int q[QPG];
    bool qlive[QPG];
#pragma unroll
    for (int i = 0; i < QPG; ++i) {
        const int qi = (qchunk * QPG + i) * QC + grp;
        qlive[i] = qi < Lq;
        q[i] = qlive[i] ? (a.q_perm ? a.q_perm[qi] : qi) : 0;
    }

    // value viewed as 16-byte (fp32) / 8-byte (bf16) channel quads: row stride M*LPG quads
    const unsigned ps = (unsigned)(M * LPG);
    const float4 *vb32 = reinterpret_cast<const float4 *>(a.value) + m * LPG + j;
    const uint2 *vb16 = reinterpret_cast<const uint2 *>(a.value) + m * LPG + j;

    float4 acc[QPG];
#pragma unroll
    for (int i = 0; i < QPG; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    int slot_base = 0;
    for (int sg = 0; sg < a.n_seg; ++sg) {
        const int P = a.seg[sg].P, K = a.seg[sg].n_slots * P;
        const float *loc = reinterpret_cast<const float *>(a.seg[sg].loc);
        const float *aw = reinterpret_cast<const float *>(a.seg[sg].aw);
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
                const float w00 = (g.ok & 1u) ? w * g.hh * g.hw : 0.f;
                const float w01 = (g.ok & 2u) ? w * g.hh * g.lw : 0.f;
                const float w10 = (g.ok & 4u) ? w * g.lh * g.hw : 0.f;
                const float w11 = (g.ok & 8u) ? w * g.lh * g.lw : 0.f;
                const unsigned rT = (unsigned)g.rowT | ((unsigned)g.dcol << 31);
                const unsigned rB = (unsigned)g.rowB;
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
                    acc[i].x += c00 * v00.x + c01 * v01.x + c10 * v10.x + c11 * v11.x;
                    acc[i].y += c00 * v00.y + c01 * v01.y + c10 * v10.y + c11 * v11.y;
                    acc[i].z += c00 * v00.z + c01 * v01.z + c10 * v10.z + c11 * v11.z;
                    acc[i].w += c00 * v00.w + c01 * v01.w + c10 * v10.w + c11 * v11.w;
                }
            }
        }
        slot_base += a.seg[sg].n_slots;
    }

#pragma unroll
    for (int i = 0; i < QPG; ++i) {
        if (!qlive[i]) continue;
        const size_t row = ((size_t)outer * Lq + q[i]) * M + m;
        if (BF16)
            reinterpret_cast<uint2 *>(a.out)[row * LPG + j] = pack_bf16x4(acc[i]);
        else
            reinterpret_cast<float4 *>(a.out)[row * LPG + j] = acc[i];
    }
}

}  // namespace devis
