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
//     attention weight folded in, and publishes a 32-byte record (4 byte offsets, 4 weight factors)
//     in a per-warp shared-memory exchange; the group then walks the LPG records with two broadcast
//     LDS.128 each (TapExchange in msda_common.cuh; shuffles were measured to cost 3x the L1 wavefronts);
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
    using X = TapExchange<LPG>;
    extern __shared__ int4 s_slot[];
    const int outer = blockIdx.y;
    build_slots(s_slot, a.src, a.d, outer, a.n_slots_total);
    float *xbuf = reinterpret_cast<float *>(s_slot + a.n_slots_total) + (threadIdx.x >> 5) * (2 * X::kWordsPerWarpBuf);

    const int M = a.d.M, Lq = a.d.Lq;
    const int j = threadIdx.x % LPG;          // channel quad owned by this lane == tap it prepares
    const int g = (threadIdx.x & 31) / LPG;   // group within the warp
    const int grp = threadIdx.x / LPG;        // group within the CTA
    const int QC = blockDim.x / LPG;          // (query, head) pairs in flight per CTA
    const int qchunk = blockIdx.x / M, m = blockIdx.x - qchunk * M;

    int q[QPG];
    bool qlive[QPG];
#pragma unroll
    for (int i = 0; i < QPG; ++i) {
        const int qi = (qchunk * QPG + i) * QC + grp;
        qlive[i] = qi < Lq;
        q[i] = qlive[i] ? (a.q_perm ? a.q_perm[qi] : qi) : 0;
    }

    // value row = M*LPG channel quads; this lane reads quad (m*LPG + j) of every row it touches
    constexpr unsigned kQuadBytes = BF16 ? 8u : 16u;
    const unsigned rowbytes = (unsigned)(M * LPG) * kQuadBytes;
    const char *vbase = reinterpret_cast<const char *>(a.value) + (size_t)(m * LPG + j) * kQuadBytes;
    // keep the per-lane base as ONE 64-bit register pair: each corner address is then base + u32 offset
    // (IADD3 + IADD3.X) instead of a re-derived IMAD.WIDE chain (5 instructions per address in round 1a)
    asm volatile("" : "+l"(vbase));

    float4 acc[QPG];
#pragma unroll
    for (int i = 0; i < QPG; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    int slot_base = 0, parity = 0;
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
                const TapGeom t = tap_geometry(xy.x, xy.y, sl, live);
                float *buf = xbuf + parity * X::kWordsPerWarpBuf;
                parity ^= 1;
                X::publish(buf, j, g, t, w, rowbytes);
                __syncwarp();
#pragma unroll
                for (int jj = 0; jj < LPG; ++jj) {
                    uint4 off;
                    float4 c;
                    X::fetch(buf, jj, g, off, c);
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
                    acc[i].x = fmaf(c.w, v11.x, fmaf(c.z, v10.x, fmaf(c.y, v01.x, fmaf(c.x, v00.x, acc[i].x))));
                    acc[i].y = fmaf(c.w, v11.y, fmaf(c.z, v10.y, fmaf(c.y, v01.y, fmaf(c.x, v00.y, acc[i].y))));
                    acc[i].z = fmaf(c.w, v11.z, fmaf(c.z, v10.z, fmaf(c.y, v01.z, fmaf(c.x, v00.z, acc[i].z))));
                    acc[i].w = fmaf(c.w, v11.w, fmaf(c.z, v10.w, fmaf(c.y, v01.w, fmaf(c.x, v00.w, acc[i].w))));
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
