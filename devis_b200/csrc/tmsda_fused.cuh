// tmsda_fused.cuh -- whole-clip temporal attention with the PROLOGUE fused in.
//
// The reference turns the Linear outputs into op operands with a chain of elementwise kernels
// (modules/ms_deform_attn.py:232-264 and :437-452): view/flatten, cat of the current and temporal logits, one
// softmax over all K = L*Pc + Wt*L*Pt taps, two slices made contiguous, `ref + off / (W, H)` for the current and
// for the temporal taps -- and autograd runs all of it backwards again.  At the DeVIS encoder shape that is
// 178 MB of locations + 89 MB of weights written and re-read per layer-clip and direction.
// Here the kernels read the RAW Linear outputs:
//     off_curr (T,Lq,M,L,Pc,2), logit_curr (T,Lq,M,L*Pc), off_temporal (T,Lq,M,Wt*L,Pt,2), logit_temporal (T,Lq,M,Wt*L*Pt)
// plus the reference points ref (T,Lq,L,2), and the lane that prepares a tap does
//     loc = ref[level] + off / (W_level, H_level)        (current taps; __fdiv_rn/__fadd_rn == torch's div, add)
//     loc = ref[level 0] + off / (W_level, H_level)      (temporal taps start from the level-0 point, :447)
//     w   = exp(logit - rowmax) / rowsum                 (rowmax, rowsum over all K taps of the (query, head))
// The backward returns d/d(off) = grad_loc / (W, H) and d/d(logit) = w * (grad_w - sum_k w_k grad_w_k) directly.
// Encoder form only (2-d reference points); the decoder's taps are 1000x fewer and keep the unfused path.
// Work split, tap exchange, reductions: identical to msda_fwd.cuh / msda_bwd.cuh (LPG = 8, D = 32).
#pragma once
#include "msda_bwd.cuh"
#include "msda_common.cuh"
#include "msda_fwd.cuh"

namespace devis {

struct FusedArgs {
    const void *value;
    const float *ref;          // (T, Lq, L, 2)
    const float *off[2];       // raw sampling offsets, current / temporal
    const float *logit[2];     // raw attention logits, current / temporal
    int n_slots[2];
    int P[2];
    int n_seg;
    ClipTable src;
    OpDims d;
    const int *q_perm;
    // forward
    void *out;
    // backward
    const void *grad_out;
    void *grad_value;          // fp32 (bf16 under HALF_ACC)
    float *grad_off[2];
    float *grad_logit[2];
};

// softmax statistics of one (query, head) row over both segments, computed by the 8 lanes of its group
__device__ __forceinline__ void row_softmax_stats(const FusedArgs &a, size_t row, int j, bool live, float &rmax,
                                                  float &rinv)
{
    float mx = -INFINITY;
    for (int sg = 0; sg < a.n_seg; ++sg) {
        const int K = a.n_slots[sg] * a.P[sg];
        const float *lg = a.logit[sg] + row * K;
        for (int k = j; k < K; k += 8)
            if (live) mx = fmaxf(mx, __ldg(lg + k));
    }
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o, 8));
    float sum = 0.f;
    for (int sg = 0; sg < a.n_seg; ++sg) {
        const int K = a.n_slots[sg] * a.P[sg];
        const float *lg = a.logit[sg] + row * K;
        for (int k = j; k < K; k += 8)
            if (live) sum += expf(__ldg(lg + k) - mx);
    }
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o, 8);
    rmax = live ? mx : 0.f;
    rinv = live ? 1.f / sum : 0.f;
}

// location and weight of tap k of segment sg for (query row): the fused prologue
__device__ __forceinline__ void fused_tap_operands(const FusedArgs &a, int sg, size_t row, size_t qrow, int k, int K,
                                                   const int4 sl, int level, float rmax, float rinv, float &x,
                                                   float &y, float &w)
{
    const float2 off = __ldg(reinterpret_cast<const float2 *>(a.off[sg] + row * K * 2) + k);
    const float2 rf = __ldg(reinterpret_cast<const float2 *>(a.ref + (qrow * a.src.L + (sg == 0 ? level : 0)) * 2));
    x = __fadd_rn(rf.x, __fdiv_rn(off.x, (float)sl.y));
    y = __fadd_rn(rf.y, __fdiv_rn(off.y, (float)sl.x));
    w = expf(__ldg(a.logit[sg] + row * K + k) - rmax) * rinv;
}

template <bool BF16>
__global__ void __launch_bounds__(256, 3) tmsda_fused_fwd_kernel(const FusedArgs a)
{
    constexpr int LPG = 8;
    extern __shared__ int4 s_slot[];
    const int outer = blockIdx.y, n_slots_total = a.n_slots[0] + (a.n_seg > 1 ? a.n_slots[1] : 0);
    build_slots(s_slot, a.src, a.d, outer, n_slots_total);
    float *xbuf = reinterpret_cast<float *>(s_slot + n_slots_total) + (threadIdx.x >> 5) * (2 * Tap16x8::kWordsPerWarpBuf);

    const int M = a.d.M, Lq = a.d.Lq;
    const int j = threadIdx.x % LPG, g = (threadIdx.x & 31) / LPG, grp = threadIdx.x / LPG, QC = blockDim.x / LPG;
    const int qchunk = blockIdx.x / M, m = blockIdx.x - qchunk * M;
    const int qi = qchunk * QC + grp;
    const bool qlive = qi < Lq;
    const int q = qlive ? (a.q_perm ? a.q_perm[qi] : qi) : 0;
    const size_t qrow = (size_t)outer * Lq + q, row = qrow * M + m;

    constexpr unsigned kQuadBytes = BF16 ? 8u : 16u;
    const unsigned rowbytes = (unsigned)(M * LPG) * kQuadBytes;
    const char *vbase = reinterpret_cast<const char *>(a.value) + (size_t)(m * LPG + j) * kQuadBytes;
    asm volatile("" : "+l"(vbase));

    float rmax, rinv;
    row_softmax_stats(a, row, j, qlive, rmax, rinv);

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int slot_base = 0, parity = 0;
    for (int sg = 0; sg < a.n_seg; ++sg) {
        const int P = a.P[sg], K = a.n_slots[sg] * P;             // P % 4 == 0 (checked by the launcher)
        for (int k0 = 0; k0 < K; k0 += LPG) {
            const int k = k0 + j;
            const bool live = k < K && qlive;
            const int ls = k < K ? k / P : 0;                      // slot within the segment
            const int4 sl = s_slot[slot_base + ls];
            const unsigned my_pitch = (unsigned)sl.y * rowbytes;
            const unsigned pitch_lo = __shfl_sync(0xffffffffu, my_pitch, 0, 8);
            const unsigned pitch_hi = __shfl_sync(0xffffffffu, my_pitch, 4, 8);
            float x = 0.f, y = 0.f, w = 0.f;
            if (live) fused_tap_operands(a, sg, row, qrow, k, K, sl, ls % a.src.L, rmax, rinv, x, y, w);
            const TapGeom t = tap_geometry(x, y, sl, live);
            float *buf = xbuf + parity * Tap16x8::kWordsPerWarpBuf;
            parity ^= 1;
            *reinterpret_cast<uint4 *>(buf + Tap16x8::word(j, g)) = make_tap16(t, w, rowbytes);
            __syncwarp();
            consume_tap16x8<BF16>(buf, g, rowbytes, pitch_lo, pitch_hi, vbase, acc);
        }
        slot_base += a.n_slots[sg];
    }
    if (qlive) {
        if (BF16)
            reinterpret_cast<uint2 *>(a.out)[row * LPG + j] = pack_bf16x4(acc);
        else
            reinterpret_cast<float4 *>(a.out)[row * LPG + j] = acc;
    }
}

// HALF_ACC: bf16 grad_value with packed bf16 reductions (see msda_bwd_kernel)
template <bool BF16, bool HALF_ACC = false>
__global__ void __launch_bounds__(256, DEVIS_BWD_MIN_BLOCKS) tmsda_fused_bwd_kernel(const FusedArgs a)
{
    constexpr int LPG = 8;
    using X = TapExchange<LPG>;
    extern __shared__ int4 s_slot[];
    const int outer = blockIdx.y, n_slots_total = a.n_slots[0] + (a.n_seg > 1 ? a.n_slots[1] : 0);
    build_slots(s_slot, a.src, a.d, outer, n_slots_total);
    float *xbuf = reinterpret_cast<float *>(s_slot + n_slots_total) + (threadIdx.x >> 5) * (2 * X::kWordsPerWarpBuf);

    const int M = a.d.M, Lq = a.d.Lq;
    const int j = threadIdx.x % LPG, g = (threadIdx.x & 31) / LPG, grp = threadIdx.x / LPG, QC = blockDim.x / LPG;
    const int qchunk = blockIdx.x / M, m = blockIdx.x - qchunk * M;
    const int qi = qchunk * QC + grp;
    const bool qlive = qi < Lq;
    const int q = qlive ? (a.q_perm ? a.q_perm[qi] : qi) : 0;
    const size_t qrow = (size_t)outer * Lq + q, row = qrow * M + m;

    float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (qlive)
        gg = BF16 ? ldg_bf16x4(reinterpret_cast<const uint2 *>(a.grad_out) + row * LPG + j)
                  : ldg_f4(reinterpret_cast<const float4 *>(a.grad_out) + row * LPG + j);

    constexpr unsigned kQuadBytes = BF16 ? 8u : 16u;
    const unsigned rowbytes = (unsigned)(M * LPG) * kQuadBytes;
    const char *vbase = reinterpret_cast<const char *>(a.value) + (size_t)(m * LPG + j) * kQuadBytes;
    asm volatile("" : "+l"(vbase));
    static_assert(!HALF_ACC || BF16, "bf16 accumulation needs bf16 value");
    char *gvb = a.grad_value ? reinterpret_cast<char *>(a.grad_value) + (size_t)(m * LPG + j) * (HALF_ACC ? 8u : 16u) : nullptr;
    constexpr unsigned kGvShift = (BF16 && !HALF_ACC) ? 1u : 0u;

    float rmax, rinv;
    row_softmax_stats(a, row, j, qlive, rmax, rinv);

    float dotp = 0.f;   // this lane's share of sum_k w_k * d(out.grad_out)/d(w_k)
    int slot_base = 0, parity = 0;
    for (int sg = 0; sg < a.n_seg; ++sg) {
        const int P = a.P[sg], K = a.n_slots[sg] * P;
        float *goff = a.grad_off[sg] + row * K * 2;
        float *glog = a.grad_logit[sg] + row * K;
        for (int k0 = 0; k0 < K; k0 += LPG) {
            const int k = k0 + j;
            const bool live = k < K && qlive;
            const int ls = live ? k / P : 0;
            const int4 sl = s_slot[slot_base + ls];
            float x = 0.f, y = 0.f, w = 0.f;
            if (live) fused_tap_operands(a, sg, row, qrow, k, K, sl, ls % a.src.L, rmax, rinv, x, y, w);
            const TapGeom t = tap_geometry(x, y, sl, live);
            float *buf = xbuf + parity * X::kWordsPerWarpBuf;
            parity ^= 1;
            X::publish(buf, j, g, t, w, rowbytes);
            __syncwarp();

            float dsum[LPG][4];
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
                dsum[jj][0] = fmaf(v00.w, gg.w, fmaf(v00.z, gg.z, fmaf(v00.y, gg.y, v00.x * gg.x)));
                dsum[jj][1] = fmaf(v01.w, gg.w, fmaf(v01.z, gg.z, fmaf(v01.y, gg.y, v01.x * gg.x)));
                dsum[jj][2] = fmaf(v10.w, gg.w, fmaf(v10.z, gg.z, fmaf(v10.y, gg.y, v10.x * gg.x)));
                dsum[jj][3] = fmaf(v11.w, gg.w, fmaf(v11.z, gg.z, fmaf(v11.y, gg.y, v11.x * gg.x)));
                if (gvb) {
                    if (c.x != 0.f) red_add_quad<HALF_ACC>(gvb + ((size_t)off.x << kGvShift), c.x * gg.x, c.x * gg.y, c.x * gg.z, c.x * gg.w);
                    if (c.y != 0.f) red_add_quad<HALF_ACC>(gvb + ((size_t)off.y << kGvShift), c.y * gg.x, c.y * gg.y, c.y * gg.z, c.y * gg.w);
                    if (c.z != 0.f) red_add_quad<HALF_ACC>(gvb + ((size_t)off.z << kGvShift), c.z * gg.x, c.z * gg.y, c.z * gg.z, c.z * gg.w);
                    if (c.w != 0.f) red_add_quad<HALF_ACC>(gvb + ((size_t)off.w << kGvShift), c.w * gg.x, c.w * gg.y, c.w * gg.z, c.w * gg.w);
                }
            }
            float A[4];
            reduce_scatter_taps<LPG>(dsum, j, A);
            if (live) {
                const bool hit = t.ok != 0u;
                const float hh = (t.ok & 1u) ? t.hh : 0.f, lh = (t.ok & 2u) ? t.lh : 0.f;
                const float hw = (t.ok & 4u) ? t.hw : 0.f, lw = (t.ok & 8u) ? t.lw : 0.f;
                const float l_in = (t.ok & 4u) ? 1.f : 0.f, r_in = (t.ok & 8u) ? 1.f : 0.f;
                const float t_in = (t.ok & 1u) ? 1.f : 0.f, b_in = (t.ok & 2u) ? 1.f : 0.f;
                const float val = hit ? hh * (hw * A[0] + lw * A[1]) + lh * (hw * A[2] + lw * A[3]) : 0.f;
                const float gx = hh * (r_in * A[1] - l_in * A[0]) + lh * (r_in * A[3] - l_in * A[2]);
                const float gy = hw * (b_in * A[2] - t_in * A[0]) + lw * (b_in * A[3] - t_in * A[1]);
                // d/d(loc) as in msda_bwd.cuh, then the chain rule of loc = ref + off / (W, H)
                const float glx = hit ? (float)sl.y * gx * w : 0.f, gly = hit ? (float)sl.x * gy * w : 0.f;
                reinterpret_cast<float2 *>(goff)[k] = make_float2(__fdiv_rn(glx, (float)sl.y), __fdiv_rn(gly, (float)sl.x));
                glog[k] = val;            // parked; finished below once the row's  sum_k w_k val_k  is known
                dotp = fmaf(w, val, dotp);
            }
        }
        slot_base += a.n_slots[sg];
    }
    // softmax backward:  d/d(logit_k) = w_k * (val_k - sum_j w_j val_j)
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) dotp += __shfl_xor_sync(0xffffffffu, dotp, o, 8);
    if (qlive) {
        for (int sg = 0; sg < a.n_seg; ++sg) {
            const int K = a.n_slots[sg] * a.P[sg];
            float *glog = a.grad_logit[sg] + row * K;
            const float *lg = a.logit[sg] + row * K;
            for (int k = j; k < K; k += 8) {
                const float w = expf(__ldg(lg + k) - rmax) * rinv;
                glog[k] = w * (glog[k] - dotp);
            }
        }
    }
}

}  // namespace devis
