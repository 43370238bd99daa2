// msda_common.cuh -- shared definitions for the sm_100a multi-scale deformable attention kernels.
//
// Vocabulary (follows the reference, modules/ms_deform_attn.py and cuda/ms_deform_im2col_cuda.cuh):
//   value row   one pixel of one level of one batch item / frame: M heads x D channels, contiguous
//   slot        one (frame, level) map a query can sample: {H, W, first value row}.  For the per-call op a
//               slot is a level (ms_deform_attn_cuda.cu:67-68); for the whole-clip temporal op slots
//               [0,L) are the query frame's own levels and slot L + j*L + l is level l of temporal
//               frame j (ms_deform_attn.py:232-238: temporal "levels" are frame-major, level-minor)
//   segment     a (sampling_loc, attn_weight) tensor pair covering a contiguous run of slots with a fixed
//               number of points per slot: one segment per call for the plain op, two (current, temporal)
//               for the whole-clip op
//   tap         one sampling point of one (query, head): a bilinear read of 4 value rows
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace devis {

constexpr int kMaxLevels = 16;        // host-side table of the whole-clip op
constexpr int kMaxFrameTable = 2048;  // T * t_window entries (uint8 each)
constexpr int kMaxSlots = 1024;       // slots per query (dynamic smem: 16 B each)

struct Segment {
    const void *loc;   // (outer, Lq, M, n_slots, P, 2)
    const void *aw;    // (outer, Lq, M, n_slots, P)
    void *grad_loc;    // backward only
    void *grad_aw;     // backward only
    int n_slots;
    int P;
};

// Slot source of the per-call op: the reference's device-resident int64 tensors.
struct DeviceLevels {
    const int64_t *shapes;  // (L,2) rows (H,W)
    const int64_t *lsi;     // (L)
};

// Slot source of the whole-clip op: host tables passed by value in the kernel parameters.
struct ClipTable {
    int H[kMaxLevels];
    int W[kMaxLevels];
    int lsi[kMaxLevels];
    int L;
    int Wt;
    uint8_t frame[kMaxFrameTable];  // [t * Wt + j]
};

struct OpDims {
    int outer;  // batch (plain op) or query frames (whole-clip op)
    int S;      // value rows per outer item
    int M, D;
    int Lq;
};

// ---------------------------------------------------------------------------------------------
// slot table in shared memory: int4 {H, W, first value row (absolute, outer offset included), W*M}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void build_slots(int4 *s_slot, const DeviceLevels &lv, const OpDims &d, int outer,
                                            int n_slots)
{
    for (int s = threadIdx.x; s < n_slots; s += blockDim.x) {
        const int H = (int)lv.shapes[2 * s], W = (int)lv.shapes[2 * s + 1];
        s_slot[s] = make_int4(H, W, outer * d.S + (int)lv.lsi[s], W * d.M);
    }
    __syncthreads();
}

__device__ __forceinline__ void build_slots(int4 *s_slot, const ClipTable &tb, const OpDims &d, int outer,
                                            int n_slots)
{
    for (int s = threadIdx.x; s < n_slots; s += blockDim.x) {
        int frame, l;
        if (s < tb.L) {
            frame = outer;
            l = s;
        } else {
            const int j = (s - tb.L) / tb.L;
            l = (s - tb.L) - j * tb.L;
            frame = tb.frame[outer * tb.Wt + j];
        }
        s_slot[s] = make_int4(tb.H[l], tb.W[l], frame * d.S + tb.lsi[l], tb.W[l] * d.M);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// One tap's geometry, computed by ONE lane and then shared with the lanes that own the channels.
// Mirrors cuda/ms_deform_im2col_cuda.cuh:281-288 (pixel coordinate and range test) and :38-52,80
// (floor cell, bilinear weights); rounding sequence of the coordinate is the reference's:
// round(round(loc*size) - 0.5), no FMA contraction.
// ---------------------------------------------------------------------------------------------
struct TapGeom {
    float lh, lw, hh, hw;  // fractional parts and their complements
    int rowT, rowB;        // clamped value rows of the top-left / bottom-left corner
    int dcol;              // 0 or 1: clamped column step to the right corners
    unsigned ok;           // bit0 TL, bit1 TR, bit2 BL, bit3 BR corner inside the map; 0 if tap out of range
};

__device__ __forceinline__ TapGeom tap_geometry(float x, float y, const int4 slot, bool live)
{
    TapGeom g;
    const int H = slot.x, W = slot.y;
    const float h_im = __fadd_rn(__fmul_rn(y, (float)H), -0.5f);
    const float w_im = __fadd_rn(__fmul_rn(x, (float)W), -0.5f);
    const bool inb = live && h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W;
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int h0 = inb ? (int)hf : 0, w0 = inb ? (int)wf : 0;
    g.lh = h_im - hf;
    g.lw = w_im - wf;
    g.hh = 1.f - g.lh;
    g.hw = 1.f - g.lw;
    const bool t_ok = h0 >= 0, b_ok = h0 + 1 <= H - 1, l_ok = w0 >= 0, r_ok = w0 + 1 <= W - 1;
    g.ok = inb ? ((t_ok && l_ok) | ((t_ok && r_ok) << 1) | ((b_ok && l_ok) << 2) | ((b_ok && r_ok) << 3)) : 0u;
    const int h0c = max(h0, 0), h1c = min(h0 + 1, H - 1), w0c = max(w0, 0), w1c = min(w0 + 1, W - 1);
    g.rowT = slot.z + h0c * W + w0c;
    g.rowB = slot.z + h1c * W + w0c;
    g.dcol = w1c - w0c;
    return g;
}

// 16-byte read-only loads / vector reductions --------------------------------------------------
__device__ __forceinline__ float4 ldg_f4(const float4 *p) { return __ldg(p); }

__device__ __forceinline__ float4 ldg_bf16x4(const uint2 *p)
{
    const uint2 r = __ldg(p);
    float4 o;
    o.x = __uint_as_float(r.x << 16);
    o.y = __uint_as_float(r.x & 0xffff0000u);
    o.z = __uint_as_float(r.y << 16);
    o.w = __uint_as_float(r.y & 0xffff0000u);
    return o;
}

__device__ __forceinline__ void red_add_f4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ uint2 pack_bf16x4(float4 v)
{
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<const unsigned *>(&lo);
    r.y = *reinterpret_cast<const unsigned *>(&hi);
    return r;
}

}  // namespace devis
