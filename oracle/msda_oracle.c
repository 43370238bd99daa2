/*
 * oracle/msda_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, optional OpenMP over heads) of the algorithm that the
 * reference's CUDA op implements for multi-scale deformable attention.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file; the product path (devis_b200/) never does.
 *
 * What it follows (paths relative to /root/reference/src/models/ops/src/cuda):
 *   forward   ms_deform_im2col_cuda.cuh:237-299  (ms_deformable_im2col_gpu_kernel)
 *             ms_deform_im2col_cuda.cuh:33-84    (ms_deform_attn_im2col_bilinear)
 *   backward  ms_deform_im2col_cuda.cuh:301-403  (col2im ... blocksize_aware_reduce_v1)
 *             ms_deform_im2col_cuda.cuh:87-159   (ms_deform_attn_col2im_bilinear)
 *   host      ms_deform_attn_cuda.cu:20-80,83-153 (zero-initialised outputs; the
 *             im2col_step batching loop is a pure chunking of the batch axis and
 *             therefore has no numerical effect -- it is not restated)
 *
 * Semantics pinned here:
 *   - locations are (x, y) in [0,1]; pixel coordinate = loc * (W, H) - 0.5
 *   - a sample contributes only if  -1 < h < H  and  -1 < w < W
 *   - bilinear interpolation with zero padding for corners outside the map
 *   - value layout (N, S, M, D); level l occupies rows [lsi[l], lsi[l] + H_l*W_l)
 *   - grad_loc = (W * d/dw, H * d/dh) * attn * grad_out summed over channels,
 *     grad_attn = bilinear value . grad_out summed over channels,
 *     grad_value = scatter of attn * corner weight * grad_out
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against fixtures in
 * tests/golden/ that were produced by the reference's own
 * ms_deform_attn_core_pytorch (functions/ms_deform_attn_func.py:102-122) and
 * autograd, see tests/golden/make_golden.py.
 *
 * Build: see oracle/build.py  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define DEFINE_ORACLE(SUFFIX, real)                                                              \
                                                                                                 \
/* one channel of one bilinear tap; cuh:33-84 */                                                 \
static real tap_##SUFFIX(const real *level, int H, int W, int M, int D, real h, real w, int m,   \
                         int c)                                                                  \
{                                                                                                \
    const int h0 = (int)floor((double)h), w0 = (int)floor((double)w);                            \
    const int h1 = h0 + 1, w1 = w0 + 1;                                                          \
    const real lh = h - (real)h0, lw = w - (real)w0;                                             \
    const real hh = (real)1 - lh, hw = (real)1 - lw;                                             \
    const int64_t ps = (int64_t)M * D;          /* pixel stride  */                              \
    const int64_t rs = (int64_t)W * ps;         /* row stride    */                              \
    const int64_t ch = (int64_t)m * D + c;                                                       \
    real v00 = 0, v01 = 0, v10 = 0, v11 = 0;                                                     \
    if (h0 >= 0 && w0 >= 0) v00 = level[h0 * rs + w0 * ps + ch];                                 \
    if (h0 >= 0 && w1 <= W - 1) v01 = level[h0 * rs + w1 * ps + ch];                             \
    if (h1 <= H - 1 && w0 >= 0) v10 = level[h1 * rs + w0 * ps + ch];                             \
    if (h1 <= H - 1 && w1 <= W - 1) v11 = level[h1 * rs + w1 * ps + ch];                         \
    const real a = hh * hw, b = hh * lw, cc = lh * hw, d = lh * lw;                              \
    return a * v00 + b * v01 + cc * v10 + d * v11;                                               \
}                                                                                                \
                                                                                                 \
/* cuh:237-299 */                                                                                \
void msda_oracle_forward_##SUFFIX(const real *value, const int64_t *shapes, const int64_t *lsi,  \
                                  const real *loc, const real *aw, int N, int S, int M, int D,   \
                                  int L, int Lq, int P, real *out)                               \
{                                                                                                \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                     \
    for (int b = 0; b < N; ++b)                                                                  \
        for (int m = 0; m < M; ++m)                                                              \
            for (int q = 0; q < Lq; ++q) {                                                       \
                const int64_t row = ((int64_t)b * Lq + q) * M + m;                               \
                const real *aw_r = aw + row * L * P;                                             \
                const real *loc_r = loc + row * L * P * 2;                                       \
                for (int c = 0; c < D; ++c) {                                                    \
                    real col = 0;                                                                \
                    for (int l = 0; l < L; ++l) {                                                \
                        const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];            \
                        const real *level = value + ((int64_t)b * S + lsi[l]) * M * D;           \
                        for (int p = 0; p < P; ++p) {                                            \
                            const real lw_ = loc_r[(l * P + p) * 2];                             \
                            const real lh_ = loc_r[(l * P + p) * 2 + 1];                         \
                            const real h = (real)((real)(lh_ * H) - 0.5);                        \
                            const real w = (real)((real)(lw_ * W) - 0.5);                        \
                            if (h > -1 && w > -1 && h < H && w < W)                              \
                                col += tap_##SUFFIX(level, H, W, M, D, h, w, m, c) *             \
                                       aw_r[l * P + p];                                          \
                        }                                                                        \
                    }                                                                            \
                    out[row * D + c] = col;                                                      \
                }                                                                                \
            }                                                                                    \
}                                                                                                \
                                                                                                 \
/* cuh:301-403 + cuh:87-159; outputs must be zero-filled by the caller like                      \
 * ms_deform_attn_cuda.cu:121-123 -- we do it here for convenience. */                           \
void msda_oracle_backward_##SUFFIX(const real *value, const int64_t *shapes, const int64_t *lsi, \
                                   const real *loc, const real *aw, const real *gout, int N,     \
                                   int S, int M, int D, int L, int Lq, int P, real *gvalue,      \
                                   real *gloc, real *gaw)                                        \
{                                                                                                \
    memset(gvalue, 0, sizeof(real) * (size_t)N * S * M * D);                                     \
    memset(gloc, 0, sizeof(real) * (size_t)N * Lq * M * L * P * 2);                              \
    memset(gaw, 0, sizeof(real) * (size_t)N * Lq * M * L * P);                                   \
    /* heads never alias in grad_value, so (b, m) is a race-free parallel axis */                \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                     \
    for (int b = 0; b < N; ++b)                                                                  \
        for (int m = 0; m < M; ++m)                                                              \
            for (int q = 0; q < Lq; ++q) {                                                       \
                const int64_t row = ((int64_t)b * Lq + q) * M + m;                               \
                const real *g_r = gout + row * D;                                                \
                for (int l = 0; l < L; ++l) {                                                    \
                    const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                \
                    const int64_t ps = (int64_t)M * D, rs = (int64_t)W * ps;                     \
                    const int64_t lvl_off = ((int64_t)b * S + lsi[l]) * M * D;                   \
                    const real *level = value + lvl_off;                                         \
                    real *glevel = gvalue + lvl_off;                                             \
                    for (int p = 0; p < P; ++p) {                                                \
                        const int64_t k = row * L * P + l * P + p;                               \
                        const real attn = aw[k];                                                 \
                        const real h = (real)((real)(loc[2 * k + 1] * H) - 0.5);                 \
                        const real w = (real)((real)(loc[2 * k] * W) - 0.5);                     \
                        if (!(h > -1 && w > -1 && h < H && w < W)) continue;                     \
                        const int h0 = (int)floor((double)h), w0 = (int)floor((double)w);        \
                        const int h1 = h0 + 1, w1 = w0 + 1;                                      \
                        const real lh = h - (real)h0, lw = w - (real)w0;                         \
                        const real hh = (real)1 - lh, hw = (real)1 - lw;                         \
                        const real a = hh * hw, bb = hh * lw, cc = lh * hw, d = lh * lw;         \
                        real acc_w = 0, acc_h = 0, acc_a = 0;                                    \
                        for (int c = 0; c < D; ++c) {                                            \
                            const int64_t ch = (int64_t)m * D + c;                               \
                            const real top = g_r[c];                                             \
                            const real tv = top * attn;                                          \
                            real gh = 0, gw = 0, v00 = 0, v01 = 0, v10 = 0, v11 = 0;             \
                            if (h0 >= 0 && w0 >= 0) {                                            \
                                v00 = level[h0 * rs + w0 * ps + ch];                             \
                                gh -= hw * v00; gw -= hh * v00;                                  \
                                glevel[h0 * rs + w0 * ps + ch] += a * tv;                        \
                            }                                                                    \
                            if (h0 >= 0 && w1 <= W - 1) {                                        \
                                v01 = level[h0 * rs + w1 * ps + ch];                             \
                                gh -= lw * v01; gw += hh * v01;                                  \
                                glevel[h0 * rs + w1 * ps + ch] += bb * tv;                       \
                            }                                                                    \
                            if (h1 <= H - 1 && w0 >= 0) {                                        \
                                v10 = level[h1 * rs + w0 * ps + ch];                             \
                                gh += hw * v10; gw -= lh * v10;                                  \
                                glevel[h1 * rs + w0 * ps + ch] += cc * tv;                       \
                            }                                                                    \
                            if (h1 <= H - 1 && w1 <= W - 1) {                                    \
                                v11 = level[h1 * rs + w1 * ps + ch];                             \
                                gh += lw * v11; gw += lh * v11;                                  \
                                glevel[h1 * rs + w1 * ps + ch] += d * tv;                        \
                            }                                                                    \
                            const real val = a * v00 + bb * v01 + cc * v10 + d * v11;            \
                            acc_a += top * val;                                                  \
                            acc_w += (real)W * gw * tv;                                          \
                            acc_h += (real)H * gh * tv;                                          \
                        }                                                                        \
                        gaw[k] = acc_a;                                                          \
                        gloc[2 * k] = acc_w;                                                     \
                        gloc[2 * k + 1] = acc_h;                                                 \
                    }                                                                            \
                }                                                                                \
            }                                                                                    \
}

DEFINE_ORACLE(f32, float)
DEFINE_ORACLE(f64, double)

/* lets the loader check it got the file it expects */
int msda_oracle_abi_version(void) { return 1; }
