// deform_conv_capi.cu -- C ABI (include/devis_deform_conv.h) of the modulated deformable convolution kernels.
#include "../../include/devis_deform_conv.h"
#include "capi_common.h"
#include "deform_conv.cuh"

using namespace devis;

namespace {

int check_dims(const DcnDims &d, int dtype)
{
    if (dtype != DEVIS_MSDA_F32 && dtype != DEVIS_MSDA_F64) return DEVIS_MSDA_ERR_BAD_DTYPE;
    if (d.N < 0 || d.H <= 0 || d.W <= 0 || d.C <= 0 || d.Ho < 0 || d.Wo < 0 || d.kh <= 0 || d.kw <= 0 || d.sh <= 0 ||
        d.sw <= 0 || d.ph < 0 || d.pw < 0 || d.dh <= 0 || d.dw <= 0)
        return DEVIS_MSDA_ERR_BAD_SHAPE;
    if ((long long)d.H * d.W >= (1LL << 31)) return DEVIS_MSDA_ERR_TOO_LARGE;
    return DEVIS_MSDA_OK;
}

// lanes per tap: a group covers its channels in 16-byte pieces; 8 lanes = one 128-byte line per corner row.
// Grid: x = output plane (wo fastest), y = kernel position, z = batch item (in slices of at most 65535).
constexpr int kMaxGridZ = 65535;

template <class F4, class F8, class S8, class S32>
int dispatch(const DcnDims &d, int dtype, F4 f4, F8 f8, S8 s8, S32 s32)
{
    const long long plane = (long long)d.Ho * d.Wo;
    if (d.N == 0 || plane == 0) return DEVIS_MSDA_OK;
    int G;
    const bool vec = dtype == DEVIS_MSDA_F32 && d.C % 4 == 0;
    if (vec) G = d.C / 4 >= 8 ? 8 : 4;
    else G = d.C >= 32 ? 32 : 8;
    if (plane * G + 255 >= (1LL << 31) || d.kh * d.kw > 65535) return DEVIS_MSDA_ERR_TOO_LARGE;
    for (int n0 = 0; n0 < d.N; n0 += kMaxGridZ) {
        const dim3 grid((unsigned)((plane * G + 255) / 256), (unsigned)(d.kh * d.kw),
                        (unsigned)(d.N - n0 < kMaxGridZ ? d.N - n0 : kMaxGridZ));
        if (vec) {
            if (G == 8) f8(grid, n0);
            else f4(grid, n0);
        } else {
            if (G == 32) s32(grid, n0);
            else s8(grid, n0);
        }
        const int rc = devis_capi_check_launch();
        if (rc) return rc;
    }
    return DEVIS_MSDA_OK;
}

}  // namespace

extern "C" {

int devis_dcn_im2col(const void *input, const void *offset, const void *mask, void *cols, int batch, int height,
                     int width, int channels, int out_h, int out_w, int kernel_h, int kernel_w, int stride_h,
                     int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int dtype, void *stream)
{
    const DcnDims d{batch, height, width, channels, out_h, out_w, kernel_h, kernel_w,
                    stride_h, stride_w, pad_h, pad_w, dil_h, dil_w};
    const int rc = check_dims(d, dtype);
    if (rc) return rc;
    const long long n_taps = (long long)batch * out_h * out_w * kernel_h * kernel_w;
    if (n_taps > 0 && (!input || !offset || !cols)) return DEVIS_MSDA_ERR_NULL_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DEVIS_MSDA_F32) {
        const float *in = (const float *)input, *of = (const float *)offset, *mk = (const float *)mask;
        float *co = (float *)cols;
        return dispatch(d, dtype,
                        [&](dim3 b, int n0) { dcn_im2col_kernel<float, 4, 4><<<b, 256, 0, st>>>(in, of, mk, co, d, n0); },
                        [&](dim3 b, int n0) { dcn_im2col_kernel<float, 4, 8><<<b, 256, 0, st>>>(in, of, mk, co, d, n0); },
                        [&](dim3 b, int n0) { dcn_im2col_kernel<float, 1, 8><<<b, 256, 0, st>>>(in, of, mk, co, d, n0); },
                        [&](dim3 b, int n0) { dcn_im2col_kernel<float, 1, 32><<<b, 256, 0, st>>>(in, of, mk, co, d, n0); });
    }
    const double *in = (const double *)input, *of = (const double *)offset, *mk = (const double *)mask;
    double *co = (double *)cols;
    return dispatch(d, dtype, [&](dim3, int) {}, [&](dim3, int) {},
                    [&](dim3 b, int n0) { dcn_im2col_kernel<double, 1, 8><<<b, 256, 0, st>>>(in, of, mk, co, d, n0); },
                    [&](dim3 b, int n0) { dcn_im2col_kernel<double, 1, 32><<<b, 256, 0, st>>>(in, of, mk, co, d, n0); });
}

int devis_dcn_col2im(const void *input, const void *offset, const void *mask, const void *grad_cols, void *grad_input,
                     void *grad_offset, void *grad_mask, int batch, int height, int width, int channels, int out_h,
                     int out_w, int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w, int dil_h,
                     int dil_w, int dtype, void *stream)
{
    const DcnDims d{batch, height, width, channels, out_h, out_w, kernel_h, kernel_w,
                    stride_h, stride_w, pad_h, pad_w, dil_h, dil_w};
    const int rc = check_dims(d, dtype);
    if (rc) return rc;
    const long long n_taps = (long long)batch * out_h * out_w * kernel_h * kernel_w;
    if (n_taps > 0 && (!input || !offset || !grad_cols || !grad_offset || (mask && !grad_mask)))
        return DEVIS_MSDA_ERR_NULL_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    if (grad_input) {   // torchvision: at::zeros_like(input)
        const size_t bytes = (size_t)batch * height * width * channels * (dtype == DEVIS_MSDA_F64 ? 8 : 4);
        if (bytes) {
            const cudaError_t e = cudaMemsetAsync(grad_input, 0, bytes, st);
            if (e != cudaSuccess) return devis_capi_cuda_fail(e);
        }
    }
    if (dtype == DEVIS_MSDA_F32) {
        const float *in = (const float *)input, *of = (const float *)offset, *mk = (const float *)mask;
        const float *gc = (const float *)grad_cols;
        float *gi = (float *)grad_input, *go = (float *)grad_offset, *gm = (float *)grad_mask;
        return dispatch(d, dtype,
                        [&](dim3 b, int n0) { dcn_col2im_kernel<float, 4, 4><<<b, 256, 0, st>>>(in, of, mk, gc, gi, go, gm, d, n0); },
                        [&](dim3 b, int n0) { dcn_col2im_kernel<float, 4, 8><<<b, 256, 0, st>>>(in, of, mk, gc, gi, go, gm, d, n0); },
                        [&](dim3 b, int n0) { dcn_col2im_kernel<float, 1, 8><<<b, 256, 0, st>>>(in, of, mk, gc, gi, go, gm, d, n0); },
                        [&](dim3 b, int n0) { dcn_col2im_kernel<float, 1, 32><<<b, 256, 0, st>>>(in, of, mk, gc, gi, go, gm, d, n0); });
    }
    const double *in = (const double *)input, *of = (const double *)offset, *mk = (const double *)mask;
    const double *gc = (const double *)grad_cols;
    double *gi = (double *)grad_input, *go = (double *)grad_offset, *gm = (double *)grad_mask;
    return dispatch(d, dtype, [&](dim3, int) {}, [&](dim3, int) {},
                    [&](dim3 b, int n0) { dcn_col2im_kernel<double, 1, 8><<<b, 256, 0, st>>>(in, of, mk, gc, gi, go, gm, d, n0); },
                    [&](dim3 b, int n0) { dcn_col2im_kernel<double, 1, 32><<<b, 256, 0, st>>>(in, of, mk, gc, gi, go, gm, d, n0); });
}


// ---- fused gather + contraction (no column matrix) -------------------------------------------------------------------

static int fused_lanes(int channels, int out_channels, int dtype)
{
    if (dtype != DEVIS_MSDA_F32 || channels <= 0 || channels % 4) return 0;
    switch (out_channels) {
    case 1: case 2: case 4: case 8: case 16: case 32: case 64: break;
    default: return 0;
    }
    return channels / 4 > 4 ? 8 : 4;
}

int devis_dcn_fused_lanes(int channels, int out_channels, int dtype) { return fused_lanes(channels, out_channels, dtype); }

size_t devis_dcn_packed_weight_elems(int channels, int out_channels, int kernel_h, int kernel_w)
{
    const int G = fused_lanes(channels, out_channels, DEVIS_MSDA_F32);
    if (!G || kernel_h <= 0 || kernel_w <= 0) return 0;
    const int nblk = (channels / 4 + G - 1) / G;
    return (size_t)kernel_h * kernel_w * nblk * out_channels * G * 4;
}

int devis_dcn_pack_weight(const void *weight, void *packed, int channels, int out_channels, int kernel_h, int kernel_w,
                          void *stream)
{
    const int G = fused_lanes(channels, out_channels, DEVIS_MSDA_F32);
    if (!G) return DEVIS_MSDA_ERR_UNSUPPORTED;
    if (kernel_h <= 0 || kernel_w <= 0) return DEVIS_MSDA_ERR_BAD_SHAPE;
    if (!weight || !packed) return DEVIS_MSDA_ERR_NULL_POINTER;
    const int nblk = (channels / 4 + G - 1) / G, K = kernel_h * kernel_w;
    const long long total = (long long)K * nblk * out_channels * G * 4;
    const unsigned blocks = (unsigned)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256);
    dcn_pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float *)weight, (float *)packed, out_channels,
                                                                    channels, K, G, nblk);
    return devis_capi_check_launch();
}

// pixels per lane group: the weights fetched for one (position, channel block) serve PPG pixels
#define DCN_FUSED_LAUNCH(KERNEL, COUT, G, PPG, ...)                                                          \
    for (int n0 = 0; n0 < d.N; n0 += kMaxGridZ) {                                                            \
        const dim3 grid((unsigned)((d.Wo + (32 / G) * PPG - 1) / ((32 / G) * PPG)), (unsigned)((d.Ho + 7) / 8), \
                        (unsigned)(d.N - n0 < kMaxGridZ ? d.N - n0 : kMaxGridZ));                            \
        KERNEL<COUT, G, PPG><<<grid, 256, 0, st>>>(__VA_ARGS__, d, n0);                                      \
        const int rc_ = devis_capi_check_launch();                                                           \
        if (rc_) return rc_;                                                                                 \
    }
#define DCN_FUSED_CASES(KERNEL, G, PPG, PPG32, ...)                              \
    switch (out_channels) {                                                      \
    case 1: DCN_FUSED_LAUNCH(KERNEL, 1, G, PPG, __VA_ARGS__) break;              \
    case 2: DCN_FUSED_LAUNCH(KERNEL, 2, G, PPG, __VA_ARGS__) break;              \
    case 4: DCN_FUSED_LAUNCH(KERNEL, 4, G, PPG, __VA_ARGS__) break;              \
    case 8: DCN_FUSED_LAUNCH(KERNEL, 8, G, PPG, __VA_ARGS__) break;              \
    case 16: DCN_FUSED_LAUNCH(KERNEL, 16, G, PPG, __VA_ARGS__) break;            \
    case 32: DCN_FUSED_LAUNCH(KERNEL, 32, G, PPG32, __VA_ARGS__) break;          \
    default: DCN_FUSED_LAUNCH(KERNEL, 64, G, 1, __VA_ARGS__) break;              \
    }

int devis_dcn_fused_forward(const void *input, const void *offset, const void *mask, const void *packed_weight,
                            const void *bias, void *out, int batch, int height, int width, int channels, int out_h,
                            int out_w, int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                            int dil_h, int dil_w, int out_channels, void *stream)
{
    const DcnDims d{batch, height, width, channels, out_h, out_w, kernel_h, kernel_w,
                    stride_h, stride_w, pad_h, pad_w, dil_h, dil_w};
    const int rc = check_dims(d, DEVIS_MSDA_F32);
    if (rc) return rc;
    const int G = fused_lanes(channels, out_channels, DEVIS_MSDA_F32);
    if (!G) return DEVIS_MSDA_ERR_UNSUPPORTED;
    const long long n_pixels = (long long)batch * out_h * out_w;
    if (n_pixels == 0) return DEVIS_MSDA_OK;
    if (!input || !offset || !packed_weight || !out) return DEVIS_MSDA_ERR_NULL_POINTER;
    if ((long long)out_h * out_w >= (1LL << 31) || (long long)height * width * channels >= (1LL << 31))
        return DEVIS_MSDA_ERR_TOO_LARGE;
    cudaStream_t st = (cudaStream_t)stream;
    const float *in = (const float *)input, *of = (const float *)offset, *mk = (const float *)mask;
    const float4 *wp = (const float4 *)packed_weight;
    const float *bs = (const float *)bias;
    float *o = (float *)out;
    if (G == 8) { DCN_FUSED_CASES(dcn_fused_fwd_kernel, 8, 2, 2, in, of, mk, wp, bs, o) }
    else { DCN_FUSED_CASES(dcn_fused_fwd_kernel, 4, 2, 2, in, of, mk, wp, bs, o) }
    return DEVIS_MSDA_OK;
}

int devis_dcn_fused_backward(const void *input, const void *offset, const void *mask, const void *packed_weight,
                             const void *grad_out, void *grad_input, void *grad_offset, void *grad_mask, int batch,
                             int height, int width, int channels, int out_h, int out_w, int kernel_h, int kernel_w,
                             int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int out_channels,
                             void *stream)
{
    const DcnDims d{batch, height, width, channels, out_h, out_w, kernel_h, kernel_w,
                    stride_h, stride_w, pad_h, pad_w, dil_h, dil_w};
    const int rc = check_dims(d, DEVIS_MSDA_F32);
    if (rc) return rc;
    const int G = fused_lanes(channels, out_channels, DEVIS_MSDA_F32);
    if (!G) return DEVIS_MSDA_ERR_UNSUPPORTED;
    const long long n_pixels = (long long)batch * out_h * out_w;
    if (n_pixels > 0 && (!input || !offset || !packed_weight || !grad_out || !grad_offset || (mask && !grad_mask)))
        return DEVIS_MSDA_ERR_NULL_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    if (grad_input) {
        const size_t bytes = (size_t)batch * height * width * channels * 4;
        if (bytes) {
            const cudaError_t e = cudaMemsetAsync(grad_input, 0, bytes, st);
            if (e != cudaSuccess) return devis_capi_cuda_fail(e);
        }
    }
    if (n_pixels == 0) return DEVIS_MSDA_OK;
    if ((long long)out_h * out_w >= (1LL << 31) || (long long)height * width * channels >= (1LL << 31))
        return DEVIS_MSDA_ERR_TOO_LARGE;
    const float *in = (const float *)input, *of = (const float *)offset, *mk = (const float *)mask;
    const float4 *wp = (const float4 *)packed_weight;
    const float *go = (const float *)grad_out;
    float *gi = (float *)grad_input, *gof = (float *)grad_offset, *gm = (float *)grad_mask;
    if (G == 8) { DCN_FUSED_CASES(dcn_fused_bwd_kernel, 8, 1, 1, in, of, mk, wp, go, gi, gof, gm) }
    else { DCN_FUSED_CASES(dcn_fused_bwd_kernel, 4, 1, 1, in, of, mk, wp, go, gi, gof, gm) }
    return DEVIS_MSDA_OK;
}

}  // extern "C"
