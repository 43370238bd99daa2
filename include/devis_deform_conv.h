/*
 * devis_deform_conv.h -- C ABI of the modulated deformable convolution kernels (same library, libdevis_msda.so).
 *
 * DeVIS's mask head evaluates every 3x3 layer with torchvision.ops.deform_conv2d
 * (/root/reference/src/models/deformable_segmentation.py:244-267 ModulatedDeformableConv2d, :323-380 MaskHeadConv;
 * SURVEY.md section 8f-3).  torchvision is a third-party dependency of the reference (pinned at 0.12,
 * docs/INSTALL.md:9); its operator is vision::ops::deform_conv2d (torchvision/csrc/ops/deform_conv2d.cpp) with the CUDA
 * kernels deformable_im2col / deformable_col2im / deformable_col2im_coord (torchvision/csrc/ops/cuda/
 * deform_conv2d_kernel.cu).  The two entry points below replace those three kernels; the matrix products with the
 * convolution weights stay in cuBLAS on the caller's side (torchvision: at::addmm / at::bmm), exactly as the attention
 * library leaves the projection GEMMs to cuBLAS.
 *
 * Layouts.  `input` / `grad_input` are CHANNELS-LAST (N, H, W, C) -- a torch tensor in torch.channels_last memory
 * format, or x.permute(0, 2, 3, 1).contiguous().  `offset` (N, 2*kh*kw, Ho, Wo) and `mask` (N, kh*kw, Ho, Wo) keep
 * torchvision's layout: channel 2k is the row offset and 2k+1 the column offset of kernel position k = ky*kw + kx;
 * mask may be NULL (plain deformable convolution).  Columns: cols[((n*Ho + ho)*Wo + wo) * kh*kw*C + k*C + c], i.e. a
 * row-major (N*Ho*Wo, kh*kw*C) matrix; multiply it with weight.permute(0, 2, 3, 1).reshape(Cout, kh*kw*C)^T.
 * groups = 1 and offset_groups = 1 only.  dtype: DEVIS_MSDA_F32 or DEVIS_MSDA_F64.  Asynchronous on `stream`,
 * no allocation, no synchronisation; returns DEVIS_MSDA_OK or a DEVIS_MSDA_ERR_* code (devis_msda.h).
 */
#ifndef DEVIS_DEFORM_CONV_H_
#define DEVIS_DEFORM_CONV_H_

#include "devis_msda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* cols = modulated deformable im2col(input): cols[pixel, k, c] = mask[k] * bilinear(input[n, :, :, c], sample point k),
 * zero padding (torchvision bilinear_interpolate).  Fully written. */
int devis_dcn_im2col(const void *input_nhwc, const void *offset, const void *mask, void *cols, int batch, int height,
                     int width, int channels, int out_h, int out_w, int kernel_h, int kernel_w, int stride_h,
                     int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int dtype, void *stream);

/* Backward of the gather given grad_cols (same layout as cols):
 *   grad_input_nhwc  like input  -- zero-filled by this call, then accumulated (may be NULL: not needed)
 *   grad_offset      like offset -- fully written
 *   grad_mask        like mask   -- fully written (NULL iff mask is NULL) */
int devis_dcn_col2im(const void *input_nhwc, const void *offset, const void *mask, const void *grad_cols,
                     void *grad_input_nhwc, void *grad_offset, void *grad_mask, int batch, int height, int width,
                     int channels, int out_h, int out_w, int kernel_h, int kernel_w, int stride_h, int stride_w,
                     int pad_h, int pad_w, int dil_h, int dil_w, int dtype, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DEVIS_DEFORM_CONV_H_ */
