/*
 * oracle/dcn_oracle.c — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * CPU restatement of the reference ModulatedDeformableConv2dTRT plugin's deformable im2col (FP32):
 *   TensorRT/plugin/modulated_deformable_conv2d/modulatedDeformableConv2dKernel.cu
 *     :85-116   dmcn_im2col_bilinear       — zero-padded bilinear tap rule
 *     :259-318  modulated_deformable_im2col_gpu_kernel<float>
 *               col[(c*kh*kw + i*kw + j), h, w] = mask * bilinear(im[c], h*s - p + i*d + off_h, w*s - p + j*d + off_w)
 *               offset channel 2*(i*kw+j) = dh, +1 = dw (mmcv convention, :297-301)
 * The GEMM out[b, g] = W_g (Co/g x Ci/g*kh*kw) . col_g and the bias add (:735-759) are done by the Python driver with
 * a float64-accumulated matmul (oracle/dcn.py).
 *
 * Pinning: mmcv's compiled op (mmcv-full 1.5.0, the reference binding's forward,
 * det2trt/models/functions/modulated_deformable_conv2d.py:84-104) is absent and un-vendored, so the reference's Python
 * path cannot run here. This file is pinned instead (1) against torchvision.ops.deform_conv2d(mask=…) — an independent
 * implementation of the same DCNv2 definition with the same offset channel order — in tests/test_oracle_golden.py, and
 * (2) on the GPU box against the reference's own CUDA launcher compiled into oracle/_ref (tests/test_dcn_gpu.py).
 */
#include <math.h>
#include <stdint.h>

#include "par.h"

typedef struct {
  const float *im, *offset, *mask;
  float *col;
  int C, H, W, kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w, dg, Ho, Wo;
} dcn_ctx;

static inline float dcn_bilinear(const float *im, int H, int W, float h, float w) {
  const int h_low = (int)floorf(h), w_low = (int)floorf(w);
  const int h_high = h_low + 1, w_high = w_low + 1;
  const float lh = h - (float)h_low, lw = w - (float)w_low;
  const float hh = 1 - lh, hw = 1 - lw;
  float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  if (h_low >= 0 && w_low >= 0) v1 = im[h_low * W + w_low];
  if (h_low >= 0 && w_high <= W - 1) v2 = im[h_low * W + w_high];
  if (h_high <= H - 1 && w_low >= 0) v3 = im[h_high * W + w_low];
  if (h_high <= H - 1 && w_high <= W - 1) v4 = im[h_high * W + w_high];
  return hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4;
}

static void dcn_body(long long begin, long long end, void *vctx) {
  const dcn_ctx *x = (const dcn_ctx *)vctx;
  const int HoWo = x->Ho * x->Wo, kk = x->kh * x->kw, cpg = x->C / x->dg;
  for (long long idx = begin; idx < end; ++idx) { /* idx = (c, h_col, w_col) */
    const int w_col = (int)(idx % x->Wo), h_col = (int)((idx / x->Wo) % x->Ho), c = (int)(idx / HoWo);
    const int g = c / cpg;
    const int h_in = h_col * x->stride_h - x->pad_h, w_in = w_col * x->stride_w - x->pad_w;
    const float *im = x->im + (long long)c * x->H * x->W;
    const float *off = x->offset + (long long)g * 2 * kk * HoWo;
    const float *msk = x->mask + (long long)g * kk * HoWo;
    for (int i = 0; i < x->kh; ++i)
      for (int j = 0; j < x->kw; ++j) {
        const int t = i * x->kw + j, p = h_col * x->Wo + w_col;
        const float oh = off[(2 * t) * HoWo + p], ow = off[(2 * t + 1) * HoWo + p], m = msk[t * HoWo + p];
        const float h_im = (float)(h_in + i * x->dil_h) + oh, w_im = (float)(w_in + j * x->dil_w) + ow;
        float val = 0.f;
        if (h_im > -1 && w_im > -1 && h_im < (float)x->H && w_im < (float)x->W)
          val = dcn_bilinear(im, x->H, x->W, h_im, w_im);
        x->col[((long long)c * kk + t) * HoWo + p] = val * m;
      }
  }
}

/* One image: im [C,H,W], offset [dg*2*kh*kw, Ho, Wo], mask [dg*kh*kw, Ho, Wo] -> col [C*kh*kw, Ho*Wo]. */
void oracle_dcn_im2col_f32(const float *im, const float *offset, const float *mask, float *col, int C, int H, int W,
                           int kh, int kw, int pad_h, int pad_w, int stride_h, int stride_w, int dil_h, int dil_w,
                           int dg, int Ho, int Wo) {
  dcn_ctx x = {im, offset, mask, col, C, H, W, kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w, dg, Ho, Wo};
  oracle_parallel_for((long long)C * Ho * Wo, dcn_body, &x);
}
