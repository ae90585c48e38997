// oracle/ref_shim.cu — TEST INFRASTRUCTURE ONLY (GPU-side oracle).
//
// Exports the reference's own kernel launchers with C linkage so tests / bench.py can call the UNMODIFIED
// reference CUDA kernels through ctypes for a same-box A/B. The launchers are declared by the reference's own
// headers, included from where they lie under /root/reference (never copied):
//   TensorRT/plugin/multi_scale_deformable_attn/multiScaleDeformableAttnKernel.h:12-38
//   TensorRT/plugin/grid_sampler/gridSamplerKernel.h:14-26
//   TensorRT/plugin/modulated_deformable_conv2d/modulatedDeformableConv2dKernel.h:11-29
//   TensorRT/plugin/rotate/rotateKernel.h:14-26
// This file contains no arithmetic. It is linked only into oracle/_ref/libref_kernels.so.
#include <cublas_v2.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "gridSamplerKernel.h"
#include "modulatedDeformableConv2dKernel.h"
#include "multiScaleDeformableAttnKernel.h"
#include "rotateKernel.h"

extern "C" {

void ref_msda_f32(const float *value, const int32_t *shapes, const float *ref, const float *off, const float *w,
                  int B, int S, int M, int C, int L, int Q, int P, int G, float *out, void *stream) {
  ms_deformable_im2col_cuda<float>(value, shapes, ref, off, w, B, S, M, C, L, Q, P, G, out, (cudaStream_t)stream);
}

void ref_msda_f16(const void *value, const int32_t *shapes, const void *ref, const void *off, const void *w, int B,
                  int S, int M, int C, int L, int Q, int P, int G, void *out, void *stream) {
  ms_deformable_im2col_cuda<__half>((const __half *)value, shapes, (const __half *)ref, (const __half *)off,
                                    (const __half *)w, B, S, M, C, L, Q, P, G, (__half *)out, (cudaStream_t)stream);
}

void ref_msda_f16_h2(const void *value, const int32_t *shapes, const void *ref, const void *off, const void *w, int B,
                     int S, int M, int C, int L, int Q, int P, int G, void *out, void *stream) {
  ms_deformable_im2col_cuda_h2((const __half2 *)value, shapes, (const __half2 *)ref, (const __half2 *)off,
                               (const __half *)w, B, S, M, C, L, Q, P, G, (__half2 *)out, (cudaStream_t)stream);
}

void ref_msda_i8_f32ref(const void *value, float sv, const int32_t *shapes, const float *ref, const void *off,
                        float so, const void *w, float sw, int B, int S, int M, int C, int L, int Q, int P, int G,
                        void *out, float sout, void *stream) {
  ms_deformable_im2col_cuda_int8<float>((const int8_4 *)value, sv, shapes, ref, (const int8_4 *)off, so,
                                        (const int8_4 *)w, sw, B, S, M, C, L, Q, P, G, (int8_4 *)out, sout,
                                        (cudaStream_t)stream);
}

void ref_msda_i8_h2ref(const void *value, float sv, const int32_t *shapes, const void *ref, const void *off, float so,
                       const void *w, float sw, int B, int S, int M, int C, int L, int Q, int P, int G, void *out,
                       float sout, void *stream) {
  ms_deformable_im2col_cuda_int8<__half2>((const int8_4 *)value, sv, shapes, (const __half2 *)ref,
                                          (const int8_4 *)off, so, (const int8_4 *)w, sw, B, S, M, C, L, Q, P, G,
                                          (int8_4 *)out, sout, (cudaStream_t)stream);
}

// dims arrays are HOST int[nb_dims]; dtype: 0 f32, 1 f16 (kLINEAR), 2 f16 as half2 (kCHW2)
void ref_grid_sample(int dtype, void *out, const void *in, const void *grid, int *out_dims, int *in_dims,
                     int *grid_dims, int nb_dims, int interp, int padding, int align_corners, void *stream) {
  auto im = (GridSamplerInterpolation)interp;
  auto pm = (GridSamplerPadding)padding;
  if (dtype == 0)
    grid_sample<float>((float *)out, (const float *)in, (const float *)grid, out_dims, in_dims, grid_dims, nb_dims,
                       im, pm, align_corners != 0, (cudaStream_t)stream);
  else if (dtype == 1)
    grid_sample<__half>((__half *)out, (const __half *)in, (const __half *)grid, out_dims, in_dims, grid_dims,
                        nb_dims, im, pm, align_corners != 0, (cudaStream_t)stream);
  else
    grid_sample<__half2>((__half2 *)out, (const __half2 *)in, (const __half2 *)grid, out_dims, in_dims, grid_dims,
                         nb_dims, im, pm, align_corners != 0, (cudaStream_t)stream);
}

void ref_grid_sample_int8(void *out, float scale_o, const void *in, float scale_i, const void *grid, float scale_g,
                          int *out_dims, int *in_dims, int *grid_dims, int nb_dims, int interp, int padding,
                          int align_corners, void *stream) {
  grid_sample_int8((int8_4 *)out, scale_o, (const int8_4 *)in, scale_i, (const int8_4 *)grid, scale_g, out_dims,
                   in_dims, grid_dims, nb_dims, (GridSamplerInterpolation)interp, (GridSamplerPadding)padding,
                   align_corners != 0, (cudaStream_t)stream);
}

static cublasHandle_t g_handle = nullptr;
static cublasHandle_t handle_for(cudaStream_t s) {
  if (!g_handle) cublasCreate(&g_handle);
  cublasSetStream(g_handle, s);
  return g_handle;
}

// dtype: 0 f32, 1 f16. workspace: device buffer sized by the caller per …Conv2dPlugin.cpp:73-115.
void ref_dcn(int dtype, const void *input, const void *weight, const void *bias, const void *offset, const void *mask,
             void *output, void *workspace, int batch, int channels, int height, int width, int channels_out,
             int kernel_w, int kernel_h, int stride_w, int stride_h, int pad_w, int pad_h, int dilation_w,
             int dilation_h, int group, int deformable_group, int im2col_step, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == 0)
    ModulatedDeformConvForwardCUDAKernel<float>((const float *)input, (const float *)weight, (const float *)bias,
                                                (const float *)offset, (const float *)mask, (float *)output, workspace,
                                                batch, channels, height, width, channels_out, kernel_w, kernel_h,
                                                stride_w, stride_h, pad_w, pad_h, dilation_w, dilation_h, group,
                                                deformable_group, im2col_step, handle_for(s), s);
  else
    ModulatedDeformConvForwardCUDAKernel<__half>(
        (const __half *)input, (const __half *)weight, (const __half *)bias, (const __half *)offset,
        (const __half *)mask, (__half *)output, workspace, batch, channels, height, width, channels_out, kernel_w,
        kernel_h, stride_w, stride_h, pad_w, pad_h, dilation_w, dilation_h, group, deformable_group, im2col_step,
        handle_for(s), s);
}

// dtype: 0 f32, 1 f16 (kLINEAR), 2 f16 as half2 (kCHW2). dims = HOST int[3] {C, H, W}; angle/center device, img dtype.
void ref_rotate(int dtype, void *out, void *in, void *angle, void *center, int *dims, int interp, void *stream) {
  auto im = (RotateInterpolation)interp;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == 0)
    rotate<float>((float *)out, (float *)in, (float *)angle, (float *)center, dims, im, s);
  else if (dtype == 1)
    rotate<__half>((__half *)out, (__half *)in, (__half *)angle, (__half *)center, dims, im, s);
  else
    rotate_h2((__half2 *)out, (__half2 *)in, (__half *)angle, (__half *)center, dims, im, s);
}

// kCHW4 image; angle/center fp32 (angle_is_half = 0) or fp16 (1).
void ref_rotate_int8(void *out, float scale_o, const void *in, float scale_i, const void *angle, const void *center,
                     int angle_is_half, int *dims, int interp, void *stream) {
  auto im = (RotateInterpolation)interp;
  cudaStream_t s = (cudaStream_t)stream;
  if (angle_is_half)
    rotate_int8<__half>((int8_4 *)out, scale_o, (const int8_4 *)in, scale_i, (const __half *)angle,
                        (const __half *)center, dims, im, s);
  else
    rotate_int8<float>((int8_4 *)out, scale_o, (const int8_4 *)in, scale_i, (const float *)angle,
                       (const float *)center, dims, im, s);
}

} // extern "C"
