/*
 * b200_bev_ops.h — C ABI of libb200_bev_ops.so: BEVFormer's attention-sampling hot path for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the three TensorRT plugins of DerryHub/BEVFormer_tensorrt that sit on that path.
 * Every entry point replaces one launcher of the reference (the level directly under IPluginV2DynamicExt::enqueue)
 * and keeps its argument order and meaning; `file:line` below are relative to the reference repository.
 *
 * Conventions (all entry points)
 *   - All data pointers are DEVICE pointers owned by the caller (TensorRT / PyTorch); nothing is allocated, freed
 *     or retained. `stream` is a cudaStream_t passed as void*; launches are asynchronous on it, no syncs.
 *   - Return value: B200_OK (0) on success; non-zero status otherwise. Never aborts/exits the process (the reference
 *     ASSERT()s -> abort(), TensorRT/common/helper.h:43-49, and printf()s launch errors, …Kernel.cu:1124-1127).
 *     B200_ERR_UNSUPPORTED == 1 matches enqueue()'s "return 1" for an unsupported dtype (…Plugin.cpp:136-139).
 *   - No CPU fallback exists: without a CUDA device the launch fails and the status says so.
 */
#ifndef B200_BEV_OPS_H
#define B200_BEV_OPS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  B200_OK = 0,
  B200_ERR_UNSUPPORTED = 1, /* dtype / format / shape combination not implemented */
  B200_ERR_BAD_PARAM = 2,   /* null pointer, non-positive dimension, size overflow, misalignment */
  B200_ERR_LAUNCH = 3       /* CUDA reported an error at launch (cudaGetLastError) */
};

/* Library / build identification: "b200_bev_ops <version> sm_100a". */
const char *b200_bev_ops_version(void);
/* Text for a status code. */
const char *b200_status_string(int status);
/* Number of kernels this library has launched since load (bench.py's gpu_launches claim is counted here). */
unsigned long long b200_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-scale deformable attention (plugin MultiScaleDeformableAttnTRT / …TRT2)
 *
 *   value            [batch, spatial_size, num_heads, channels]      levels concatenated along spatial_size
 *   spatial_shapes   int32 [num_levels, 2] (h, w), device memory
 *   reference_points [batch, num_query, 1, 2*points_per_group]       normalised (x, y) per group (Z anchor)
 *   sampling_offsets [batch, num_query, num_heads, num_levels*num_point*2]   pixels of that level, (x, y) innermost
 *   attn_weight      [batch, num_query, num_heads, num_levels*num_point]     PRE-softmax logits
 *   out              [batch, num_query, num_heads, channels]
 *
 * Semantics = the reference FP32 kernel (…/multiScaleDeformableAttnKernel.cu:611-688, :133-178): softmax over
 * all num_levels*num_point logits, loc = fma(ref, size, off) - 0.5 (bit-exact index arithmetic), zero-padded bilinear
 * taps. The FP16 entry stores FP16 but evaluates indices and accumulates in FP32 (the reference's __half kernels do
 * the coordinate math in half precision, :740-756; that loss is deliberately not reproduced).
 * ---------------------------------------------------------------------------------------------------------- */

/* replaces ms_deformable_im2col_cuda<float>   — …/multiScaleDeformableAttnKernel.h:12-19, .cu:1106-1128 */
int b200_msda_f32(const float *value, const int32_t *spatial_shapes, const float *reference_points,
                  const float *sampling_offsets, const float *attn_weight, int batch, int spatial_size, int num_heads,
                  int channels, int num_levels, int num_query, int num_point, int points_per_group, float *out,
                  void *stream);

/* replaces ms_deformable_im2col_cuda<__half>  — …Kernel.h:12-19 (T=__half), .cu:1130-1149; pointers are __half* */
int b200_msda_f16(const void *value, const int32_t *spatial_shapes, const void *reference_points,
                  const void *sampling_offsets, const void *attn_weight, int batch, int spatial_size, int num_heads,
                  int channels, int num_levels, int num_query, int num_point, int points_per_group, void *out,
                  void *stream);

/* replaces ms_deformable_im2col_cuda_h2        — …Kernel.h:21-27, .cu:1151-1169. Same memory layout as _f16
 * (__half2 is a reinterpretation); needs even `channels` like the reference (…Plugin.cpp:240-245). */
int b200_msda_f16_h2(const void *value, const int32_t *spatial_shapes, const void *reference_points,
                     const void *sampling_offsets, const void *attn_weight, int batch, int spatial_size,
                     int num_heads, int channels, int num_levels, int num_query, int num_point, int points_per_group,
                     void *out, void *stream);

/* replaces ms_deformable_im2col_cuda_int8<float> / <__half2> — …Kernel.h:29-38, .cu:1172-1218.
 * value / offsets / attn_weight / out are int8 with per-tensor scales (real = q * scale, TensorRT PTQ convention,
 * …Plugin.cpp:75-77); reference_points are float (ref_is_half == 0) or __half (ref_is_half != 0).
 * Needs channels % 4 == 0 and num_point % 4 == 0 like the reference (…Plugin.cpp:151-156).
 * Arithmetic: in-register dequantisation, FP32 evaluation, one requantisation of the result with
 * round-half-away-from-zero and saturation (T2int8, .cu:51-55). */
int b200_msda_i8(const int8_t *value, float scale_value, const int32_t *spatial_shapes, const void *reference_points,
                 int ref_is_half, const int8_t *sampling_offsets, float scale_offset, const int8_t *attn_weight,
                 float scale_weight, int batch, int spatial_size, int num_heads, int channels, int num_levels,
                 int num_query, int num_point, int points_per_group, int8_t *out, float scale_out, void *stream);

/* Test/diagnostic entry: writes, for every (batch, query, head, level*point), the sampling-index record
 * {in_range, h_low, w_low, tap_mask} (4 x int32) computed by the same device code as the kernels above.
 * dtype: 0 = float inputs, 1 = __half inputs. Used by the bit-exact index parity tests. */
int b200_msda_debug_indices(int dtype, const int32_t *spatial_shapes, const void *reference_points,
                            const void *sampling_offsets, int batch, int num_heads, int num_levels, int num_query,
                            int num_point, int points_per_group, int32_t *records, void *stream);

/* Selects the FP16 accumulate flavour for b200_msda_f16*: 0 = exact (convert taps to FP32, FFMA),
 * 1 = mixed (fma.rn.f32.f16: FP16 tap x FP16 weight -> FP32 accumulator). Returns the previous setting. */
int b200_msda_set_f16_mode(int mode);

/* ------------------------------------------------------------------------------------------------------------
 * Plugin-shaped entry: the argument list of IPluginV2DynamicExt::enqueue
 * (…/multiScaleDeformableAttnPlugin.cpp:71-140) with a POD mirror of nvinfer1::PluginTensorDesc, so a TensorRT
 * plugin shell (csrc/trt_plugin/, compiled only where NvInfer.h exists) or any other host can forward verbatim.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t nbDims;
  int32_t d[8];
} b200_dims; /* layout of nvinfer1::Dims */

typedef struct {
  b200_dims dims;
  int32_t type;   /* nvinfer1::DataType: 0 kFLOAT, 1 kHALF, 2 kINT8, 3 kINT32 */
  int32_t format; /* nvinfer1::TensorFormat: 0 kLINEAR, 1 kCHW2 (fp16x2), 2 kHWC8, 3 kCHW4 (int8x4) */
  float scale;
} b200_tensor_desc; /* layout of nvinfer1::PluginTensorDesc */

/* inputs: value, spatial_shapes, reference_points, sampling_offsets, attn_weight; outputs: out.
 * use_h2 selects the …TRT2 flavour. Returns 0 ok / 1 unsupported dtype, as the reference does. */
int b200_msda_enqueue(const b200_tensor_desc *input_desc, const b200_tensor_desc *output_desc,
                      const void *const *inputs, void *const *outputs, void *workspace, void *stream, int use_h2);

/* Mirror of supportsFormatCombination (…Plugin.cpp:148-189): 1 if descriptor `pos` of in_out[0..5] is acceptable. */
int b200_msda_supports_format(int pos, const b200_tensor_desc *in_out, int nb_inputs, int nb_outputs);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* B200_BEV_OPS_H */
