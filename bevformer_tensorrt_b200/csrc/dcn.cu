// dcn.cu — modulated deformable convolution (DCNv2, plugin ModulatedDeformableConv2dTRT / …TRT2) for B200 (sm_100a).
//
// Replaces ModulatedDeformConvForwardCUDAKernel<float|__half|__half2>
// (TensorRT/plugin/modulated_deformable_conv2d/modulatedDeformableConv2dKernel.cu:695-895) and its kernels
// (deformable im2col :259-461, bias epilogue :550-568).
//
// Round-1 structure (v1): gather + library GEMM, but batched and without the reference's serial per-image loop.
//   1. dcn_im2col_kernel: thread = (image, output pixel, slice of input channels). The nine (kh*kw) sampling
//      positions, their tap offsets, bilinear weights x mask and validity are computed ONCE per thread in fp32 and
//      reused for every channel of the slice (the reference recomputes them per channel, :295-313); columns are written
//      pixel-contiguous (coalesced) as col[b][c*kh*kw + t][p].
//   2. one cublasGemmStridedBatchedEx over the images (fp32 accumulate — the reference accumulates FP16 GEMMs in FP16,
//      common/cuda_helper.cu:101-110), output pre-loaded with the bias (beta = 1) so no separate bias pass reads it back.
// The dense contraction is the one place on this path where tensor cores apply; the fused implicit-GEMM
// (gather straight into the UMMA operand tile in shared memory, tcgen05.mma, accumulators in TMEM) is the planned v2
//  — v1 exists to have a correct, measured DCN behind the final ABI first.
#include <cublas_v2.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace b200 {

// dcn_fused.cu
size_t dcn_fused_workspace_bytes(int batch, int channels, int height, int width, int channels_out, int kk);
bool dcn_fused_supported(int channels, int channels_out, int kk, int group, int deformable_group);
int dcn_fused_f16(const __half *input, const __half *weight, const __half *bias, const __half *offset,
                  const __half *mask, __half *output, void *workspace, int batch, int channels, int height, int width,
                  int channels_out, int kernel_w, int kernel_h, int stride_w, int stride_h, int pad_w, int pad_h,
                  int dilation_w, int dilation_h, int Ho, int Wo, int flags, cudaStream_t stream);
int dcn_pack_weights_f16(const __half *weight, __half *packed, int channels_out, int channels, int kk, cudaStream_t stream);

int dcn_fused_i8(const int8_t *input, float scale_i, const int8_t *weight, float scale_w, const void *bias, int bias_is_half,
                 const int8_t *offset, float scale_off, const int8_t *mask, float scale_mask, int8_t *output, float scale_o,
                 void *workspace, int batch, int channels, int height, int width, int channels_out, int kernel_w,
                 int kernel_h, int stride_w, int stride_h, int pad_w, int pad_h, int dilation_w, int dilation_h, int Ho,
                 int Wo, cudaStream_t stream);

static int g_dcn_fused = -1;  // -1: read B200_DCN_FUSED once (default on)
static bool dcn_fused_enabled() {
  if (g_dcn_fused < 0) {
    const char *e = getenv("B200_DCN_FUSED");
    g_dcn_fused = (e && e[0] == '0') ? 0 : 1;
  }
  return g_dcn_fused == 1;
}

struct DcnParams {
  const void *im, *offset, *mask;
  void *col;
  int B, C, H, W, kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w, dg, Ho, Wo;
  int slices, cps;  // channel slices per (image, pixel), channels per slice
};

template <typename T>
__device__ __forceinline__ float ld_f(const T *p);
template <>
__device__ __forceinline__ float ld_f<float>(const float *p) {
  return __ldg(p);
}
template <>
__device__ __forceinline__ float ld_f<__half>(const __half *p) {
  return __half2float(__ldg(p));
}
template <typename T>
__device__ __forceinline__ void st_f(T *p, float v);
template <>
__device__ __forceinline__ void st_f<float>(float *p, float v) {
  *p = v;
}
template <>
__device__ __forceinline__ void st_f<__half>(__half *p, float v) {
  *p = __float2half_rn(v);
}

// KK = kh*kw as a template constant (9 for every DCN in the BEVFormer backbones) so the tap table lives in registers;
// KK = 0 selects the generic loop.
template <typename T, int KK>
__global__ void __launch_bounds__(256) dcn_im2col_kernel(const DcnParams p) {
  const int HoWo = p.Ho * p.Wo;
  const int kk = KK > 0 ? KK : p.kh * p.kw;
  const long long total = static_cast<long long>(p.B) * p.slices * HoWo;
  const int cpg = p.C / p.dg;  // channels per deformable group
  const T *im = static_cast<const T *>(p.im);
  const T *offset = static_cast<const T *>(p.offset);
  const T *mask = static_cast<const T *>(p.mask);
  T *col = static_cast<T *>(p.col);
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pix = static_cast<int>(idx % HoWo);
    const int s = static_cast<int>((idx / HoWo) % p.slices);
    const int b = static_cast<int>(idx / (static_cast<long long>(HoWo) * p.slices));
    const int w_col = pix % p.Wo, h_col = pix / p.Wo;
    const int c0 = s * p.cps, c1 = min(p.C, c0 + p.cps);
    const int g = c0 / cpg;  // a slice never straddles deformable groups (host guarantees cps | cpg)
    const int h_in = h_col * p.stride_h - p.pad_h, w_in = w_col * p.stride_w - p.pad_w;
    const T *off_b = offset + (static_cast<long long>(b) * p.dg + g) * 2 * kk * HoWo + pix;
    const T *msk_b = mask + (static_cast<long long>(b) * p.dg + g) * kk * HoWo + pix;
    const T *im_b = im + (static_cast<long long>(b) * p.C + c0) * p.H * p.W;
    T *col_b = col + (static_cast<long long>(b) * p.C + c0) * kk * HoWo + pix;

    for (int t = 0; t < kk; ++t) {
      const int i = t / p.kw, j = t - i * p.kw;
      const float oh = ld_f(off_b + static_cast<long long>(2 * t) * HoWo);
      const float ow = ld_f(off_b + static_cast<long long>(2 * t + 1) * HoWo);
      const float m = ld_f(msk_b + static_cast<long long>(t) * HoWo);
      // h_im = (h_in + i*dil) + offset : integer part exact, one fp32 add (…Kernel.cu:306-307)
      const float h_im = __fadd_rn(static_cast<float>(h_in + i * p.dil_h), oh);
      const float w_im = __fadd_rn(static_cast<float>(w_in + j * p.dil_w), ow);
      const bool ok = h_im > -1.f && w_im > -1.f && h_im < static_cast<float>(p.H) && w_im < static_cast<float>(p.W);
      const float hf = floorf(h_im), wf = floorf(w_im);
      const int h_low = ok ? static_cast<int>(hf) : 0, w_low = ok ? static_cast<int>(wf) : 0;
      const float lh = __fsub_rn(h_im, hf), lw = __fsub_rn(w_im, wf), hh = 1.f - lh, hw = 1.f - lw;
      const bool tp = h_low >= 0, bt = h_low + 1 <= p.H - 1, lf = w_low >= 0, rt = w_low + 1 <= p.W - 1;
      const float w1 = (ok && tp && lf) ? hh * hw * m : 0.f, w2 = (ok && tp && rt) ? hh * lw * m : 0.f;
      const float w3 = (ok && bt && lf) ? lh * hw * m : 0.f, w4 = (ok && bt && rt) ? lh * lw * m : 0.f;
      const int h0 = max(h_low, 0), w0 = max(w_low, 0);
      const int o1 = h0 * p.W + w0, dx = (lf && rt) ? 1 : 0, dy = (tp && bt) ? p.W : 0;
      const T *ip = im_b;
      T *cp = col_b + static_cast<long long>(t) * HoWo;
#pragma unroll 4
      for (int c = c0; c < c1; ++c, ip += p.H * p.W, cp += static_cast<long long>(kk) * HoWo) {
        // the reference multiplies the interpolated value by the mask (:313); folding the mask into the four weights
        // differs by rounding only
        const float v = fmaf(w4, ld_f(ip + o1 + dy + dx),
                             fmaf(w3, ld_f(ip + o1 + dy), fmaf(w2, ld_f(ip + o1 + dx), w1 * ld_f(ip + o1))));
        st_f(cp, v);
      }
    }
  }
}

template <typename T>
__global__ void dcn_fill_bias_kernel(T *out, const T *bias, int Co, int HoWo, long long n) {
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n;
       idx += static_cast<long long>(gridDim.x) * blockDim.x)
    out[idx] = bias[(idx / HoWo) % Co];
}

static cublasHandle_t g_cublas = nullptr;
static std::mutex g_cublas_mutex;

static cublasHandle_t get_handle() {
  std::lock_guard<std::mutex> lock(g_cublas_mutex);
  if (!g_cublas && cublasCreate(&g_cublas) != CUBLAS_STATUS_SUCCESS) g_cublas = nullptr;
  return g_cublas;
}

static size_t col_bytes(int batch, int channels, int kh, int kw, int Ho, int Wo, size_t elem) {
  size_t b = static_cast<size_t>(batch) * channels * kh * kw * Ho * Wo * elem;
  return (b + 255) / 256 * 256;
}

template <typename T>
static int dcn_forward(const T *input, const T *weight, const T *bias, const T *offset, const T *mask, T *output,
                       void *workspace, int batch, int channels, int height, int width, int channels_out, int kernel_w,
                       int kernel_h, int stride_w, int stride_h, int pad_w, int pad_h, int dilation_w, int dilation_h,
                       int group, int deformable_group, void *cublas_handle, cudaStream_t stream) {
  if (!input || !weight || !offset || !mask || !output || !workspace) return B200_ERR_BAD_PARAM;
  if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || channels_out <= 0 || kernel_w <= 0 ||
      kernel_h <= 0 || stride_w <= 0 || stride_h <= 0 || dilation_w <= 0 || dilation_h <= 0 || group <= 0 ||
      deformable_group <= 0)
    return B200_ERR_BAD_PARAM;
  // the reference exit(1)s on these (…Conv2dPlugin.cpp:300-336); here they are a status
  if (channels % group || channels_out % group || channels % deformable_group) return B200_ERR_UNSUPPORTED;
  const int Ho = (height + 2 * pad_h - (dilation_h * (kernel_h - 1) + 1)) / stride_h + 1;
  const int Wo = (width + 2 * pad_w - (dilation_w * (kernel_w - 1) + 1)) / stride_w + 1;
  if (Ho <= 0 || Wo <= 0) return B200_ERR_BAD_PARAM;
  const long long HoWo = static_cast<long long>(Ho) * Wo;
  const int kk = kernel_h * kernel_w;
  if (static_cast<long long>(channels) * kk * HoWo >= (1ll << 31)) return B200_ERR_BAD_PARAM;

  if (sizeof(T) == 2 && dcn_fused_enabled() && dcn_fused_supported(channels, channels_out, kk, group, deformable_group) &&
      reinterpret_cast<uintptr_t>(workspace) % 256 == 0)
    return dcn_fused_f16(reinterpret_cast<const __half *>(input), reinterpret_cast<const __half *>(weight),
                         reinterpret_cast<const __half *>(bias), reinterpret_cast<const __half *>(offset),
                         reinterpret_cast<const __half *>(mask), reinterpret_cast<__half *>(output), workspace, batch,
                         channels, height, width, channels_out, kernel_w, kernel_h, stride_w, stride_h, pad_w, pad_h,
                         dilation_w, dilation_h, Ho, Wo, 0, stream);

  DcnParams p{};
  p.im = input, p.offset = offset, p.mask = mask, p.col = workspace;
  p.B = batch, p.C = channels, p.H = height, p.W = width, p.kh = kernel_h, p.kw = kernel_w;
  p.pad_h = pad_h, p.pad_w = pad_w, p.stride_h = stride_h, p.stride_w = stride_w;
  p.dil_h = dilation_h, p.dil_w = dilation_w, p.dg = deformable_group, p.Ho = Ho, p.Wo = Wo;
  const int cpg = channels / deformable_group;
  int cps = 16;  // channels per thread: amortises the 9-tap index math; must divide the deformable group size
  while (cps > 1 && cpg % cps) cps >>= 1;
  p.cps = cps, p.slices = channels / cps;
  const long long total = static_cast<long long>(batch) * p.slices * HoWo;
  const unsigned blocks = static_cast<unsigned>(total / 256 + 1 < (1 << 22) ? total / 256 + 1 : (1 << 22));
  if (kk == 9)
    dcn_im2col_kernel<T, 9><<<blocks, 256, 0, stream>>>(p);
  else
    dcn_im2col_kernel<T, 0><<<blocks, 256, 0, stream>>>(p);
  int st = check_launch();
  if (st != B200_OK) return st;

  const long long out_n = static_cast<long long>(batch) * channels_out * HoWo;
  float beta = 0.f;
  if (bias) {
    const unsigned fb = static_cast<unsigned>(out_n / 256 + 1 < (1 << 20) ? out_n / 256 + 1 : (1 << 20));
    dcn_fill_bias_kernel<T><<<fb, 256, 0, stream>>>(output, bias, channels_out, static_cast<int>(HoWo), out_n);
    st = check_launch();
    if (st != B200_OK) return st;
    beta = 1.f;
  }

  cublasHandle_t h = cublas_handle ? static_cast<cublasHandle_t>(cublas_handle) : get_handle();
  if (!h) return B200_ERR_LAUNCH;
  if (cublasSetStream(h, stream) != CUBLAS_STATUS_SUCCESS) return B200_ERR_LAUNCH;
  const float alpha = 1.f;
  const int m = channels_out / group, n = static_cast<int>(HoWo), k = channels / group * kk;
  const cudaDataType_t dt = sizeof(T) == 2 ? CUDA_R_16F : CUDA_R_32F;
  const T *col = static_cast<const T *>(workspace);
  for (int g = 0; g < group; ++g) {
    // row-major out_b[g] (m x n) = W_g (m x k) . col_b[g] (k x n)  ==  column-major (n x m) = col^T-view . W^T-view,
    // the same operand order the reference hands cuBLAS (…Kernel.cu:749-751), batched over images with strides.
    const cublasStatus_t cs = cublasGemmStridedBatchedEx(
        h, CUBLAS_OP_N, CUBLAS_OP_N, n, m, k, &alpha, col + static_cast<long long>(g) * k * n, dt, n,
        static_cast<long long>(channels) * kk * n, weight + static_cast<long long>(g) * m * k, dt, k, 0, &beta,
        output + static_cast<long long>(g) * m * n, dt, n, static_cast<long long>(channels_out) * n, batch,
        CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (cs != CUBLAS_STATUS_SUCCESS) return B200_ERR_LAUNCH;
  }
  return B200_OK;
}

// ---- INT8 for the shapes the fused tensor-core kernel does not take (groups / deformable groups > 1, other channel
// counts): dequantise into the workspace, run the FP16 gather + cuBLAS path (FP32 accumulation), requantise once.
// A correctness path: every DCN of the BEVFormer backbones goes through dcn_fused_i8 instead.
__global__ void dcn_i8_chw4_to_nchw_f16_kernel(const int8_t *in, float scale, __half *out, int C, long long plane,
                                               long long total) {  // [N, C/4, plane, 4] -> [N, C, plane]
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = idx % plane, nc4 = idx / plane;
    const long long n = nc4 / (C / 4), c4 = nc4 % (C / 4);
    const uint32_t u = ldg32(in + idx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      out[(n * C + c4 * 4 + i) * plane + pix] = __float2half_rn(static_cast<float>(static_cast<int8_t>(u >> (8 * i))) * scale);
  }
}
__global__ void dcn_i8_to_f16_kernel(const int8_t *in, float scale, __half *out, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __float2half_rn(static_cast<float>(in[i]) * scale);
}
__global__ void dcn_bias_to_f16_kernel(const void *bias, int is_half, __half *out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = is_half ? static_cast<const __half *>(bias)[i] : __float2half_rn(static_cast<const float *>(bias)[i]);
}
__global__ void dcn_f16_to_i8_kernel(const __half *in, float inv_scale, int8_t *out, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = static_cast<int8_t>(to_int8_sat(__half2float(in[i]) * inv_scale));
}

static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

struct DcnI8Plan {  // byte offsets into the workspace
  size_t x, w, off, mask, bias, out, col, total;
};
static DcnI8Plan dcn_i8_plan(int batch, int channels, int height, int width, int channels_out, int kh, int kw, int Ho,
                             int Wo, int group, int dg) {
  DcnI8Plan p{};
  size_t o = 0;
  p.x = o, o += align256(static_cast<size_t>(batch) * channels * height * width * 2);
  p.w = o, o += align256(static_cast<size_t>(channels_out) * (channels / group) * kh * kw * 2);
  p.off = o, o += align256(static_cast<size_t>(batch) * dg * 2 * kh * kw * Ho * Wo * 2);
  p.mask = o, o += align256(static_cast<size_t>(batch) * dg * kh * kw * Ho * Wo * 2);
  p.bias = o, o += align256(static_cast<size_t>(channels_out) * 2);
  p.out = o, o += align256(static_cast<size_t>(batch) * channels_out * Ho * Wo * 2);
  p.col = o, o += col_bytes(batch, channels, kh, kw, Ho, Wo, 2);
  p.total = o;
  return p;
}

static unsigned grid_for(long long n) { return static_cast<unsigned>(n / 256 + 1 < (1 << 20) ? n / 256 + 1 : (1 << 20)); }

static int dcn_i8_unfused(const int8_t *input, float scale_i, const int8_t *weight, float scale_w, const void *bias,
                          int bias_is_half, const int8_t *offset, float scale_off, const int8_t *mask, float scale_mask,
                          int8_t *output, float scale_o, void *workspace, int batch, int channels, int height, int width,
                          int channels_out, int kernel_w, int kernel_h, int stride_w, int stride_h, int pad_w, int pad_h,
                          int dilation_w, int dilation_h, int Ho, int Wo, int group, int dg, void *cublas_handle,
                          cudaStream_t s) {
  if (group <= 0 || dg <= 0) return B200_ERR_BAD_PARAM;
  if (channels % 4 || (channels / group) % 4 || channels % group || channels_out % group || channels % dg)
    return B200_ERR_UNSUPPORTED;  // kCHW4 packs 4 channels; the reference exit(1)s on the group mismatches
  const DcnI8Plan pl = dcn_i8_plan(batch, channels, height, width, channels_out, kernel_h, kernel_w, Ho, Wo, group, dg);
  char *ws = static_cast<char *>(workspace);
  __half *x = reinterpret_cast<__half *>(ws + pl.x), *w = reinterpret_cast<__half *>(ws + pl.w);
  __half *off = reinterpret_cast<__half *>(ws + pl.off), *msk = reinterpret_cast<__half *>(ws + pl.mask);
  __half *b = reinterpret_cast<__half *>(ws + pl.bias), *out = reinterpret_cast<__half *>(ws + pl.out);
  const long long plane = static_cast<long long>(height) * width, kk = static_cast<long long>(kernel_h) * kernel_w;
  const long long nx = static_cast<long long>(batch) * (channels / 4) * plane;
  const long long nw = static_cast<long long>(channels_out) * (channels / group / 4) * kk;
  const long long noff = static_cast<long long>(batch) * dg * 2 * kk * Ho * Wo, nout = static_cast<long long>(batch) * channels_out * Ho * Wo;
  dcn_i8_chw4_to_nchw_f16_kernel<<<grid_for(nx), 256, 0, s>>>(input, scale_i, x, channels, plane, nx);
  dcn_i8_chw4_to_nchw_f16_kernel<<<grid_for(nw), 256, 0, s>>>(weight, scale_w, w, channels / group, kk, nw);
  dcn_i8_to_f16_kernel<<<grid_for(noff), 256, 0, s>>>(offset, scale_off, off, noff);
  dcn_i8_to_f16_kernel<<<grid_for(noff / 2), 256, 0, s>>>(mask, scale_mask, msk, noff / 2);
  if (bias) dcn_bias_to_f16_kernel<<<(channels_out + 255) / 256, 256, 0, s>>>(bias, bias_is_half, b, channels_out);
  g_launch_count.fetch_add(bias ? 4 : 3, std::memory_order_relaxed);
  int st = check_launch();
  if (st != B200_OK) return st;
  st = dcn_forward<__half>(x, w, bias ? b : nullptr, off, msk, out, ws + pl.col, batch, channels, height, width,
                           channels_out, kernel_w, kernel_h, stride_w, stride_h, pad_w, pad_h, dilation_w, dilation_h,
                           group, dg, cublas_handle, s);
  if (st != B200_OK) return st;
  dcn_f16_to_i8_kernel<<<grid_for(nout), 256, 0, s>>>(out, 1.f / scale_o, output, nout);
  return check_launch();
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200_dcn_workspace_size(int dtype, int batch, int channels, int height, int width, int kernel_w, int kernel_h,
                               int stride_w, int stride_h, int pad_w, int pad_h, int dilation_w, int dilation_h) {
  if (stride_h <= 0 || stride_w <= 0) return 0;
  const int Ho = (height + 2 * pad_h - (dilation_h * (kernel_h - 1) + 1)) / stride_h + 1;
  const int Wo = (width + 2 * pad_w - (dilation_w * (kernel_w - 1) + 1)) / stride_w + 1;
  if (Ho <= 0 || Wo <= 0 || batch <= 0 || channels <= 0) return 0;
  const size_t v1 = col_bytes(batch, channels, kernel_h, kernel_w, Ho, Wo, dtype == 0 ? 4 : 2);
  // the fused FP16 path needs the NHWC copy of the input + the permuted weights instead (channels_out is not an
  // argument here; 512 is the largest the fused path takes)
  const size_t fused = dtype == 0 ? 0 : dcn_fused_workspace_bytes(batch, channels, height, width, 512, kernel_h * kernel_w);
  return v1 > fused ? v1 : fused;
}

size_t b200_dcn_i8_workspace_size(int batch, int channels, int height, int width, int channels_out, int kernel_w,
                                  int kernel_h, int stride_w, int stride_h, int pad_w, int pad_h, int dilation_w,
                                  int dilation_h, int group, int deformable_group) {
  if (stride_h <= 0 || stride_w <= 0 || group <= 0 || deformable_group <= 0) return 0;
  const int Ho = (height + 2 * pad_h - (dilation_h * (kernel_h - 1) + 1)) / stride_h + 1;
  const int Wo = (width + 2 * pad_w - (dilation_w * (kernel_w - 1) + 1)) / stride_w + 1;
  if (Ho <= 0 || Wo <= 0 || batch <= 0 || channels <= 0 || channels_out <= 0) return 0;
  const size_t fused = b200_dcn_workspace_size(1, batch, channels, height, width, kernel_w, kernel_h, stride_w, stride_h,
                                               pad_w, pad_h, dilation_w, dilation_h);
  const size_t unfused = dcn_i8_plan(batch, channels, height, width, channels_out, kernel_h, kernel_w, Ho, Wo, group,
                                     deformable_group).total;
  return fused > unfused ? fused : unfused;
}

int b200_dcn_set_fused(int enabled) {  // enabled < 0: query only
  const int prev = dcn_fused_enabled() ? 1 : 0;
  if (enabled >= 0) g_dcn_fused = enabled ? 1 : 0;
  return prev;
}

int b200_dcn_pack_weights_f16(const void *weight, void *packed, int channels_out, int channels, int kernel_h, int kernel_w,
                              void *stream) {
  if (!weight || !packed || channels_out <= 0 || channels <= 0 || kernel_h <= 0 || kernel_w <= 0) return B200_ERR_BAD_PARAM;
  if (!dcn_fused_supported(channels, channels_out, kernel_h * kernel_w, 1, 1)) return B200_ERR_UNSUPPORTED;
  return dcn_pack_weights_f16(static_cast<const __half *>(weight), static_cast<__half *>(packed), channels_out, channels,
                              kernel_h * kernel_w, static_cast<cudaStream_t>(stream));
}

int b200_dcn_f16_ex(const void *input, const void *weight, const void *bias, const void *offset, const void *mask,
                    void *output, void *workspace, int batch, int channels, int height, int width, int channels_out,
                    int kernel_w, int kernel_h, int stride_w, int stride_h, int pad_w, int pad_h, int dilation_w,
                    int dilation_h, int group, int deformable_group, int flags, void *stream) {
  if (!input || !weight || !offset || !mask || !output || !workspace) return B200_ERR_BAD_PARAM;
  if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || channels_out <= 0 || kernel_w <= 0 || kernel_h <= 0 ||
      stride_w <= 0 || stride_h <= 0 || dilation_w <= 0 || dilation_h <= 0)
    return B200_ERR_BAD_PARAM;
  const int Ho = (height + 2 * pad_h - (dilation_h * (kernel_h - 1) + 1)) / stride_h + 1;
  const int Wo = (width + 2 * pad_w - (dilation_w * (kernel_w - 1) + 1)) / stride_w + 1;
  if (Ho <= 0 || Wo <= 0) return B200_ERR_BAD_PARAM;
  if (!dcn_fused_supported(channels, channels_out, kernel_h * kernel_w, group, deformable_group) ||
      reinterpret_cast<uintptr_t>(workspace) % 256 != 0 || reinterpret_cast<uintptr_t>(input) % 16 != 0 ||
      reinterpret_cast<uintptr_t>(weight) % 16 != 0)
    return B200_ERR_UNSUPPORTED;
  return dcn_fused_f16(static_cast<const __half *>(input), static_cast<const __half *>(weight),
                       static_cast<const __half *>(bias), static_cast<const __half *>(offset),
                       static_cast<const __half *>(mask), static_cast<__half *>(output), workspace, batch, channels,
                       height, width, channels_out, kernel_w, kernel_h, stride_w, stride_h, pad_w, pad_h, dilation_w,
                       dilation_h, Ho, Wo, flags, static_cast<cudaStream_t>(stream));
}

int b200_dcn_i8(const int8_t *input, float scale_i, const int8_t *weight, float scale_w, const void *bias, int bias_is_half,
                const int8_t *offset, float scale_off, const int8_t *mask, float scale_mask, int8_t *output, float scale_o,
                void *workspace, int batch, int channels, int height, int width, int channels_out, int kernel_w,
                int kernel_h, int stride_w, int stride_h, int pad_w, int pad_h, int dilation_w, int dilation_h, int group,
                int deformable_group, int im2col_step, void *cublas_handle, void *stream) {
  (void)im2col_step;
  if (!input || !weight || !offset || !mask || !output || !workspace) return B200_ERR_BAD_PARAM;
  if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || channels_out <= 0 || kernel_w <= 0 || kernel_h <= 0 ||
      stride_w <= 0 || stride_h <= 0 || dilation_w <= 0 || dilation_h <= 0 || !(scale_o > 0.f))
    return B200_ERR_BAD_PARAM;
  const int Ho = (height + 2 * pad_h - (dilation_h * (kernel_h - 1) + 1)) / stride_h + 1;
  const int Wo = (width + 2 * pad_w - (dilation_w * (kernel_w - 1) + 1)) / stride_w + 1;
  if (Ho <= 0 || Wo <= 0) return B200_ERR_BAD_PARAM;
  if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return B200_ERR_UNSUPPORTED;
  // every DCN of the BEVFormer backbones takes the fused tensor-core path; other shapes dequantise into the workspace
  // (sized by b200_dcn_i8_workspace_size) and run the gather + cuBLAS path
  if (!dcn_fused_enabled() || !dcn_fused_supported(channels, channels_out, kernel_h * kernel_w, group, deformable_group))
    return dcn_i8_unfused(input, scale_i, weight, scale_w, bias, bias_is_half, offset, scale_off, mask, scale_mask,
                          output, scale_o, workspace, batch, channels, height, width, channels_out, kernel_w, kernel_h,
                          stride_w, stride_h, pad_w, pad_h, dilation_w, dilation_h, Ho, Wo, group, deformable_group,
                          cublas_handle, static_cast<cudaStream_t>(stream));
  return dcn_fused_i8(input, scale_i, weight, scale_w, bias, bias_is_half, offset, scale_off, mask, scale_mask, output,
                      scale_o, workspace, batch, channels, height, width, channels_out, kernel_w, kernel_h, stride_w,
                      stride_h, pad_w, pad_h, dilation_w, dilation_h, Ho, Wo, static_cast<cudaStream_t>(stream));
}

int b200_dcn_f32(const float *input, const float *weight, const float *bias, const float *offset, const float *mask,
                 float *output, void *workspace, int batch, int channels, int height, int width, int channels_out,
                 int kernel_w, int kernel_h, int stride_w, int stride_h, int pad_w, int pad_h, int dilation_w,
                 int dilation_h, int group, int deformable_group, int im2col_step, void *cublas_handle, void *stream) {
  (void)im2col_step;  // the reference computes it but loops per image anyway (…Kernel.cu:711-735)
  return dcn_forward<float>(input, weight, bias, offset, mask, output, workspace, batch, channels, height, width,
                            channels_out, kernel_w, kernel_h, stride_w, stride_h, pad_w, pad_h, dilation_w, dilation_h,
                            group, deformable_group, cublas_handle, static_cast<cudaStream_t>(stream));
}

int b200_dcn_f16(const void *input, const void *weight, const void *bias, const void *offset, const void *mask,
                 void *output, void *workspace, int batch, int channels, int height, int width, int channels_out,
                 int kernel_w, int kernel_h, int stride_w, int stride_h, int pad_w, int pad_h, int dilation_w,
                 int dilation_h, int group, int deformable_group, int im2col_step, void *cublas_handle, void *stream) {
  (void)im2col_step;
  return dcn_forward<__half>(static_cast<const __half *>(input), static_cast<const __half *>(weight),
                             static_cast<const __half *>(bias), static_cast<const __half *>(offset),
                             static_cast<const __half *>(mask), static_cast<__half *>(output), workspace, batch,
                             channels, height, width, channels_out, kernel_w, kernel_h, stride_w, stride_h, pad_w,
                             pad_h, dilation_w, dilation_h, group, deformable_group, cublas_handle,
                             static_cast<cudaStream_t>(stream));
}

}  // extern "C"
