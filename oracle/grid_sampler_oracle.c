/*
 * oracle/grid_sampler_oracle.c — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * CPU restatement of the reference GridSampler2DTRT plugin's FP32 arithmetic:
 *   TensorRT/plugin/grid_sampler/gridSamplerKernel.cu
 *     :82-92    grid_sampler_unnormalize      — grid range is [-10, 10], not [-1, 1]
 *     :157-162  clip_coordinates
 *     :226-247  reflect_coordinates
 *     :342-351  safe_downgrade_to_int_range
 *     :372-393  compute_coordinates
 *     :439-442  within_bounds_2d
 *     :457-466, :509-533 cubic convolution coefficients (A = -0.75), cubic_interp1d
 *     :615-629  get_value_bounded
 *     :666-795  grid_sampler_2d_kernel<float>: bilinear / nearest (::round, half away from zero) / bicubic
 * Layout: input [N, C, Hi, Wi]; grid [N, 2, Ho, Wo] (x plane, then y plane — channel-first, :694-698); out [N,C,Ho,Wo].
 *
 * Pinning: tests/golden/make_golden_grid_sampler.py runs the reference's own Python binding
 * (det2trt/models/functions/grid_sampler.py:7-36, pure torch: aten.grid_sampler(input, grid.permute(0,2,3,1)/10, …))
 * and stores its outputs; tests/test_oracle_golden.py compares this file against them. The binding divides the grid by
 * 10 and then un-normalises from [-1,1] while the kernel un-normalises from [-10,10] directly, so the two differ by
 * rounding in the source index (values agree to ~1e-5 for bilinear/bicubic; nearest can flip at exact half-way points,
 * which the reference's own test tolerates with delta 0.1, test_grid_sampler.py:146-147). Index parity of the product
 * is defined against THIS file (the kernel's formulas), and this file is additionally A/B-checked on the GPU box
 * against the compiled reference kernel (oracle/_ref).
 *
 * Compile with -ffp-contract=off (see msda_oracle.c).
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>

#include "par.h"

typedef struct {
  int32_t ix, iy; /* bilinear/bicubic: floor of the source index (north-west tap); nearest: rounded index */
} gs_index_rec;

static inline float gs_unnormalize(float coord, int size, int align) {
  if (align) return ((coord + 10.f) / 2) * ((float)(size - 1) / 10.f);
  /* "- 0.5" is a double literal in the reference: the float product is widened, 0.5 subtracted exactly, and the
   * result rounded back to float once — identical to a single float subtraction. */
  return (float)((double)(((coord + 10.f) / 2) * ((float)size / 10.f)) - 0.5);
}

static inline float gs_clip(float in, int limit) { return fminf((float)(limit - 1), fmaxf(in, 0.f)); }

static inline float gs_reflect(float in, int twice_low, int twice_high) {
  if (twice_low == twice_high) return 0.f;
  const float mn = (float)twice_low / 2;
  const float span = (float)(twice_high - twice_low) / 2;
  in = fabsf(in - mn);
  const float extra = fmodf(in, span);
  const int flips = (int)floorf(in / span);
  return (flips % 2 == 0) ? extra + mn : span - extra + mn;
}

static inline float gs_safe(float x) {
  if (x > (float)(INT_MAX - 1) || x < (float)INT_MIN || !isfinite((double)x)) return -100.0f;
  return x;
}

static inline float gs_compute_coordinates(float coord, int size, int padding, int align) {
  if (padding == 1) {
    coord = gs_clip(coord, size);
  } else if (padding == 2) {
    coord = align ? gs_reflect(coord, 0, 2 * (size - 1)) : gs_reflect(coord, -1, 2 * size - 1);
    coord = gs_clip(coord, size);
  }
  return gs_safe(coord);
}

static inline float gs_source_index(float coord, int size, int padding, int align) {
  return gs_compute_coordinates(gs_unnormalize(coord, size, align), size, padding, align);
}

static inline int gs_within(int h, int w, int H, int W) { return h >= 0 && h < H && w >= 0 && w < W; }

static inline float cubic1(float x, float A) { return ((A + 2) * x - (A + 3)) * x * x + 1; }
static inline float cubic2(float x, float A) { return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A; }
static inline void cubic_coeffs(float c[4], float t) {
  const float A = -0.75f;
  c[0] = cubic2((float)(t + 1.0), A);
  c[1] = cubic1(t, A);
  const float x2 = (float)(1.0 - t);
  c[2] = cubic1(x2, A);
  c[3] = cubic2((float)(x2 + 1.0), A);
}
static inline float cubic_interp1d(float x0, float x1, float x2, float x3, float t) {
  float c[4];
  cubic_coeffs(c, t);
  return x0 * c[0] + x1 * c[1] + x2 * c[2] + x3 * c[3];
}

static inline float gs_bounded(const float *plane, float x, float y, int W, int H, int padding, int align) {
  x = gs_compute_coordinates(x, W, padding, align);
  y = gs_compute_coordinates(y, H, padding, align);
  const int ix = (int)x, iy = (int)y;
  return gs_within(iy, ix, H, W) ? plane[iy * W + ix] : 0.f;
}

typedef struct {
  const float *input, *grid;
  float *out;
  int N, C, Hi, Wi, Ho, Wo, interp, padding, align;
  gs_index_rec *idx;
} gs_ctx;

static void gs_body(long long begin, long long end, void *vctx) {
  const gs_ctx *x = (const gs_ctx *)vctx;
  const int C = x->C, Hi = x->Hi, Wi = x->Wi, Ho = x->Ho, Wo = x->Wo;
  const long long plane_o = (long long)Ho * Wo, plane_i = (long long)Hi * Wi;
  for (long long p = begin; p < end; ++p) {
    const int w = (int)(p % Wo), h = (int)((p / Wo) % Ho), n = (int)(p / plane_o);
    const float gx = x->grid[((long long)n * 2 + 0) * plane_o + (long long)h * Wo + w];
    const float gy = x->grid[((long long)n * 2 + 1) * plane_o + (long long)h * Wo + w];
    float ix = gs_source_index(gx, Wi, x->padding, x->align);
    float iy = gs_source_index(gy, Hi, x->padding, x->align);
    const float *in_n = x->input + (long long)n * C * plane_i;
    float *out_p = x->out + (long long)n * C * plane_o + (long long)h * Wo + w;
    if (x->interp == 0) { /* bilinear :700-740 */
      const int ix_nw = (int)floorf(ix), iy_nw = (int)floorf(iy);
      const int ix_se = ix_nw + 1, iy_se = iy_nw + 1;
      const float nw = ((float)ix_se - ix) * ((float)iy_se - iy);
      const float ne = (ix - (float)ix_nw) * ((float)iy_se - iy);
      const float sw = ((float)ix_se - ix) * (iy - (float)iy_nw);
      const float se = (ix - (float)ix_nw) * (iy - (float)iy_nw);
      if (x->idx) x->idx[p].ix = ix_nw, x->idx[p].iy = iy_nw;
      for (int c = 0; c < C; ++c) {
        const float *pl = in_n + c * plane_i;
        float o = 0.f;
        if (gs_within(iy_nw, ix_nw, Hi, Wi)) o += pl[iy_nw * Wi + ix_nw] * nw;
        if (gs_within(iy_nw, ix_se, Hi, Wi)) o += pl[iy_nw * Wi + ix_se] * ne;
        if (gs_within(iy_se, ix_nw, Hi, Wi)) o += pl[iy_se * Wi + ix_nw] * sw;
        if (gs_within(iy_se, ix_se, Hi, Wi)) o += pl[iy_se * Wi + ix_se] * se;
        out_p[c * plane_o] = o;
      }
    } else if (x->interp == 1) { /* nearest :741-756 — ::round, i.e. half away from zero */
      const int ixn = (int)roundf(ix), iyn = (int)roundf(iy);
      if (x->idx) x->idx[p].ix = ixn, x->idx[p].iy = iyn;
      for (int c = 0; c < C; ++c)
        out_p[c * plane_o] = gs_within(iyn, ixn, Hi, Wi) ? in_n[c * plane_i + iyn * Wi + ixn] : 0.f;
    } else { /* bicubic :757-792 */
      ix = gs_unnormalize(gx, Wi, x->align);
      iy = gs_unnormalize(gy, Hi, x->align);
      const float ix_nw = floorf(ix), iy_nw = floorf(iy);
      const float tx = ix - ix_nw, ty = iy - iy_nw;
      if (x->idx) x->idx[p].ix = (int)gs_safe(ix_nw), x->idx[p].iy = (int)gs_safe(iy_nw);
      for (int c = 0; c < C; ++c) {
        const float *pl = in_n + c * plane_i;
        float co[4];
        for (int i = 0; i < 4; ++i)
          co[i] = cubic_interp1d(gs_bounded(pl, ix_nw - 1, iy_nw - 1 + i, Wi, Hi, x->padding, x->align),
                                 gs_bounded(pl, ix_nw + 0, iy_nw - 1 + i, Wi, Hi, x->padding, x->align),
                                 gs_bounded(pl, ix_nw + 1, iy_nw - 1 + i, Wi, Hi, x->padding, x->align),
                                 gs_bounded(pl, ix_nw + 2, iy_nw - 1 + i, Wi, Hi, x->padding, x->align), tx);
        out_p[c * plane_o] = cubic_interp1d(co[0], co[1], co[2], co[3], ty);
      }
    }
  }
}

/* interp: 0 bilinear, 1 nearest, 2 bicubic; padding: 0 zeros, 1 border, 2 reflection (gridSamplerKernel.h:11-12) */
void oracle_grid_sample_2d_f32(const float *input, const float *grid, float *out, int N, int C, int Hi, int Wi, int Ho,
                               int Wo, int interp, int padding, int align, gs_index_rec *idx) {
  gs_ctx x = {input, grid, out, N, C, Hi, Wi, Ho, Wo, interp, padding, align, idx};
  oracle_parallel_for((long long)N * Ho * Wo, gs_body, &x);
}
