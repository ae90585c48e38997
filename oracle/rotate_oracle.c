/*
 * oracle/rotate_oracle.c — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * CPU restatement of the reference RotateTRT plugin's FP32 arithmetic:
 *   TensorRT/plugin/rotate/rotateKernel.cu
 *     :67-70    within_bounds_2d
 *     :73-78    safe_downgrade_to_int_range
 *     :103-109  grid_sampler_compute_source_index: ((coord + 1) * size - 1) / 2
 *     :128-210  rotateKernel<float>: matrix (:137-143), base grid and source index (:145-155), bilinear (:157-188),
 *               nearest with ::round, half away from zero (:189-205)
 * Layout: img / out [C, H, W]; angle in degrees (counter-clockwise); center = (x, y) in pixels.
 *
 * The source index decides which pixel "nearest" copies, so the places where nvcc contracts the reference's
 * expressions into FFMA (read off the SASS of the reference kernel compiled for sm_100a, oracle/_ref) are written
 * with fmaf() here; everything else is separately rounded (compile with -ffp-contract=off). cos / sin are the one
 * part that cannot be restated bit for bit (CUDA's cosf/sinf are accurate to ~1 ulp, not correctly rounded): this
 * file rounds the double-precision value; callers may pass the device's own cos/sin instead (use_trig).
 *
 * Pinning: tests/golden/make_golden_rotate.py runs the reference's own Python binding
 * (det2trt/models/functions/rotate.py:12-84, pure torch: affine grid through bmm + aten.grid_sampler) and stores its
 * outputs; tests/test_oracle_golden.py compares this file against them. The binding builds the grid through a
 * rescaled theta and a batched matmul while the kernel evaluates the matrix per pixel, so source indices differ by
 * rounding (bilinear agrees to ~1e-4 at 512 px; nearest flips a pixel where the index lands within rounding of x.5,
 * which the reference's own test absorbs in a mean-abs tolerance, test_rotate.py:98-105). On the GPU box this file and
 * the product are additionally A/B-checked against the compiled reference kernel (oracle/_ref, ref_rotate).
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>

#include "par.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct {
  float m[4]; /* cos, sin, matrix[2], matrix[5] */
} rot_matrix;

static rot_matrix rot_make_matrix(float angle, float center_x, float center_y, int H, int W, int use_trig, float c_in,
                                  float s_in) {
  rot_matrix r;
  /* -(*angle) * M_PI / 180.f: float * double / float -> double arithmetic, one rounding to float (:137) */
  const float ang = (float)((double)(-angle) * M_PI / 180.0);
  const float c = use_trig ? c_in : (float)cos((double)ang);
  const float s = use_trig ? s_in : (float)sin((double)ang);
  const float cx = center_x - 0.5f * (float)W, cy = center_y - 0.5f * (float)H;
  r.m[0] = c;
  r.m[1] = s;
  r.m[2] = cx + fmaf(-cx, c, -(cy * s)); /* -cx*cos - cy*sin + cx, as contracted by nvcc */
  r.m[3] = cy + fmaf(-cy, c, cx * s);    /*  cx*sin - cy*cos + cy */
  return r;
}

static inline float rot_safe(float x) {
  if (x > (float)(INT_MAX - 1) || x < (float)INT_MIN || !isfinite((double)x)) return -100.f;
  return x;
}

static inline void rot_source_index(const rot_matrix *r, int w, int h, int W, int H, float *ix, float *iy) {
  const float x = -(float)W * 0.5f + 0.5f + (float)w, y = -(float)H * 0.5f + 0.5f + (float)h;
  const float nx = r->m[2] + fmaf(r->m[0], x, r->m[1] * y);
  const float ny = r->m[3] + fmaf(r->m[0], y, -(r->m[1] * x));
  const float gx = nx / (0.5f * (float)W), gy = ny / (0.5f * (float)H);
  *ix = rot_safe(fmaf((float)W, gx + 1.f, -1.f) * 0.5f);
  *iy = rot_safe(fmaf((float)H, gy + 1.f, -1.f) * 0.5f);
}

static inline int inside(int h, int w, int H, int W) { return h >= 0 && h < H && w >= 0 && w < W; }

typedef struct {
  const float *img;
  float *out;
  float *src_xy;
  rot_matrix r;
  int C, H, W, interp;
} rot_ctx;

static void rot_body(long long begin, long long end, void *vctx) {
  const rot_ctx *k = (const rot_ctx *)vctx;
  const int H = k->H, W = k->W;
  const long long plane = (long long)H * W;
  for (long long pix = begin; pix < end; ++pix) {
    const int w = (int)(pix % W), h = (int)(pix / W);
    float ix, iy;
    rot_source_index(&k->r, w, h, W, H, &ix, &iy);
    if (k->src_xy) k->src_xy[2 * pix] = ix, k->src_xy[2 * pix + 1] = iy;
    if (!k->out) continue;
    if (k->interp == 0) {
      const int x0 = (int)floorf(ix), y0 = (int)floorf(iy), x1 = x0 + 1, y1 = y0 + 1;
      const float nw = ((float)x1 - ix) * ((float)y1 - iy), ne = (ix - (float)x0) * ((float)y1 - iy);
      const float sw = ((float)x1 - ix) * (iy - (float)y0), se = (ix - (float)x0) * (iy - (float)y0);
      for (int c = 0; c < k->C; ++c) {
        const float *ip = k->img + c * plane;
        float acc = 0.f; /* "*out += v * w" contracts to one FFMA per tap */
        if (inside(y0, x0, H, W)) acc = fmaf(ip[(long long)y0 * W + x0], nw, acc);
        if (inside(y0, x1, H, W)) acc = fmaf(ip[(long long)y0 * W + x1], ne, acc);
        if (inside(y1, x0, H, W)) acc = fmaf(ip[(long long)y1 * W + x0], sw, acc);
        if (inside(y1, x1, H, W)) acc = fmaf(ip[(long long)y1 * W + x1], se, acc);
        k->out[c * plane + pix] = acc;
      }
    } else {
      const int xn = (int)roundf(ix), yn = (int)roundf(iy);
      const int ok = inside(yn, xn, H, W);
      for (int c = 0; c < k->C; ++c) k->out[c * plane + pix] = ok ? k->img[c * plane + (long long)yn * W + xn] : 0.f;
    }
  }
}

/* out may be NULL (indices only); src_xy may be NULL. use_trig: take cos/sin of the rotation from (c, s). */
void oracle_rotate_f32(const float *img, float angle, float center_x, float center_y, int C, int H, int W, int interp,
                       int use_trig, float c, float s, float *out, float *src_xy) {
  rot_ctx k;
  k.img = img, k.out = out, k.src_xy = src_xy;
  k.r = rot_make_matrix(angle, center_x, center_y, H, W, use_trig, c, s);
  k.C = C, k.H = H, k.W = W, k.interp = interp;
  oracle_parallel_for((long long)H * W, rot_body, &k);
}
