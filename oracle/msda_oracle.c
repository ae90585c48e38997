/*
 * oracle/msda_oracle.c — TEST INFRASTRUCTURE ONLY. Never linked into, imported by, or called from the
 * product path (bevformer_tensorrt_b200/). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.
 *
 * A plain-C, CPU restatement of the arithmetic of the reference's MultiScaleDeformableAttn TensorRT plugin:
 *
 *   FP32 semantic definition : TensorRT/plugin/multi_scale_deformable_attn/multiScaleDeformableAttnKernel.cu:611-688
 *                              (kernel body) and :133-178 (bilinear tap rule)
 *   INT8 (float ref points)  : same file :849-955 (kernel) and :279-361 (bilinear with quantised intermediates),
 *                              helpers T2int8 :51-55, qmulf :88-94
 *
 * Pinning status: pinned. tests/test_oracle_golden.py checks this file against
 *   (1) vectors produced by importing the reference's own pure-PyTorch implementation
 *       det2trt/models/utils/trt_ops.py:4-85 through the adapter of
 *       det2trt/models/functions/multi_scale_deformable_attn.py:58-92 (tests/golden/make_golden_msda.py), and
 *   (2) vectors produced on a B200 by the reference's own CUDA kernels compiled unmodified into oracle/_ref
 *       (tests/golden/make_golden_ref_gpu.py) — FP32 kernel and both INT8 kernels.
 *
 * Index arithmetic. The reference computes  loc = ref * (float)size + off  in a .cu file compiled with nvcc's
 * default -fmad=true, i.e. as ONE fused multiply-add (single rounding), and then  im = loc - 0.5f  as a separate
 * add (…Kernel.cu:662-672). fmaf() reproduces the fused rounding exactly on the CPU; this file must therefore be
 * compiled with -ffp-contract=off so that the compiler neither fuses nor un-fuses anything else.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "par.h"

#define ORACLE_MAX_C 1024 /* channels per head */

/* One record per (batch, query, head, level*point): what "sampling-index arithmetic" means for the parity tests. */
typedef struct {
  int32_t in_range; /* the gate at …Kernel.cu:674 */
  int32_t h_low;    /* floorf(h_im), only meaningful when in_range */
  int32_t w_low;
  int32_t tap_mask; /* bit0: (h_low,w_low) bit1: (h_low,w_high) bit2: (h_high,w_low) bit3: (h_high,w_high) */
} msda_index_rec;


/* …Kernel.cu:138-172 — the four taps and their validity. Returns the tap mask. */
static inline int tap_rule(int H, int W, float h_im, float w_im, int *h_low_o, int *w_low_o, float wt[4]) {
  const int h_low = (int)floorf(h_im);
  const int w_low = (int)floorf(w_im);
  const int h_high = h_low + 1;
  const int w_high = w_low + 1;
  const float lh = h_im - (float)h_low;
  const float lw = w_im - (float)w_low;
  const float hh = 1.0f - lh, hw = 1.0f - lw;
  int mask = 0;
  if (h_low >= 0 && w_low >= 0) mask |= 1;
  if (h_low >= 0 && w_high <= W - 1) mask |= 2;
  if (h_high <= H - 1 && w_low >= 0) mask |= 4;
  if (h_high <= H - 1 && w_high <= W - 1) mask |= 8;
  wt[0] = hh * hw;
  wt[1] = hh * lw;
  wt[2] = lh * hw;
  wt[3] = lh * lw;
  *h_low_o = h_low;
  *w_low_o = w_low;
  return mask;
}

/*
 * FP32 MSDA in the plugin's launcher signature (…Kernel.h:12-19):
 *   value   [B, S, M, C]   (S = sum_l H_l*W_l, levels concatenated)
 *   shapes  int32 [L, 2]   (h, w)
 *   ref     [B, Q, 1, 2*G] (x, y per group; normalised)
 *   off     [B, Q, M, L*P*2] (pixel units, (x, y) innermost)
 *   logits  [B, Q, M, L*P] (pre-softmax)
 *   out     [B, Q, M, C]
 * idx (optional, may be NULL): [B, Q, M, L*P] records.
 */
typedef struct {
  const float *value, *ref, *off, *logits;
  const int32_t *shapes;
  int B, S, M, C, L, Q, P, G;
  float *out;
  msda_index_rec *idx;
} f32_ctx;

static void f32_body(long long it_begin, long long it_end, void *vctx) {
  const f32_ctx *x = (const f32_ctx *)vctx;
  const float *value = x->value, *ref = x->ref, *off = x->off, *logits = x->logits;
  const int32_t *shapes = x->shapes;
  const int S = x->S, M = x->M, C = x->C, L = x->L, Q = x->Q, P = x->P, G = x->G;
  float *out = x->out;
  msda_index_rec *idx = x->idx;
  const int NP = L * P;
  for (long long it = it_begin; it < it_end; ++it) {
    const int m = (int)(it % M);
    const long long bq = it / M;
    const int q = (int)(bq % Q);
    const int b = (int)(bq / Q);
    const float *lg = logits + it * NP;
    const float *of = off + it * NP * 2;
    const float *rp = ref + ((long long)b * Q + q) * G * 2;
    float *o = out + it * C;
    msda_index_rec *ir = idx ? idx + it * NP : NULL;

    float maxw = -INFINITY; /* …Kernel.cu:642-648 */
    for (int i = 0; i < NP; ++i) maxw = fmaxf(maxw, lg[i]);

    float acc[ORACLE_MAX_C];
    for (int c = 0; c < C; ++c) acc[c] = 0.0f;
    float sumw = 0.0f;

    long long level_start = 0;
    int k = 0;
    for (int l = 0; l < L; ++l) {
      const int H = shapes[2 * l], W = shapes[2 * l + 1];
      const float *vl = value + (((long long)b * S + level_start) * M + m) * C;
      for (int p = 0; p < P; ++p, ++k) {
        const int g = p % G; /* …Kernel.cu:656 */
        const float loc_w = fmaf(rp[2 * g], (float)W, of[2 * k]);
        const float loc_h = fmaf(rp[2 * g + 1], (float)H, of[2 * k + 1]);
        const float e = expf(lg[k] - maxw);
        sumw += e; /* counted even when the point is out of range (:667-669) */
        const float h_im = loc_h - 0.5f;
        const float w_im = loc_w - 0.5f;
        const int inr = (h_im > -1 && w_im > -1 && h_im < (float)H && w_im < (float)W);
        if (ir) {
          ir[k].in_range = inr;
          ir[k].h_low = 0;
          ir[k].w_low = 0;
          ir[k].tap_mask = 0;
        }
        if (!inr) continue;
        int h_low, w_low;
        float wt[4];
        const int mask = tap_rule(H, W, h_im, w_im, &h_low, &w_low, wt);
        if (ir) {
          ir[k].h_low = h_low;
          ir[k].w_low = w_low;
          ir[k].tap_mask = mask;
        }
        const long long step = (long long)M * C;
        const float *t0 = vl + ((long long)h_low * W + w_low) * step;
        const float *t1 = t0 + step;
        const float *t2 = t0 + (long long)W * step;
        const float *t3 = t2 + step;
        for (int c = 0; c < C; ++c) {
          const float v1 = (mask & 1) ? t0[c] : 0.0f;
          const float v2 = (mask & 2) ? t1[c] : 0.0f;
          const float v3 = (mask & 4) ? t2[c] : 0.0f;
          const float v4 = (mask & 8) ? t3[c] : 0.0f;
          /* …Kernel.cu:174-178 then :675-678; nvcc contracts a*b+c chains into FMAs, which changes the value by
           * at most an ulp or two — irrelevant at the value tolerances, so plain arithmetic is used here. */
          const float val = wt[0] * v1 + wt[1] * v2 + wt[2] * v3 + wt[3] * v4;
          acc[c] += val * e;
        }
      }
      level_start += (long long)H * W;
    }
    for (int c = 0; c < C; ++c) o[c] = acc[c] / sumw; /* :686 */
  }
}

void oracle_msda_f32(const float *value, const int32_t *shapes, const float *ref, const float *off,
                     const float *logits, int B, int S, int M, int C, int L, int Q, int P, int G, float *out,
                     msda_index_rec *idx) {
  f32_ctx x = {value, ref, off, logits, shapes, B, S, M, C, L, Q, P, G, out, idx};
  oracle_parallel_for((long long)B * Q * M, f32_body, &x);
}

/* T2int8<float> (…Kernel.cu:51-55): clamp to [-128,127], then round half away from zero by truncation. */
static inline int8_t t2int8f(float a) {
  a = a > 127.0f ? 127.0f : a;
  a = a < -128.0f ? -128.0f : a;
  return (int8_t)(a + (a > 0 ? 0.5f : -0.5f));
}

/*
 * "In-register dequant" INT8 definition (what the product computes; SURVEY Appendix A.4 "parity plan, primary"):
 * the FP32 formulas of oracle_msda_f32 evaluated on value*scale_value, off*scale_offset, logits*scale_weight,
 * with the result requantised by T2int8(result / scale_out). ref points are given as float (callers convert fp16).
 * out_real (optional) receives the un-quantised float result.
 */
typedef struct {
  const int8_t *value, *off, *logits;
  const float *ref;
  const int32_t *shapes;
  float scale_value, scale_offset, scale_weight, scale_out;
  int B, S, M, C, L, Q, P, G;
  int8_t *out;
  float *out_real;
} i8_ctx;

#define I8_UNPACK                                                                                                   \
  const i8_ctx *x = (const i8_ctx *)vctx;                                                                           \
  const int8_t *value = x->value, *off = x->off, *logits = x->logits;                                               \
  const float *ref = x->ref;                                                                                        \
  const int32_t *shapes = x->shapes;                                                                                \
  const float scale_value = x->scale_value, scale_offset = x->scale_offset, scale_weight = x->scale_weight,         \
              scale_out = x->scale_out;                                                                             \
  const int S = x->S, M = x->M, C = x->C, L = x->L, Q = x->Q, P = x->P, G = x->G;                                   \
  int8_t *out = x->out;                                                                                             \
  float *out_real = x->out_real;                                                                                    \
  (void)out_real; (void)scale_value; (void)scale_out;                                                               \
  const int NP = L * P;

static void i8_dequant_body(long long it_begin, long long it_end, void *vctx) {
  I8_UNPACK
  for (long long it = it_begin; it < it_end; ++it) {
    const int m = (int)(it % M);
    const long long bq = it / M;
    const int q = (int)(bq % Q);
    const int b = (int)(bq / Q);
    const int8_t *lg = logits + it * NP;
    const int8_t *of = off + it * NP * 2;
    const float *rp = ref + ((long long)b * Q + q) * G * 2;

    float maxw = -INFINITY;
    for (int i = 0; i < NP; ++i) maxw = fmaxf(maxw, (float)lg[i] * scale_weight);

    float acc[ORACLE_MAX_C];
    for (int c = 0; c < C; ++c) acc[c] = 0.0f;
    float sumw = 0.0f;
    long long level_start = 0;
    int k = 0;
    for (int l = 0; l < L; ++l) {
      const int H = shapes[2 * l], W = shapes[2 * l + 1];
      const int8_t *vl = value + (((long long)b * S + level_start) * M + m) * C;
      for (int p = 0; p < P; ++p, ++k) {
        const int g = p % G;
        /* …Kernel.cu:916-921: ref*W + (int8 off * scale_offset); the product off*scale is rounded to float first,
         * then fused with ref*W. */
        const float ox = (float)of[2 * k] * scale_offset;
        const float oy = (float)of[2 * k + 1] * scale_offset;
        const float loc_w = fmaf(rp[2 * g], (float)W, ox);
        const float loc_h = fmaf(rp[2 * g + 1], (float)H, oy);
        const float e = expf((float)lg[k] * scale_weight - maxw);
        sumw += e;
        const float h_im = loc_h - 0.5f;
        const float w_im = loc_w - 0.5f;
        if (!(h_im > -1 && w_im > -1 && h_im < (float)H && w_im < (float)W)) continue;
        int h_low, w_low;
        float wt[4];
        const int mask = tap_rule(H, W, h_im, w_im, &h_low, &w_low, wt);
        const long long step = (long long)M * C;
        const int8_t *t0 = vl + ((long long)h_low * W + w_low) * step;
        const int8_t *t1 = t0 + step;
        const int8_t *t2 = t0 + (long long)W * step;
        const int8_t *t3 = t2 + step;
        for (int c = 0; c < C; ++c) {
          const float v1 = (mask & 1) ? (float)t0[c] : 0.0f;
          const float v2 = (mask & 2) ? (float)t1[c] : 0.0f;
          const float v3 = (mask & 4) ? (float)t2[c] : 0.0f;
          const float v4 = (mask & 8) ? (float)t3[c] : 0.0f;
          const float val = wt[0] * v1 + wt[1] * v2 + wt[2] * v3 + wt[3] * v4;
          acc[c] += val * e;
        }
      }
      level_start += (long long)H * W;
    }
    for (int c = 0; c < C; ++c) {
      const float real = acc[c] * scale_value / sumw;
      if (out_real) out_real[it * C + c] = real;
      out[it * C + c] = t2int8f(real / scale_out);
    }
  }
}

void oracle_msda_i8_dequant(const int8_t *value, float scale_value, const int32_t *shapes, const float *ref,
                            const int8_t *off, float scale_offset, const int8_t *logits, float scale_weight, int B,
                            int S, int M, int C, int L, int Q, int P, int G, int8_t *out, float scale_out,
                            float *out_real) {
  i8_ctx x = {value, off, logits, ref, shapes, scale_value, scale_offset, scale_weight, scale_out,
              B, S, M, C, L, Q, P, G, out, out_real};
  oracle_parallel_for((long long)B * Q * M, i8_dequant_body, &x);
}

/*
 * Emulation of the reference's INT8 kernel with float reference points (…Kernel.cu:849-955, :279-361):
 * quantised intermediates — softmax weights as int8 at x127, bilinear weights as int8 at x127, per-point sampled
 * value requantised to int8 in value scale, int32 accumulation, final  T2int8(acc * (scale_value/scale_out) / sumw).
 * Requires P % 4 == 0 (…Plugin.cpp:151-156). Used to check oracle/_ref's INT8 kernel and to quantify how far the
 * in-register-dequant definition is from the reference's quantised-intermediate result.
 */
static void i8_refemu_body(long long it_begin, long long it_end, void *vctx) {
  I8_UNPACK
  /* scale_o = scale_value * __frcp_rn(scale_out) (:879) — correctly rounded reciprocal, then one multiply. */
  const float scale_o = scale_value * (1.0f / scale_out);
  for (long long it = it_begin; it < it_end; ++it) {
    const int m = (int)(it % M);
    const long long bq = it / M;
    const int q = (int)(bq % Q);
    const int b = (int)(bq / Q);
    const int8_t *lg = logits + it * NP;
    const int8_t *of = off + it * NP * 2;
    const float *rp = ref + ((long long)b * Q + q) * G * 2;

    float maxw = -INFINITY;
    for (int i = 0; i < NP; ++i) maxw = fmaxf(maxw, (float)lg[i] * scale_weight);

    int32_t acc[ORACLE_MAX_C];
    for (int c = 0; c < C; ++c) acc[c] = 0;
    float sumw = 0.0f;
    long long level_start = 0;
    int k = 0;
    for (int l = 0; l < L; ++l) {
      const int H = shapes[2 * l], W = shapes[2 * l + 1];
      const int8_t *vl = value + (((long long)b * S + level_start) * M + m) * C;
      for (int p = 0; p < P; ++p, ++k) {
        const int g = p % G;
        const float ox = (float)of[2 * k] * scale_offset;
        const float oy = (float)of[2 * k + 1] * scale_offset;
        const float loc_w = fmaf(rp[2 * g], (float)W, ox);
        const float loc_h = fmaf(rp[2 * g + 1], (float)H, oy);
        const float h_im = loc_h - 0.5f;
        const float w_im = loc_w - 0.5f;
        const int8_t wq = t2int8f(expf((float)lg[k] * scale_weight - maxw) * 127.0f); /* :926-929 */
        sumw += (float)wq;                                                                /* :930 */
        if (!(h_im > -1 && w_im > -1 && h_im < (float)H && w_im < (float)W)) continue;
        int h_low, w_low;
        float wt[4];
        const int mask = tap_rule(H, W, h_im, w_im, &h_low, &w_low, wt);
        const float scale_area = 1.0f / 127.0f; /* :298 */
        const int8_t bw[4] = {t2int8f(wt[0] / scale_area), t2int8f(wt[1] / scale_area), t2int8f(wt[2] / scale_area),
                              t2int8f(wt[3] / scale_area)};
        const long long step = (long long)M * C;
        const int8_t *t0 = vl + ((long long)h_low * W + w_low) * step;
        const int8_t *t1 = t0 + step;
        const int8_t *t2 = t0 + (long long)W * step;
        const int8_t *t3 = t2 + step;
        for (int c = 0; c < C; ++c) {
          int32_t t = 0;
          if (mask & 1) t += (int32_t)t0[c] * bw[0];
          if (mask & 2) t += (int32_t)t1[c] * bw[1];
          if (mask & 4) t += (int32_t)t2[c] * bw[2];
          if (mask & 8) t += (int32_t)t3[c] * bw[3];
          const int8_t sv = t2int8f((float)t * scale_area); /* :346-358 */
          acc[c] += (int32_t)sv * (int32_t)wq;              /* :944-947 */
        }
      }
      level_start += (long long)H * W;
    }
    /* qmulf(output, data_output, scale_o * __frcp_rn(sum_weight)) (:951-953) */
    const float mul = scale_o * (1.0f / sumw);
    for (int c = 0; c < C; ++c) out[it * C + c] = t2int8f((float)acc[c] * mul);
  }
}

void oracle_msda_i8_refemu(const int8_t *value, float scale_value, const int32_t *shapes, const float *ref,
                           const int8_t *off, float scale_offset, const int8_t *logits, float scale_weight, int B,
                           int S, int M, int C, int L, int Q, int P, int G, int8_t *out, float scale_out) {
  i8_ctx x = {value, off, logits, ref, shapes, scale_value, scale_offset, scale_weight, scale_out,
              B, S, M, C, L, Q, P, G, out, NULL};
  oracle_parallel_for((long long)B * Q * M, i8_refemu_body, &x);
}
