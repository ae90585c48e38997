// msda_v2.cu — second-generation multi-scale deformable attention for B200 (FP16 and INT8, C = 32, 16 <= L*P <= 32).
//
// Why a second kernel (profiles/r02*_micro_gather_bw2.txt, profiles/r01_msda_f16_U_ncu_summary.txt): the round-1 kernel
// (msda.cu) gathers each bilinear tap of a head as its own 64-byte (FP16) / 32-byte (INT8) piece out of the reference's
// value layout [S, heads, 32]. On B200 the L1/LSU path retires about ONE distinct 128-byte line per clock per SM whatever
// the piece size (LSU write-back 86 % busy at 1.03 ms), and the SIMT accumulate spends 8 HADD2.F32 widenings + 4 FFMA2 per
// 16 bytes. Both limits are functions of the LAYOUT, so this path changes the layout first:
//
//   1. PACK pre-pass (msda_pack_kernel): `value` is re-laid per (camera, head) as COLUMN-PAIR ENTRIES
//          entry(level, parity, r, x) = { pixel(y0, x), pixel(y0 + 1, x) },  y0 = 2r - parity,
//      with the two rows interleaved per channel: [ch0:y0, ch0:y1, ch1:y0, ch1:y1, ...] (out-of-image rows are zeros).
//      A bilinear sample (h_low, w_low) needs exactly the two entries (r, w_low), (r, w_low + 1) of parity h_low & 1:
//      FP16 2 x 128 B = two whole lines, INT8 2 x 64 B = one 128-byte run. Levels get the second (odd) parity copy
//      from the coarsest level down while the packed stack stays within an L2 budget; a level without it serves odd
//      h_low from two entries per column with one zero weight each. Value tiles enter shared memory by TMA bulk
//      copies (cp.async.bulk, SASS UBLKCP: rows of the [S, heads*32] stack are contiguous), are interleaved by the
//      threads and leave as 16-byte coalesced stores.
//   2. GATHER kernels, one item (batch, query, head) per warp at a time, lane = sampling point in the prologue
//      (32 points -> 32 lanes: bit-exact index arithmetic of …Kernel.cu:657-674 / :138-172, softmax by warp shuffles),
//      which emits a COMPACTED list of entry fetches (points that are out of range emit nothing) into shared memory:
//        FP16: the interleaved entry IS the A fragment of mma.sync.m16n8k16 (rows = channels, k = taps): each lane
//              loads 16 B = 4 channels x (y0, y1) and feeds two MMAs untouched; the tap weights ride in two B columns
//              as an fp16 hi/lo split (22-bit weights, fp32 accumulate: exact to ~2^-22 — no FP16 widening, no FFMA).
//        INT8: the interleaved bytes are dp2a operands: b = {ch_i:y0, ch_i:y1, ch_j:y0, ch_j:y1}, a = two 16-bit fixed
//              point weights; 8 IDP per 16 bytes instead of 36 convert/FMA slots; int32 accumulate, one requantisation.
//      Semantics are those of msda.cu (FP32 index math, softmax over all L*P logits, zero-padded taps).
//
// Replaces (same as msda.cu): ms_deformable_im2col_cuda<__half>, _h2, _int8<float|__half2>
// (TensorRT/plugin/multi_scale_deformable_attn/multiScaleDeformableAttnKernel.cu:1130-1218, kernels :691-1104).
#include <cuda.h>

#include "common.cuh"

namespace b200 {

constexpr int kV2MaxLevels = 8;
constexpr int kV2Threads = 256;
constexpr int kV2Warps = kV2Threads / 32;
constexpr int kPackCols = 64;  // pixels per pack tile row


// The packed-stack plan is a pure function of (spatial_shapes, batch*heads, entry bytes, L2 budget) and is evaluated ON
// THE DEVICE by both kernels from the same device-resident spatial_shapes the reference plugin receives (TensorRT hands
// the plugin device pointers only), so the host never needs the level sizes. Parity-1 (odd-row) copies are granted
// from the coarsest level down while the whole stack stays within `budget` bytes (L2 residency).
struct LevelInfo {
  int H, W;
  int pix0;     // first pixel of the level inside S
  int e0, e1;   // first entry of the parity-0 / parity-1 block inside a (camera, head) slab; e1 < 0: no odd copy
  int entries;  // entries per (camera, head) slab (all levels)
  int tiles0;   // pack tiles (row pair x 64-column chunk x parity) before this level, and of this level
  int tiles;
};


__device__ __forceinline__ long long level_block(int H, int W) { return static_cast<long long>(H / 2 + 1) * W; }

__device__ LevelInfo plan_level(const int32_t *shapes, int L, long long bm, int EB, long long budget, int want) {
  long long total = 0;
  for (int l = 0; l < L; ++l) total += level_block(__ldg(shapes + 2 * l), __ldg(shapes + 2 * l + 1));
  unsigned dup = 0u;
  for (int l = L - 1; l >= 0; --l) {
    const long long blk = level_block(__ldg(shapes + 2 * l), __ldg(shapes + 2 * l + 1));
    if ((total + blk) * EB * bm <= budget) dup |= 1u << l, total += blk;
  }
  LevelInfo r{};
  long long e = 0;
  int pix = 0, tiles = 0;
  for (int l = 0; l < L; ++l) {
    const int H = __ldg(shapes + 2 * l), W = __ldg(shapes + 2 * l + 1);
    const long long blk = level_block(H, W);
    const bool d = (dup >> l) & 1u;
    const int nt = (H / 2 + 1) * ((W + kPackCols - 1) / kPackCols) * (d ? 2 : 1);
    if (l == want) {
      r.H = H, r.W = W, r.pix0 = pix, r.e0 = static_cast<int>(e), r.e1 = d ? static_cast<int>(e + blk) : -1;
      r.tiles0 = tiles, r.tiles = nt;
    }
    e += blk * (d ? 2 : 1), pix += H * W, tiles += nt;
  }
  r.entries = static_cast<int>(e);
  if (want >= L) r.tiles0 = tiles;  // total tile count
  return r;
}

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---------------------------------------------------------------------------------------------------------------
// 1. pack pre-pass
// ---------------------------------------------------------------------------------------------------------------
struct PackParams {
  const char *value;  // [B, S, M*C] elements of EL bytes
  char *packed;       // [B, M, entries, 2*C*EL]
  const int32_t *shapes;
  long long budget;
  int B, S, M, L;
};

// One tile = one (level, parity, row pair r, 64-column chunk) of one camera: two rows x 64 pixels x (M*C*EL) bytes arrive
// in shared memory by two TMA bulk copies (the pixels of a row chunk are contiguous in `value`), then every thread builds
// 16-byte pieces of the interleaved entries. Persistent CTAs loop over the tiles (their number is known on the device only).
template <int EL>  // bytes per element: 2 (fp16) or 1 (int8); C = 32
__global__ void __launch_bounds__(256) msda_pack_kernel(const PackParams p) {
  constexpr int C = 32;
  constexpr int EB = 2 * C * EL;    // entry bytes: 128 (fp16) / 64 (int8)
  extern __shared__ __align__(128) char tile[];  // [2][kPackCols][M*C*EL]
  __shared__ __align__(8) unsigned long long bar;
  const int b = blockIdx.y;
  const long long bm = static_cast<long long>(p.B) * p.M;
  const int tiles_total = plan_level(p.shapes, p.L, bm, EB, p.budget, p.L).tiles0;
  const uint32_t barr = smem_addr(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barr) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;
  for (int tidx = blockIdx.x; tidx < tiles_total; tidx += gridDim.x) {
  // decode the tile index -> (level, parity, r, chunk)
  int l = 0;
  LevelInfo lv = plan_level(p.shapes, p.L, bm, EB, p.budget, 0);
  while (l + 1 < p.L && tidx >= lv.tiles0 + lv.tiles) lv = plan_level(p.shapes, p.L, bm, EB, p.budget, ++l);
  const int t = tidx - lv.tiles0;
  const int chunks = (lv.W + kPackCols - 1) / kPackCols, per_par = (lv.H / 2 + 1) * chunks;
  const int par = t / per_par, r = (t % per_par) / chunks, chunk = t % chunks;
  const int x0 = chunk * kPackCols, ncol = min(kPackCols, lv.W - x0);
  const int pix_bytes = p.M * C * EL;
  const int row_bytes = ncol * pix_bytes;
  const int y0 = 2 * r - par;
  bool valid[2];
#pragma unroll
  for (int rho = 0; rho < 2; ++rho) valid[rho] = (y0 + rho >= 0) && (y0 + rho < lv.H);
  if (threadIdx.x == 0) {
    // the tile buffer was last touched through the generic proxy (zero fill / reads of the previous tile): order those
    // accesses before the async-proxy writes of the bulk copies
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t tx = (valid[0] ? row_bytes : 0) + (valid[1] ? row_bytes : 0);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barr), "r"(tx) : "memory");
#pragma unroll
    for (int rho = 0; rho < 2; ++rho) {
      if (!valid[rho]) continue;
      const char *src = p.value + (static_cast<long long>(b) * p.S + lv.pix0 + static_cast<long long>(y0 + rho) * lv.W + x0) * pix_bytes;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_addr(tile + rho * kPackCols * pix_bytes)),
                   "l"(src), "r"(row_bytes), "r"(barr)
                   : "memory");
    }
  }
  // rows outside the image are zeros (the gather never weights them, but they must be finite)
#pragma unroll
  for (int rho = 0; rho < 2; ++rho)
    if (!valid[rho])
      for (int i = threadIdx.x; i < row_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(tile + rho * kPackCols * pix_bytes)[i] = make_uint4(0u, 0u, 0u, 0u);
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                   : "=r"(done) : "r"(barr), "r"(phase) : "memory");
    phase ^= 1u;
  }
  __syncthreads();

  constexpr int CPE = EB / 16;      // 16-byte pieces per entry: 8 / 4
  constexpr int HALF = 8;           // bytes of one row inside a 16-byte piece
  const int pieces = p.M * ncol * CPE;
  const long long slab = static_cast<long long>(lv.entries) * EB;
  const long long ent = static_cast<long long>(par ? lv.e1 : lv.e0) + static_cast<long long>(r) * lv.W + x0;
  for (int i = threadIdx.x; i < pieces; i += blockDim.x) {
    const int j = i % CPE, x = (i / CPE) % ncol, m = i / (CPE * ncol);
    const char *s0 = tile + x * pix_bytes + m * C * EL + j * HALF;
    const uint2 a = *reinterpret_cast<const uint2 *>(s0);                             // row y0: 8 bytes of channels
    const uint2 c = *reinterpret_cast<const uint2 *>(s0 + kPackCols * pix_bytes);     // row y0+1, same channels
    uint4 o;
    if (EL == 2) {  // halves: (a0 c0)(a1 c1)(a2 c2)(a3 c3)
      o.x = __byte_perm(a.x, c.x, 0x5410), o.y = __byte_perm(a.x, c.x, 0x7632);
      o.z = __byte_perm(a.y, c.y, 0x5410), o.w = __byte_perm(a.y, c.y, 0x7632);
    } else {  // bytes: (a0 c0 a1 c1)(a2 c2 a3 c3)...
      o.x = __byte_perm(a.x, c.x, 0x5140), o.y = __byte_perm(a.x, c.x, 0x7362);
      o.z = __byte_perm(a.y, c.y, 0x5140), o.w = __byte_perm(a.y, c.y, 0x7362);
    }
    char *dst = p.packed + (static_cast<long long>(b) * p.M + m) * slab + (ent + x) * EB + j * 16;
    *reinterpret_cast<uint4 *>(dst) = o;
  }
  __syncthreads();  // the tile buffer is reused by the next bulk copies
  }  // tiles
}

// ---------------------------------------------------------------------------------------------------------------
// 2. gather
// ---------------------------------------------------------------------------------------------------------------
struct V2Params {
  const char *packed;
  const void *ref, *off, *logits;
  void *out;
  const int32_t *shapes;
  long long budget;
  int B, M, Q, P, G, NP, L;
  long long items;
  int ref_is_half;
  const float *mask;  // EPI 1
  float *accum;
  float scale_value, scale_offset, scale_weight, scale_out;
  int4 *trace;
};

__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ int dp2a_lo(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, d));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFullMask, v, d);
  return v;
}

// fp16 hi/lo split of two weights: returns (hi pair, lo pair), each packed (y0 in the low half, y1 in the high half)
__device__ __forceinline__ uint2 split_pair(float w0, float w1) {
  const __half h0 = __float2half_rn(w0), h1 = __float2half_rn(w1);
  const __half l0 = __float2half_rn(w0 - __half2float(h0)), l1 = __float2half_rn(w1 - __half2float(h1));
  const __half2 hi = __halves2half2(h0, h1), lo = __halves2half2(l0, l1);
  return make_uint2(*reinterpret_cast<const uint32_t *>(&hi), *reinterpret_cast<const uint32_t *>(&lo));
}
__device__ __forceinline__ uint32_t fixed_pair(float w0, float w1) {
  const uint32_t q0 = min(__float2uint_rn(w0 * 65536.f), 65535u), q1 = min(__float2uint_rn(w1 * 65536.f), 65535u);
  return q0 | (q1 << 16);
}

// T: __half or int8_t (storage of value / offsets / logits / out). R: reference-point storage (__half / float).
// EPI 0: plugin output; EPI 1 (FP16 only): accum[q, m*32 + c] += bev_mask[b, q] * out (fused SCA sampling).
template <typename T, typename R, int EPI, bool DBG>
__global__ void __launch_bounds__(kV2Threads, 4) msda_v2_kernel(const V2Params prm) {
  constexpr bool I8 = sizeof(T) == 1;
  constexpr int EB = I8 ? 64 : 128;
  constexpr int MAXREC = 32 * 4 + 8;  // worst case: every point on a level without the odd copy, plus padding
  __shared__ uint4 recs[kV2Warps][MAXREC];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int M = prm.M, Q = prm.Q, P = prm.P, G = prm.G, NP = prm.NP;
  uint4 *rec = recs[warp];

  // per-lane constants of "my" sampling point (lane = point index inside the item)
  const bool have_pt = lane < NP;
  const int pt = have_pt ? lane : 0;
  const int lvl = pt / P, grp = (pt % P) % G;
  const LevelInfo lv = plan_level(prm.shapes, prm.L, static_cast<long long>(prm.B) * M, EB, prm.budget, lvl);
  const float Hf = static_cast<float>(lv.H), Wf = static_cast<float>(lv.W);
  const long long slab = static_cast<long long>(lv.entries) * EB;

  const unsigned warps_total = gridDim.x * kV2Warps, items = static_cast<unsigned>(prm.items);  // items < 2^31 (host)
  for (unsigned it32 = blockIdx.x * kV2Warps + warp; it32 < items; it32 += warps_total) {
    const long long it = it32;
    const unsigned bq32 = it32 / static_cast<unsigned>(M);
    const long long bq = bq32;
    const int m = static_cast<int>(it32 - bq32 * M);
    const int b = static_cast<int>(bq32 / static_cast<unsigned>(Q));
    float mk = 1.f;
    if (EPI == 1) {
      mk = __ldg(prm.mask + bq);
      if (mk == 0.f) continue;  // warp-uniform: this camera does not see the query, nothing is added
    }
    // ---- phase A (bit-exact with the reference): loc = fma(ref, size, off) - 0.5, range gate
    float ox, oy, rx, ry;
    if (I8) {
      const unsigned short o2 = __ldg(reinterpret_cast<const unsigned short *>(prm.off) + it * NP + pt);
      ox = static_cast<float>(static_cast<int8_t>(o2 & 0xff)) * prm.scale_offset;
      oy = static_cast<float>(static_cast<int8_t>(o2 >> 8)) * prm.scale_offset;
    } else {
      const float2 o2 = h2_to_f2(ldg32_stream(reinterpret_cast<const uint32_t *>(prm.off) + it * NP + pt));
      ox = o2.x, oy = o2.y;
    }
    if (sizeof(R) == 2) {
      const float2 r2 = h2_to_f2(__ldg(reinterpret_cast<const uint32_t *>(prm.ref) + bq * G + grp));
      rx = r2.x, ry = r2.y;
    } else {
      const float2 r2 = __ldg(reinterpret_cast<const float2 *>(prm.ref) + bq * G + grp);
      rx = r2.x, ry = r2.y;
    }
    const float w_im = __fadd_rn(__fmaf_rn(rx, Wf, ox), -0.5f);
    const float h_im = __fadd_rn(__fmaf_rn(ry, Hf, oy), -0.5f);
    const bool ok = have_pt && h_im > -1.f && w_im > -1.f && h_im < Hf && w_im < Wf;
    const unsigned okm = __ballot_sync(kFullMask, ok);
    T *out_item = static_cast<T *>(prm.out) + it * 32;
    if (okm == 0u) {  // nothing in range: the result is exactly 0 (= 0 / sum); logits are never read
      if (EPI == 0) {
        if (I8) { if (lane < 8) reinterpret_cast<uint32_t *>(out_item)[lane] = 0u; }
        else { if (lane < 16) reinterpret_cast<uint32_t *>(out_item)[lane] = 0u; }
      }
      continue;
    }
    // ---- phase B: softmax statistics over all NP logits
    float lg = -INFINITY;
    if (have_pt) {
      if (I8) lg = static_cast<float>(__ldg(reinterpret_cast<const int8_t *>(prm.logits) + it * NP + pt)) * prm.scale_weight;
      else lg = __half2float(__ldg(reinterpret_cast<const __half *>(prm.logits) + it * NP + pt));
    }
    const float mx = warp_max(lg);
    float e = expf(lg - mx);  // exp(-inf) = 0 for lanes beyond NP
    const float sum = warp_sum(e);
    if (EPI == 1) e *= mk;  // bev_mask folded into the tap weights

    // ---- phase C: taps -> fetch records
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int h_low = ok ? static_cast<int>(hf) : 0, w_low = ok ? static_cast<int>(wf) : 0;
    const float lh = __fsub_rn(h_im, hf), lw = __fsub_rn(w_im, wf), hh = 1.f - lh, hw = 1.f - lw;
    const bool t = h_low >= 0, bt = h_low + 1 <= lv.H - 1, lf = w_low >= 0, rt = w_low + 1 <= lv.W - 1;
    const float w00 = (ok && t && lf) ? hh * hw * e : 0.f, w01 = (ok && t && rt) ? hh * lw * e : 0.f;
    const float w10 = (ok && bt && lf) ? lh * hw * e : 0.f, w11 = (ok && bt && rt) ? lh * lw * e : 0.f;
    if (DBG && have_pt)
      prm.trace[it * NP + pt] = ok ? make_int4(1, h_low, w_low, ((t && lf) ? 1 : 0) | ((t && rt) ? 2 : 0) |
                                                                    ((bt && lf) ? 4 : 0) | ((bt && rt) ? 8 : 0))
                                   : make_int4(0, 0, 0, 0);
    const int par = h_low & 1;
    const bool split = ok && par == 1 && lv.e1 < 0;  // odd h_low on a level without the odd-parity copy
    const int r = (h_low + 1) >> 1;
    const int xl = max(w_low, 0), xr = min(w_low + 1, lv.W - 1);
    // entries per point: FP16 records are per ENTRY (2, or 4 when split); INT8 records are per SAMPLE (1, or 2)
    const int cnt = ok ? (I8 ? (split ? 2 : 1) : (split ? 4 : 2)) : 0;
    int pos = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(kFullMask, pos, d);
      if (lane >= d) pos += v;
    }
    const int total = __shfl_sync(kFullMask, pos, 31);
    pos -= cnt;
    constexpr int STEP = I8 ? 4 : 8;  // records consumed per loop iteration
    const int steps = (total + STEP - 1) / STEP;
    if (ok) {
      if (!split) {
        const unsigned base = static_cast<unsigned>(((lv.e1 >= 0 && par) ? lv.e1 : lv.e0) + r * lv.W);
        const unsigned aL = (base + xl) * EB, aR = (base + xr) * EB;
        if (I8) {
          rec[pos] = make_uint4(aL, fixed_pair(w00, w10), aR, fixed_pair(w01, w11));
        } else {
          const uint2 L = split_pair(w00, w10), Rr = split_pair(w01, w11);
          rec[pos] = make_uint4(aL, L.x, L.y, 0u), rec[pos + 1] = make_uint4(aR, Rr.x, Rr.y, 0u);
        }
      } else {
        // rows h_low (= y1 of even entry r-1) and h_low+1 (= y0 of even entry r)
        const unsigned top = static_cast<unsigned>(lv.e0 + max(r - 1, 0) * lv.W);
        const unsigned bot = static_cast<unsigned>(lv.e0 + r * lv.W);
        const unsigned tL = (top + xl) * EB, tR = (top + xr) * EB, bL = (bot + xl) * EB, bR = (bot + xr) * EB;
        if (I8) {
          rec[pos] = make_uint4(tL, fixed_pair(0.f, w00), tR, fixed_pair(0.f, w01));
          rec[pos + 1] = make_uint4(bL, fixed_pair(w10, 0.f), bR, fixed_pair(w11, 0.f));
        } else {
          const uint2 a = split_pair(0.f, w00), c = split_pair(0.f, w01), d = split_pair(w10, 0.f), f = split_pair(w11, 0.f);
          rec[pos] = make_uint4(tL, a.x, a.y, 0u), rec[pos + 1] = make_uint4(tR, c.x, c.y, 0u);
          rec[pos + 2] = make_uint4(bL, d.x, d.y, 0u), rec[pos + 3] = make_uint4(bR, f.x, f.y, 0u);
        }
      }
    }
    if (lane < STEP && total + lane < steps * STEP) rec[total + lane] = make_uint4(0u, 0u, 0u, 0u);  // zero-weight padding
    __syncwarp();

    const char *vslab = prm.packed + (static_cast<long long>(b) * M + m) * slab;
    if (!I8) {
      // ---- FP16: tensor-core accumulate. Lane (g = lane>>2, tg = lane&3) loads 16 B = channels 4g..4g+3 x (y0,y1) of
      // entry records tg and tg+4 of the step; MMA #1 takes channels (4g, 4g+1) as rows (g, g+8), MMA #2 (4g+2, 4g+3).
      const int g = lane >> 2, tg = lane & 3;
      float d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
      const char *vl = vslab + g * 16;
#pragma unroll 2
      for (int s = 0; s < steps; ++s) {
        const uint4 r0 = rec[s * 8 + tg], r1 = rec[s * 8 + 4 + tg];
        const uint4 a = ldg128(vl + r0.x), c = ldg128(vl + r1.x);
        const uint32_t b0 = g == 0 ? r0.y : (g == 1 ? r0.z : 0u), b1 = g == 0 ? r1.y : (g == 1 ? r1.z : 0u);
        mma_f16(d1, a.x, a.y, c.x, c.y, b0, b1);
        mma_f16(d2, a.z, a.w, c.z, c.w, b0, b1);
      }
      __syncwarp();
      if (tg == 0) {  // columns 0 (hi) and 1 (lo) of D live in the lanes with tg == 0
        const float sc = 1.f / sum;  // EPI 1: bev_mask is already inside the tap weights
        const float o0 = (d1[0] + d1[1]) * sc, o1 = (d1[2] + d1[3]) * sc, o2 = (d2[0] + d2[1]) * sc, o3 = (d2[2] + d2[3]) * sc;
        if (EPI == 0) {
          asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(reinterpret_cast<__half *>(out_item) + 4 * g),
                       "r"(f2_to_h2(o0, o1)), "r"(f2_to_h2(o2, o3))
                       : "memory");
        } else {
          float *dst = prm.accum + ((bq - static_cast<long long>(b) * Q) * M + m) * 32 + 4 * g;
          asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(o0), "f"(o1), "f"(o2), "f"(o3) : "memory");
        }
      }
    } else {
      // ---- INT8: dp2a accumulate. 8 lanes per sample record: column = (lane>>2)&1, 8 channels per lane.
      const int slot = lane >> 3, col = (lane >> 2) & 1, cj = lane & 3;
      int acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const char *vl = vslab + cj * 16;
#pragma unroll 2
      for (int s = 0; s < steps; ++s) {
        const uint2 rr = reinterpret_cast<const uint2 *>(rec + s * 4 + slot)[col];  // {address, weight pair}
        const uint4 v = ldg128(vl + rr.x);
        acc[0] = dp2a_lo(rr.y, v.x, acc[0]), acc[1] = dp2a_hi(rr.y, v.x, acc[1]);
        acc[2] = dp2a_lo(rr.y, v.y, acc[2]), acc[3] = dp2a_hi(rr.y, v.y, acc[3]);
        acc[4] = dp2a_lo(rr.y, v.z, acc[4]), acc[5] = dp2a_hi(rr.y, v.z, acc[5]);
        acc[6] = dp2a_lo(rr.y, v.w, acc[6]), acc[7] = dp2a_hi(rr.y, v.w, acc[7]);
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i] += __shfl_xor_sync(kFullMask, acc[i], 4);
        acc[i] += __shfl_xor_sync(kFullMask, acc[i], 8);
        acc[i] += __shfl_xor_sync(kFullMask, acc[i], 16);
      }
      if (lane < 4) {
        // real = acc / 65536 * scale_value / sum ; q = T2int8(real / scale_out)
        const float mul = prm.scale_value / (65536.f * sum * prm.scale_out);
        uint32_t o[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          uint32_t word = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            word |= (static_cast<uint32_t>(to_int8_sat(static_cast<float>(acc[4 * i + j]) * mul)) & 0xffu) << (8 * j);
          o[i] = word;
        }
        asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(reinterpret_cast<int8_t *>(out_item) + 8 * cj),
                     "r"(o[0]), "r"(o[1])
                     : "memory");
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static std::atomic<long long> g_v2_budget{112ll << 20};  // bytes the packed stack may take (B200: 126 MB L2)

struct V2Call {
  const void *value, *ref, *off, *logits;
  const int32_t *shapes;  // device
  void *out, *workspace;
  size_t workspace_bytes;
  int B, S, M, C, L, Q, P, G;
  int el;  // 2 fp16, 1 int8
  int ref_is_half;
  float sv, so, sw, sout;
  const float *mask;
  float *accum;
  int4 *trace;
};

static bool v2_shape_ok(int C, int L, int P, int G) {
  const int NP = L * P;
  return C == 32 && L >= 1 && L <= kV2MaxLevels && NP >= 16 && NP <= 32 && G >= 1 && G <= P && (P % G) == 0;
}

// Upper bound of the packed stack that needs only the tensor DIMENSIONS (what getWorkspaceSize sees): both parity copies
// of every level, sum_l 2 * (H_l/2 + 1) * W_l <= S + 2 * sum_l W_l <= 3 * S entries per (camera, head).
static size_t v2_workspace_bound(int B, int S, int M, int el) {
  return static_cast<size_t>(3) * S * (64 * el) * B * M;
}

template <typename T, typename R, int EPI>
static int v2_launch_gather(const V2Params &p, cudaStream_t s) {
  int dev = 0, sms = 148, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msda_v2_kernel<T, R, EPI, false>, kV2Threads, 0) != cudaSuccess ||
      per_sm < 1)
    per_sm = 2;
  long long blocks = static_cast<long long>(sms) * per_sm;  // persistent: every resident warp loops over items
  const long long need = (p.items + kV2Warps - 1) / kV2Warps;
  if (blocks > need) blocks = need;
  if (p.trace != nullptr) {
    if constexpr (EPI == 0) {
      msda_v2_kernel<T, R, 0, true><<<static_cast<unsigned>(blocks), kV2Threads, 0, s>>>(p);
      return check_launch();
    }
  }
  msda_v2_kernel<T, R, EPI, false><<<static_cast<unsigned>(blocks), kV2Threads, 0, s>>>(p);
  return check_launch();
}

static int v2_run(const V2Call &c, cudaStream_t s) {
  if (!c.value || !c.shapes || !c.ref || !c.off || !c.logits || !c.workspace || !(c.out || (c.accum && c.mask)))
    return B200_ERR_BAD_PARAM;
  if (c.B <= 0 || c.S <= 0 || c.M <= 0 || c.Q <= 0) return B200_ERR_BAD_PARAM;
  if (!v2_shape_ok(c.C, c.L, c.P, c.G)) return B200_ERR_UNSUPPORTED;
  const int EB = 64 * c.el;
  if (c.workspace_bytes < v2_workspace_bound(c.B, c.S, c.M, c.el)) return B200_ERR_BAD_PARAM;
  if (static_cast<long long>(3) * c.S * EB >= (1ll << 31)) return B200_ERR_UNSUPPORTED;   // 32-bit entry offsets in a slab
  if (static_cast<long long>(c.B) * c.Q * c.M >= (1ll << 31)) return B200_ERR_UNSUPPORTED;
  const uintptr_t al = reinterpret_cast<uintptr_t>(c.value) | reinterpret_cast<uintptr_t>(c.workspace) |
                       reinterpret_cast<uintptr_t>(c.off) | reinterpret_cast<uintptr_t>(c.logits) |
                       reinterpret_cast<uintptr_t>(c.out) | reinterpret_cast<uintptr_t>(c.ref) |
                       reinterpret_cast<uintptr_t>(c.accum);
  if (al % 16 || reinterpret_cast<uintptr_t>(c.shapes) % 4) return B200_ERR_UNSUPPORTED;
  const long long budget = g_v2_budget.load(std::memory_order_relaxed);

  // 1. pack: persistent CTAs, 2 rows x 64 pixels x (M*32*el) bytes of shared memory each
  PackParams pp{};
  pp.value = static_cast<const char *>(c.value), pp.packed = static_cast<char *>(c.workspace), pp.shapes = c.shapes;
  pp.budget = budget, pp.B = c.B, pp.S = c.S, pp.M = c.M, pp.L = c.L;
  const int smem = 2 * kPackCols * c.M * 32 * c.el;
  if (smem > 200 * 1024) return B200_ERR_UNSUPPORTED;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int per_sm = (200 * 1024) / (smem + 1024) > 0 ? (200 * 1024) / (smem + 1024) : 1;
  unsigned gx = static_cast<unsigned>((sms * per_sm + c.B - 1) / c.B);
  const dim3 grid(gx < 1 ? 1 : gx, static_cast<unsigned>(c.B));
  cudaError_t ce;
  if (c.el == 2) {
    ce = cudaFuncSetAttribute(msda_pack_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    msda_pack_kernel<2><<<grid, 256, smem, s>>>(pp);
  } else {
    ce = cudaFuncSetAttribute(msda_pack_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    msda_pack_kernel<1><<<grid, 256, smem, s>>>(pp);
  }
  int st = check_launch();
  if (ce != cudaSuccess || st != B200_OK) return B200_ERR_LAUNCH;

  // 2. gather
  V2Params gp{};
  gp.packed = static_cast<const char *>(c.workspace), gp.ref = c.ref, gp.off = c.off, gp.logits = c.logits, gp.out = c.out;
  gp.shapes = c.shapes, gp.budget = budget, gp.L = c.L;
  gp.B = c.B, gp.M = c.M, gp.Q = c.Q, gp.P = c.P, gp.G = c.G, gp.NP = c.L * c.P;
  gp.items = static_cast<long long>(c.B) * c.Q * c.M;
  gp.ref_is_half = c.ref_is_half, gp.mask = c.mask, gp.accum = c.accum, gp.trace = c.trace;
  gp.scale_value = c.sv, gp.scale_offset = c.so, gp.scale_weight = c.sw, gp.scale_out = c.sout;
  if (c.el == 2) {
    if (c.accum) return v2_launch_gather<__half, __half, 1>(gp, s);
    return v2_launch_gather<__half, __half, 0>(gp, s);
  }
  if (c.ref_is_half) return v2_launch_gather<int8_t, __half, 0>(gp, s);
  return v2_launch_gather<int8_t, float, 0>(gp, s);
}

static int v2_trace_prologue(V2Call &c, int32_t *trace_records, cudaStream_t s) {
  if (!trace_records) return B200_OK;
  const long long n = static_cast<long long>(c.B) * c.Q * c.M * c.L * c.P;
  if (n <= 0 || cudaMemsetAsync(trace_records, 0, static_cast<size_t>(n) * 16, s) != cudaSuccess) return B200_ERR_LAUNCH;
  c.trace = reinterpret_cast<int4 *>(trace_records);
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" {

long long b200_msda_set_pack_budget(long long bytes) {
  return bytes > 0 ? g_v2_budget.exchange(bytes) : g_v2_budget.load();
}

size_t b200_msda_workspace_size(int dtype, int batch, int spatial_size, int num_heads, int channels, int num_levels,
                                int num_point, int points_per_group) {
  if (batch <= 0 || num_heads <= 0 || spatial_size <= 0) return 0;
  if (dtype != 1 && dtype != 2) return 0;
  if (!v2_shape_ok(channels, num_levels, num_point, points_per_group)) return 0;
  return v2_workspace_bound(batch, spatial_size, num_heads, dtype == 1 ? 2 : 1);
}

int b200_msda_f16_ws(const void *value, const int32_t *spatial_shapes, const void *reference_points,
                     const void *sampling_offsets, const void *attn_weight, int batch, int spatial_size, int num_heads,
                     int channels, int num_levels, int num_query, int num_point, int points_per_group, void *out,
                     void *workspace, size_t workspace_bytes, int32_t *trace_records, void *stream) {
  V2Call c{};
  c.value = value, c.shapes = spatial_shapes, c.ref = reference_points, c.off = sampling_offsets;
  c.logits = attn_weight, c.out = out, c.workspace = workspace, c.workspace_bytes = workspace_bytes;
  c.B = batch, c.S = spatial_size, c.M = num_heads, c.C = channels, c.L = num_levels, c.Q = num_query, c.P = num_point;
  c.G = points_per_group, c.el = 2, c.ref_is_half = 1, c.sv = c.so = c.sw = c.sout = 1.f;
  if (!v2_shape_ok(c.C, c.L, c.P, c.G)) return B200_ERR_UNSUPPORTED;
  const int st = v2_trace_prologue(c, trace_records, static_cast<cudaStream_t>(stream));
  return st != B200_OK ? st : v2_run(c, static_cast<cudaStream_t>(stream));
}

int b200_msda_i8_ws(const int8_t *value, float scale_value, const int32_t *spatial_shapes, const void *reference_points,
                    int ref_is_half, const int8_t *sampling_offsets, float scale_offset, const int8_t *attn_weight,
                    float scale_weight, int batch, int spatial_size, int num_heads, int channels, int num_levels,
                    int num_query, int num_point, int points_per_group, int8_t *out, float scale_out, void *workspace,
                    size_t workspace_bytes, int32_t *trace_records, void *stream) {
  if (!(scale_out > 0.f)) return B200_ERR_BAD_PARAM;
  V2Call c{};
  c.value = value, c.shapes = spatial_shapes, c.ref = reference_points, c.off = sampling_offsets;
  c.logits = attn_weight, c.out = out, c.workspace = workspace, c.workspace_bytes = workspace_bytes;
  c.B = batch, c.S = spatial_size, c.M = num_heads, c.C = channels, c.L = num_levels, c.Q = num_query, c.P = num_point;
  c.G = points_per_group, c.el = 1, c.ref_is_half = ref_is_half;
  c.sv = scale_value, c.so = scale_offset, c.sw = scale_weight, c.sout = scale_out;
  if (!v2_shape_ok(c.C, c.L, c.P, c.G)) return B200_ERR_UNSUPPORTED;
  const int st = v2_trace_prologue(c, trace_records, static_cast<cudaStream_t>(stream));
  return st != B200_OK ? st : v2_run(c, static_cast<cudaStream_t>(stream));
}

int b200_msda_sca_f16_ws(const void *value, const int32_t *spatial_shapes, const void *reference_points,
                         const void *sampling_offsets, const void *attn_weight, const float *bev_mask, int batch,
                         int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                         int points_per_group, float *accum, void *workspace, size_t workspace_bytes, void *stream) {
  V2Call c{};
  c.value = value, c.shapes = spatial_shapes, c.ref = reference_points, c.off = sampling_offsets;
  c.logits = attn_weight, c.out = nullptr, c.workspace = workspace, c.workspace_bytes = workspace_bytes;
  c.B = batch, c.S = spatial_size, c.M = num_heads, c.C = channels, c.L = num_levels, c.Q = num_query, c.P = num_point;
  c.G = points_per_group, c.el = 2, c.ref_is_half = 1, c.sv = c.so = c.sw = c.sout = 1.f;
  c.mask = bev_mask, c.accum = accum;
  return v2_run(c, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
