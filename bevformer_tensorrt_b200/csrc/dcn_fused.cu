// dcn_fused.cu — DCNv2 as ONE fused implicit GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), FP16.
//
// out[b, co, p] = bias[co] + sum_{t, c} W[co, c, t] * mask[b,t,p] * bilinear(x[b, c], p, t)     (reference:
// modulatedDeformableConv2dKernel.cu:259-318 im2col + :735-759 cuBLAS GEMM + bias kernel, with a 26.7 MB column buffer
// per image in between). Here the column tile never exists in global memory:
//
//   GEMM view   D[M = Co, N = 128 output pixels of one image] += A[M, K] * B[N, K]^T,  K = kh*kw*C ordered
//               (64-channel chunk, tap, channel)
//   A (weights) : pre-permuted once per call to [Co][C/64][tap][64] (K-major rows); producers copy the k-block slice into
//                 shared memory in the canonical K-major SWIZZLE_128B layout (128-byte rows, 16-byte chunks XOR row&7).
//   B (columns) : GATHERED by the producer warps straight into the same canonical layout: the input is first brought to
//                 NHWC so that the 64 channels of a k-block are 128 contiguous bytes per bilinear corner; a thread
//                 does 4 x LDG.128, blends 8 channels in fp32 with the (mask-folded) corner weights, packs to fp16 and
//                 issues one swizzled STS.128. Sampling positions / corner weights of the tile's 128 pixels x kh*kw taps
//                 are computed once per tile into a shared-memory table (fp32, op-for-op the reference's index math).
//   MMA         : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M128 x N128 x K16, fp32 accumulate in
//                 TMEM), Co/128 accumulators of 128 TMEM columns each; smem stages are released with tcgen05.commit on
//                 mbarriers; producers signal filled stages after fence.proxy.async.
//   epilogue    : the producer warps read their TMEM lane quadrant with tcgen05.ld.32x32b.x32 (thread = output channel,
//                 registers = 32 consecutive pixels), add the bias, convert to fp16 and store NCHW rows (64-byte runs).
// Persistent CTAs (one per SM) loop over (image, 256- or 128-pixel tile) work items.
//
// Requirements of this path (otherwise dcn.cu's v1 path runs): FP16, groups == 1, deformable_groups == 1,
// C % 64 == 0, Co in {128, 256, 512}, kh*kw <= 9.
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"

namespace b200 {

constexpr int kMaxBN = 256;         // output pixels per tile (UMMA N): 256 when the accumulators fit TMEM, else 128
constexpr int kBK = 64;             // channels per k-block: 64 fp16 = one 128-byte swizzle row
constexpr int kProducerWarps = 16;  // 512 gather threads: the kernel is latency-bound on the corner loads
constexpr int kFusedThreads = (kProducerWarps + 2) * 32;  // + 1 MMA/TMEM warp + 1 TMA warp (weights)
constexpr int kMaxTaps = 9;

struct DcnFusedParams {
  const __half *x_nhwc, *w_r;
  const void *bias, *offset, *mask;  // FP16 path: __half; INT8 path: offset/mask int8, bias float (converted by the host)
  void *out;                         // FP16 path: __half; INT8 path: int8
  float scale_off, scale_mask, out_mul, out_div;  // INT8 path: dequant scales; out = T2int8((acc*out_mul + bias)/out_div)
  float out_inv;                                  // 1 / out_div, rounded once on the host
  int B, C, H, W, Co, kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w, Ho, Wo;
  int tiles_per_img, num_tiles, kb_per_tap, num_kb;
};

struct __align__(16) TapEntry {  // 16 bytes: one LDS.128 / STS.128
  int pix;       // (h0*W + w0) of the top-left corner, clamped into the image
  int step;      // bit0: column neighbour usable, bit1: row neighbour usable
  uint32_t w12;  // fp16 pair: corner weights (top-left, top-right), mask folded in
  uint32_t w34;  // fp16 pair: (bottom-left, bottom-right)
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(=1) | SBO(1024 B) |
// version 1 | layout SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, both K-major, N>>3, M>>4.
template <int BN>
constexpr uint32_t kIdesc = (1u << 4) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- pre-passes -----------------------------------------------------------------------------------------------------
// NCHW -> NHWC (fp16): 64 channels x 64 pixels per block through shared memory; 128-bit global accesses on both sides
// when HW % 8 == 0 (rows of the NCHW planes are then 16-byte aligned), scalar loads otherwise.
__global__ void __launch_bounds__(512) dcn_nchw_to_nhwc_kernel(const __half *__restrict__ in, __half *__restrict__ out,
                                                               int C, int HW) {
  __shared__ __half tile[64][66];  // row stride 33 words: column reads are conflict-free
  const int b = blockIdx.z, p0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int t = threadIdx.x;
  {
    const int c = t >> 3, pj = (t & 7) * 8;  // 64 channel rows x 8 chunks of 8 pixels
    const __half *src = in + (static_cast<long long>(b) * C + c0 + c) * HW + p0 + pj;
    if ((HW & 7) == 0 && p0 + pj + 8 <= HW) {
      const uint4 v = ldg128(src);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint32_t *>(&tile[c][pj + 2 * i]) = w[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) tile[c][pj + i] = (p0 + pj + i < HW) ? src[i] : __float2half(0.f);
    }
  }
  __syncthreads();
  {
    const int pp = t >> 3, cj = (t & 7) * 8;  // 64 pixels x 8 chunks of 8 channels
    if (p0 + pp < HW) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __half2 h = __halves2half2(tile[cj + 2 * i][pp], tile[cj + 2 * i + 1][pp]);
        w[i] = *reinterpret_cast<const uint32_t *>(&h);
      }
      *reinterpret_cast<uint4 *>(out + (static_cast<long long>(b) * HW + p0 + pp) * C + c0 + cj) =
          make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// W[co][c][t] -> Wr[co][k],  k = (c / 64) * (kk * 64) + t * 64 + (c % 64): K is ordered (channel chunk, tap, channel).
// Consecutive k-blocks are then the kh*kw taps of ONE 64-channel chunk, whose bilinear corners overlap spatially, so the
// gather of tap t+1 mostly hits in L1 what tap t brought in.
__global__ void dcn_weight_reorder_kernel(const __half *__restrict__ w, __half *__restrict__ wr, int Co, int C, int kk) {
  const long long n = static_cast<long long>(Co) * C * kk;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % kBK);
    const int t = static_cast<int>((i / kBK) % kk);
    const int cc = static_cast<int>((i / (static_cast<long long>(kBK) * kk)) % (C / kBK));
    const long long co = i / (static_cast<long long>(C) * kk);
    wr[i] = w[(co * C + cc * kBK + ci) * kk + t];
  }
}

// T2int8(real / div) of the INT8 epilogue (…Conv2dKernel.cu:578-579, T2int8 :51-55) without the generic division
// sequence (reciprocal, fix-ups, slow-path check: ~10 issue slots on each of 65 k outputs per tile, on the warps that
// also feed the tensor core): quotient from the host-rounded reciprocal plus one Newton step on the exact remainder
// (two FMAs) — the correctly rounded quotient except for isolated last-bit cases, which move a result only when it sits
// within 2^-24 of a rounding boundary. Saturation and round-half-away-from-zero in one conversion: trunc(q + copysign(0.5, q))
// with cvt.sat clamps to [-128, 127] exactly like the clamp-then-round of to_int8_sat (NaN -> 0 in both).
__device__ __forceinline__ int requant_i8(float real, float div, float inv) {
  float q = real * inv;
  q = fmaf(fmaf(-q, div, real), inv, q);
  const float h = __uint_as_float((__float_as_uint(q) & 0x80000000u) | 0x3f000000u);
  int r;
  asm("cvt.rzi.sat.s8.f32 %0, %1;" : "=r"(r) : "f"(q + h));
  return r;
}

// ---- the fused kernel -----------------------------------------------------------------------------------------------
template <int MH>
struct FusedCfg {
  // A 128x128x16 MMA reads 8 KB of operands per 64 tensor cycles — the whole shared-memory port — so the tile is made
  // 256 pixels wide whenever MH x 256 fp32 accumulator columns fit the 512 TMEM columns (Co <= 256): half the A re-reads
  // per MAC. Co = 512 keeps 128-pixel tiles.
  static constexpr int kBN = MH <= 2 ? 256 : 128;
  static constexpr int kABytes = MH * 128 * 128;  // Co rows x 128 B
  static constexpr int kBBytes = kBN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTableBytes = kMaxTaps * kBN * static_cast<int>(sizeof(TapEntry));
  static constexpr int kStages = (220 * 1024 - kTableBytes) / kStageBytes >= 4 ? 4 : (220 * 1024 - kTableBytes) / kStageBytes;
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kTableBytes + 256 /*barriers*/;
  static constexpr int kTmemCols = MH * kBN <= 128 ? 128 : (MH * kBN <= 256 ? 256 : 512);
  static_assert(kStages >= 2 && MH * kBN <= 512, "tile does not fit shared memory / TMEM");
};

// I8 = true: the INT8 plugin flavour computed by in-register dequantisation — the pre-passes hand this kernel the int8
// activations / weights as (exact) fp16 integers, offsets and masks are dequantised while the sampling table is built,
// the fp32 accumulators are rescaled by scale_in*scale_w, biased and requantised once (T2int8) in the epilogue.
template <int MH, bool I8>
__global__ void __launch_bounds__(kFusedThreads, 1) dcn_fused_kernel(const DcnFusedParams p,
                                                                     const __grid_constant__ CUtensorMap tmap_w) {
  using Cfg = FusedCfg<MH>;
  constexpr int S = Cfg::kStages;
  constexpr int kBN = Cfg::kBN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *stage_base = smem;
  TapEntry *table = reinterpret_cast<TapEntry *>(smem + S * Cfg::kStageBytes);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S * Cfg::kStageBytes + Cfg::kTableBytes);
  // bars[0..S) full, [S..2S) empty, [2S] tmem_full, [2S+1] tmem_empty, then the TMEM base pointer
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + 2 * S + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
  const uint32_t tmem_full = smem_u32(bars + 2 * S), tmem_empty = smem_u32(bars + 2 * S + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full0 + 8 * s, kProducerWarps + 1);  // 16 gather warps + the TMA warp's expect_tx arrive
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, kProducerWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kProducerWarps) {  // the MMA warp owns TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int HoWo = p.Ho * p.Wo, HW = p.H * p.W, kk = p.kh * p.kw, K = kk * p.C;
  const int my_tiles = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / gridDim.x;

  if (warp == kProducerWarps + 1) {
    // =================================== TMA: weight slices (A) ===================================
    // cp.async.bulk.tensor writes the [128 rows x 64 k] box in the same 128-byte-swizzled layout the UMMA descriptor
    // expects; completion is counted in bytes on the stage's full barrier.
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
      uint32_t kbt = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int kb = 0; kb < p.num_kb; ++kb, ++kbt) {
          const int s = kbt % S;
          mbar_wait(empty0 + 8 * s, ((kbt / S) & 1) ^ 1);
          const uint32_t bar = full0 + 8 * s;
          const uint32_t dst = smem_u32(stage_base + s * Cfg::kStageBytes);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(Cfg::kABytes) : "memory");
#pragma unroll
          for (int mh = 0; mh < MH; ++mh)
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                ::"r"(dst + mh * 16384), "l"(&tmap_w), "r"(kb * kBK), "r"(mh * 128), "r"(bar)
                : "memory");
        }
      }
    }
  } else if (warp == kProducerWarps) {
    // =================================== MMA issuer ===================================
    uint32_t kbt = 0;
    for (int it = 0; it < my_tiles; ++it) {
      if (it > 0) mbar_wait(tmem_empty, (it - 1) & 1);  // epilogue of the previous tile has drained TMEM
      tc_fence_after();
      for (int kb = 0; kb < p.num_kb; ++kb, ++kbt) {
        const int s = kbt % S;
        mbar_wait(full0 + 8 * s, (kbt / S) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(stage_base + s * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int mh = 0; mh < MH; ++mh) {
#pragma unroll
            for (int k4 = 0; k4 < kBK / 16; ++k4) {
              umma_f16(tmem_base + mh * kBN, make_sw128_desc(a_addr + mh * 16384 + k4 * 32),
                       make_sw128_desc(b_addr + k4 * 32), kIdesc<kBN>, (kb | k4) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty0 + 8 * s);                          // frees the stage when these MMAs have read it
          if (kb == p.num_kb - 1) umma_commit(tmem_full);       // accumulators complete
        }
        __syncwarp();
      }
    }
  } else {
    // =================================== producers + epilogue ===================================
    const int pt = threadIdx.x;  // 0..kProducerWarps*32-1
    uint32_t kbt = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int b = tile / p.tiles_per_img;
      const int p0 = (tile - b * p.tiles_per_img) * kBN;

      // ---- sampling table of this tile: (tap, pixel) -> clamped corner + mask-folded bilinear weights
      for (int e = pt; e < kk * kBN; e += kProducerWarps * 32) {
        const int n = e % kBN, t = e / kBN;
        const int pix = p0 + n;
        TapEntry te{0, 0, 0u, 0u};
        if (pix < HoWo) {
          const int h_col = pix / p.Wo, w_col = pix - h_col * p.Wo;
          const int i = t / p.kw, j = t - i * p.kw;
          const long long oi = (static_cast<long long>(b) * 2 * kk + 2 * t) * HoWo + pix;
          const long long mi = (static_cast<long long>(b) * kk + t) * HoWo + pix;
          float oh, ow, m;
          if (I8) {  // off*scale rounded to fp32 first, like the reference INT8 im2col (…Conv2dKernel.cu:509-514)
            oh = static_cast<float>(__ldg(static_cast<const int8_t *>(p.offset) + oi)) * p.scale_off;
            ow = static_cast<float>(__ldg(static_cast<const int8_t *>(p.offset) + oi + HoWo)) * p.scale_off;
            m = static_cast<float>(__ldg(static_cast<const int8_t *>(p.mask) + mi)) * p.scale_mask;
          } else {
            oh = __half2float(__ldg(static_cast<const __half *>(p.offset) + oi));
            ow = __half2float(__ldg(static_cast<const __half *>(p.offset) + oi + HoWo));
            m = __half2float(__ldg(static_cast<const __half *>(p.mask) + mi));
          }
          const float h_im = __fadd_rn(static_cast<float>(h_col * p.stride_h - p.pad_h + i * p.dil_h), oh);
          const float w_im = __fadd_rn(static_cast<float>(w_col * p.stride_w - p.pad_w + j * p.dil_w), ow);
          if (h_im > -1.f && w_im > -1.f && h_im < static_cast<float>(p.H) && w_im < static_cast<float>(p.W)) {
            const float hf = floorf(h_im), wf = floorf(w_im);
            const int h_low = static_cast<int>(hf), w_low = static_cast<int>(wf);
            const float lh = __fsub_rn(h_im, hf), lw = __fsub_rn(w_im, wf), hh = 1.f - lh, hw = 1.f - lw;
            const bool tp = h_low >= 0, bt = h_low + 1 <= p.H - 1, lf = w_low >= 0, rt = w_low + 1 <= p.W - 1;
            te.pix = max(h_low, 0) * p.W + max(w_low, 0);
            te.step = ((lf && rt) ? 1 : 0) | ((tp && bt) ? 2 : 0);
            te.w12 = f2_to_h2((tp && lf) ? hh * hw * m : 0.f, (tp && rt) ? hh * lw * m : 0.f);
            te.w34 = f2_to_h2((bt && lf) ? lh * hw * m : 0.f, (bt && rt) ? lh * lw * m : 0.f);
          }
        }
        table[e] = te;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kProducerWarps * 32) : "memory");

      // ---- k-blocks: fill stage s with the weight slice (A) and the gathered column tile (B)
      constexpr int kRowsPerPass = kProducerWarps * 32 / 8;
      const int j8 = pt & 7, rowp = pt >> 3;  // 16-byte chunk within the 128-byte row; row within a pass
      const __half *xb = p.x_nhwc + static_cast<long long>(b) * HW * p.C;
      constexpr int kPasses = kBN / kRowsPerPass;
      for (int kb = 0; kb < p.num_kb; ++kb, ++kbt) {
        const int s = kbt % S;
        uint8_t *b_st = stage_base + s * Cfg::kStageBytes + Cfg::kABytes;
        const int cc = kb / kk, t = kb - cc * kk;  // K order: (channel chunk, tap, channel)
        // Issue every corner load of this k-block first and only then wait for the stage to be free: the loads fly
        // while the tensor core still reads the stage's previous contents. (The weight slice A arrives by TMA.)
        TapEntry te[kPasses];
        uint4 v[kPasses][4];
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass) te[pass] = table[t * kBN + pass * kRowsPerPass + rowp];
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass) {
          const __half *c00 = xb + static_cast<long long>(te[pass].pix) * p.C + cc * kBK + j8 * 8;
          const int dx = (te[pass].step & 1) ? p.C : 0, dy = (te[pass].step & 2) ? p.W * p.C : 0;
          v[pass][0] = ldg128(c00), v[pass][1] = ldg128(c00 + dx), v[pass][2] = ldg128(c00 + dy),
          v[pass][3] = ldg128(c00 + dy + dx);
        }
        mbar_wait(empty0 + 8 * s, ((kbt / S) & 1) ^ 1);
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass) {
          const int n = pass * kRowsPerPass + rowp;
          // The column operand is FP16 in any case (tensor-core input); blending the four corners with packed HFMA2
          // (fp16 weights = mask x bilinear, <= 1) instead of fp32 costs two extra 2^-11 roundings on values that are
          // then summed over K = kh*kw*C with fp32 accumulation — far below the FP16 output rounding — and removes all
          // conversions from the producers' inner loop (the reference's FP16 im2col is fp16 arithmetic too, :321-388).
          const __half2 w12 = *reinterpret_cast<const __half2 *>(&te[pass].w12);
          const __half2 w34 = *reinterpret_cast<const __half2 *>(&te[pass].w34);
          const __half2 w1 = __low2half2(w12), w2 = __high2half2(w12), w3 = __low2half2(w34), w4 = __high2half2(w34);
          const uint32_t a[4] = {v[pass][0].x, v[pass][0].y, v[pass][0].z, v[pass][0].w};
          const uint32_t bq[4] = {v[pass][1].x, v[pass][1].y, v[pass][1].z, v[pass][1].w};
          const uint32_t c[4] = {v[pass][2].x, v[pass][2].y, v[pass][2].z, v[pass][2].w};
          const uint32_t d[4] = {v[pass][3].x, v[pass][3].y, v[pass][3].z, v[pass][3].w};
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            __half2 r = __hmul2(w1, *reinterpret_cast<const __half2 *>(&a[q]));
            r = __hfma2(w2, *reinterpret_cast<const __half2 *>(&bq[q]), r);
            r = __hfma2(w3, *reinterpret_cast<const __half2 *>(&c[q]), r);
            r = __hfma2(w4, *reinterpret_cast<const __half2 *>(&d[q]), r);
            o[q] = *reinterpret_cast<const uint32_t *>(&r);
          }
          *reinterpret_cast<uint4 *>(b_st + n * 128 + ((j8 ^ (n & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async();  // make the generic-proxy stores visible to the tensor-core (async) proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + 8 * s);
      }

      // ---- epilogue: TMEM -> registers -> +bias -> fp16 -> NCHW
      mbar_wait(tmem_full, it & 1);
      tc_fence_after();
      // a warp may only touch its own TMEM lane quadrant (warp % 4); the warps of a quadrant split the
      // (accumulator, 32-column chunk) units between them
      const int quad = warp & 3;
      for (int u = warp >> 2; u < MH * (kBN / 32); u += kProducerWarps / 4) {
        const int mh = u / (kBN / 32), c0 = (u % (kBN / 32)) * 32;
        const int co = mh * 128 + quad * 32 + lane;
        if (I8) {
          const float bias = __ldg(static_cast<const float *>(p.bias) + co);
          int8_t *orow = static_cast<int8_t *>(p.out) + (static_cast<long long>(b) * p.Co + co) * HoWo + p0;
          uint32_t r[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + mh * kBN + c0, r);
          // out = T2int8((acc * scale_i*scale_w + bias) / scale_o)   (…Conv2dKernel.cu:578-579)
          // The 32 outputs of this lane are 32 consecutive bytes of ONE output row; word j holds pixels 4j .. 4j+3.
          uint32_t ow[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint32_t word = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float real = fmaf(__uint_as_float(r[4 * j + i]), p.out_mul, bias);
              word |= (static_cast<uint32_t>(requant_i8(real, p.out_div, p.out_inv)) & 0xffu) << (8 * i);
            }
            ow[j] = word;
          }
          // An INT8 row is Ho*Wo BYTES long, so rows start 16-byte aligned only when Ho*Wo % 16 == 0 (the R101 stage-3
          // map is 58 x 100 = 5800 = 8 mod 16: every other row starts on an 8-byte boundary). Widest store the row's
          // alignment allows; byte stores (one 32-byte sector touched per byte) only for ragged tails and odd widths.
          const uintptr_t oa = reinterpret_cast<uintptr_t>(orow + c0);
          if (p0 + c0 + 32 <= HoWo && (oa & 3) == 0) {
            if ((oa & 15) == 0) {
              *reinterpret_cast<uint4 *>(orow + c0) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
              *reinterpret_cast<uint4 *>(orow + c0 + 16) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
            } else if ((oa & 7) == 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j) *reinterpret_cast<uint2 *>(orow + c0 + 8 * j) = make_uint2(ow[2 * j], ow[2 * j + 1]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) *reinterpret_cast<uint32_t *>(orow + c0 + 4 * j) = ow[j];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (p0 + c0 + i < HoWo) orow[c0 + i] = static_cast<int8_t>((ow[i >> 2] >> (8 * (i & 3))) & 0xffu);
          }
        } else {
        const float bias = __half2float(__ldg(static_cast<const __half *>(p.bias) + co));
        __half *orow = static_cast<__half *>(p.out) + (static_cast<long long>(b) * p.Co + co) * HoWo + p0;
        {
          uint32_t r[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + mh * kBN + c0, r);
          // the lane's 32 outputs = 64 consecutive bytes of one output row; word j holds pixels 2j, 2j+1
          uint32_t hw[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            hw[j] = f2_to_h2(__uint_as_float(r[2 * j]) + bias, __uint_as_float(r[2 * j + 1]) + bias);
          // Rows are Ho*Wo halves long: 16-byte aligned when Ho*Wo % 8 == 0 (58 x 100), only 4-byte aligned for the R101
          // stage-4 map (29 x 50 = 1450). Widest store the row's alignment allows; 2-byte stores only for ragged tails
          // and odd Ho*Wo.
          const uintptr_t oa = reinterpret_cast<uintptr_t>(orow + c0);
          if (p0 + c0 + 32 <= HoWo && (oa & 3) == 0) {
            if ((oa & 15) == 0) {
#pragma unroll
              for (int v = 0; v < 4; ++v)
                *reinterpret_cast<uint4 *>(orow + c0 + 8 * v) = make_uint4(hw[4 * v], hw[4 * v + 1], hw[4 * v + 2], hw[4 * v + 3]);
            } else if ((oa & 7) == 0) {
#pragma unroll
              for (int v = 0; v < 8; ++v) *reinterpret_cast<uint2 *>(orow + c0 + 4 * v) = make_uint2(hw[2 * v], hw[2 * v + 1]);
            } else {
#pragma unroll
              for (int v = 0; v < 16; ++v) *reinterpret_cast<uint32_t *>(orow + c0 + 2 * v) = hw[v];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (p0 + c0 + i < HoWo)
                reinterpret_cast<unsigned short *>(orow)[c0 + i] = static_cast<unsigned short>(hw[i >> 1] >> (16 * (i & 1)));
          }
        }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kProducerWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
  }
}

// ---- host -----------------------------------------------------------------------------------------------------------
size_t dcn_fused_workspace_bytes(int batch, int channels, int height, int width, int channels_out, int kk) {
  const size_t x = (static_cast<size_t>(batch) * height * width * channels * 2 + 255) / 256 * 256;
  const size_t w = (static_cast<size_t>(channels_out) * channels * kk * 2 + 255) / 256 * 256;
  return x + w + 2048;  // + a zero / fp32 bias vector (<= 512 channels x 4 B)
}

bool dcn_fused_supported(int channels, int channels_out, int kk, int group, int deformable_group) {
  return group == 1 && deformable_group == 1 && channels % kBK == 0 &&
         (channels_out == 128 || channels_out == 256 || channels_out == 512) && kk <= kMaxTaps;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D tensor map over the permuted weights Wr[Co][K] (fp16): box = 64 k x 128 rows, 128-byte swizzle.
static int make_weight_tmap(CUtensorMap *tm, const __half *w_r, int Co, int K) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
      return B200_ERR_LAUNCH;
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(Co)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(K) * 2};
  const cuuint32_t box[2] = {kBK, 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half *>(w_r), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B200_OK : B200_ERR_LAUNCH;
}

template <int MH, bool I8>
static int launch_fused(const DcnFusedParams &p, cudaStream_t stream) {
  using Cfg = FusedCfg<MH>;
  // per launch: the attribute is per device (and this may run on several devices / host threads); the call is cheap
  if (cudaFuncSetAttribute(dcn_fused_kernel<MH, I8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) !=
      cudaSuccess)
    return B200_ERR_LAUNCH;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  CUtensorMap tmap;
  const int ts = make_weight_tmap(&tmap, p.w_r, p.Co, p.kh * p.kw * p.C);
  if (ts != B200_OK) return ts;
  dcn_fused_kernel<MH, I8><<<grid, kFusedThreads, Cfg::kSmemBytes, stream>>>(p, tmap);
  return check_launch();
}

int dcn_fused_f16(const __half *input, const __half *weight, const __half *bias, const __half *offset,
                  const __half *mask, __half *output, void *workspace, int batch, int channels, int height, int width,
                  int channels_out, int kernel_w, int kernel_h, int stride_w, int stride_h, int pad_w, int pad_h,
                  int dilation_w, int dilation_h, int Ho, int Wo, int flags, cudaStream_t stream) {
  // flags: bit 0 = `input` is already NHWC (channels-last), bit 1 = `weight` is already permuted (b200_dcn_pack_weights_f16)
  const int kk = kernel_h * kernel_w, HW = height * width;
  __half *x_nhwc = static_cast<__half *>(workspace);
  __half *w_r = reinterpret_cast<__half *>(static_cast<uint8_t *>(workspace) +
                                           (static_cast<size_t>(batch) * HW * channels * 2 + 255) / 256 * 256);
  if (!bias) {  // the kernel always reads a bias vector (no per-thread null test in the epilogue): use zeros
    __half *zb = reinterpret_cast<__half *>(reinterpret_cast<uint8_t *>(w_r) +
                                            (static_cast<size_t>(channels_out) * channels * kk * 2 + 255) / 256 * 256);
    if (cudaMemsetAsync(zb, 0, static_cast<size_t>(channels_out) * 2, stream) != cudaSuccess) return B200_ERR_LAUNCH;
    bias = zb;
  }
  int st = B200_OK;
  if (flags & 1) {
    x_nhwc = const_cast<__half *>(input);
  } else {
    dcn_nchw_to_nhwc_kernel<<<dim3((HW + 63) / 64, channels / 64, batch), 512, 0, stream>>>(input, x_nhwc, channels, HW);
    st = check_launch();
    if (st != B200_OK) return st;
  }
  if (flags & 2) {
    w_r = const_cast<__half *>(weight);
  } else {
    const long long wn = static_cast<long long>(channels_out) * channels * kk;
    dcn_weight_reorder_kernel<<<static_cast<unsigned>((wn + 255) / 256), 256, 0, stream>>>(weight, w_r, channels_out,
                                                                                           channels, kk);
    st = check_launch();
    if (st != B200_OK) return st;
  }

  DcnFusedParams p{};
  p.x_nhwc = x_nhwc, p.w_r = w_r, p.bias = bias, p.offset = offset, p.mask = mask, p.out = output;
  p.B = batch, p.C = channels, p.H = height, p.W = width, p.Co = channels_out, p.kh = kernel_h, p.kw = kernel_w;
  p.pad_h = pad_h, p.pad_w = pad_w, p.stride_h = stride_h, p.stride_w = stride_w, p.dil_h = dilation_h,
  p.dil_w = dilation_w, p.Ho = Ho, p.Wo = Wo;
  p.kb_per_tap = channels / kBK;
  p.num_kb = kk * p.kb_per_tap;
  const int mh = channels_out / 128;
  const int bn = mh == 1 ? FusedCfg<1>::kBN : (mh == 2 ? FusedCfg<2>::kBN : FusedCfg<4>::kBN);
  p.tiles_per_img = (Ho * Wo + bn - 1) / bn;
  p.num_tiles = p.tiles_per_img * batch;
  switch (mh) {
    case 1: return launch_fused<1, false>(p, stream);
    case 2: return launch_fused<2, false>(p, stream);
    case 4: return launch_fused<4, false>(p, stream);
    default: return B200_ERR_UNSUPPORTED;
  }
}

int dcn_pack_weights_f16(const __half *weight, __half *packed, int channels_out, int channels, int kk, cudaStream_t stream) {
  const long long wn = static_cast<long long>(channels_out) * channels * kk;
  dcn_weight_reorder_kernel<<<static_cast<unsigned>((wn + 255) / 256), 256, 0, stream>>>(weight, packed, channels_out,
                                                                                         channels, kk);
  return check_launch();
}

// ---- INT8 flavour: pre-passes + launch ---------------------------------------------------------------------------------
// input kCHW4 int8 [B, C/4, H, W, 4] -> NHWC fp16 (the int8 values as exact fp16 integers): 64 pixels x 64 channels per
// block through shared memory, so that both sides are coalesced (reads: 64 consecutive pixels of a channel quad = 256
// contiguous bytes; writes: the 64 channels of a pixel = one 128-byte row)
__global__ void __launch_bounds__(256) dcn_chw4_to_nhwc_f16_kernel(const int8_t *__restrict__ in, __half *__restrict__ out,
                                                                   int C, int HW) {
  __shared__ uint32_t tile[16][65];  // [channel quad][pixel], 65: the transposed reads below are conflict-free
  const int b = blockIdx.z, p0 = blockIdx.x * 64, q0 = blockIdx.y * 16;
  const int t = threadIdx.x;
  {
    const int quad = t >> 4, pj = (t & 15) * 4;  // 16 quads x 16 groups of 4 pixels
    const uint32_t *src = reinterpret_cast<const uint32_t *>(in) + (static_cast<long long>(b) * (C / 4) + q0 + quad) * HW + p0 + pj;
#pragma unroll
    for (int i = 0; i < 4; ++i) tile[quad][pj + i] = (p0 + pj + i < HW) ? __ldg(src + i) : 0u;
  }
  __syncthreads();
  {
    const int pp = t >> 2, cj = t & 3;  // 64 pixels x 4 chunks of 16 channels
    if (p0 + pp < HW) {
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t u = tile[cj * 4 + i][pp];
        o[2 * i] = f2_to_h2(static_cast<float>(static_cast<int8_t>(u)), static_cast<float>(static_cast<int8_t>(u >> 8)));
        o[2 * i + 1] = f2_to_h2(static_cast<float>(static_cast<int8_t>(u >> 16)), static_cast<float>(static_cast<int8_t>(u >> 24)));
      }
      uint4 *dst = reinterpret_cast<uint4 *>(out + (static_cast<long long>(b) * HW + p0 + pp) * C + q0 * 4 + cj * 16);
      dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
      dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
  }
}

// weight kCHW4 int8 [Co, C/4, kh, kw, 4] -> Wr fp16 [Co][C/64][t][64]
__global__ void dcn_weight_reorder_i8_kernel(const int8_t *__restrict__ w, __half *__restrict__ wr, int Co, int C, int kk) {
  const long long n = static_cast<long long>(Co) * C * kk;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % kBK);
    const int t = static_cast<int>((i / kBK) % kk);
    const int cc = static_cast<int>((i / (static_cast<long long>(kBK) * kk)) % (C / kBK));
    const long long co = i / (static_cast<long long>(C) * kk);
    const int c = cc * kBK + ci;
    wr[i] = __float2half_rn(static_cast<float>(w[((co * (C / 4) + c / 4) * kk + t) * 4 + (c & 3)]));
  }
}

// bias (float or half, may be null) -> float vector
__global__ void dcn_bias_to_f32_kernel(const void *bias, int is_half, float *out, int Co) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Co) out[i] = !bias ? 0.f : (is_half ? __half2float(static_cast<const __half *>(bias)[i]) : static_cast<const float *>(bias)[i]);
}

int dcn_fused_i8(const int8_t *input, float scale_i, const int8_t *weight, float scale_w, const void *bias, int bias_is_half,
                 const int8_t *offset, float scale_off, const int8_t *mask, float scale_mask, int8_t *output, float scale_o,
                 void *workspace, int batch, int channels, int height, int width, int channels_out, int kernel_w,
                 int kernel_h, int stride_w, int stride_h, int pad_w, int pad_h, int dilation_w, int dilation_h, int Ho,
                 int Wo, cudaStream_t stream) {
  const int kk = kernel_h * kernel_w, HW = height * width;
  uint8_t *ws = static_cast<uint8_t *>(workspace);
  __half *x_nhwc = reinterpret_cast<__half *>(ws);
  ws += (static_cast<size_t>(batch) * HW * channels * 2 + 255) / 256 * 256;
  __half *w_r = reinterpret_cast<__half *>(ws);
  ws += (static_cast<size_t>(channels_out) * channels * kk * 2 + 255) / 256 * 256;
  float *bias_f = reinterpret_cast<float *>(ws);  // 512 floats fit the 2 KB tail reserved by dcn_fused_workspace_bytes

  dcn_chw4_to_nhwc_f16_kernel<<<dim3((HW + 63) / 64, channels / 64, batch), 256, 0, stream>>>(input, x_nhwc, channels, HW);
  int st = check_launch();
  if (st != B200_OK) return st;
  const long long wn = static_cast<long long>(channels_out) * channels * kk;
  dcn_weight_reorder_i8_kernel<<<static_cast<unsigned>((wn + 255) / 256), 256, 0, stream>>>(weight, w_r, channels_out,
                                                                                            channels, kk);
  st = check_launch();
  if (st != B200_OK) return st;
  dcn_bias_to_f32_kernel<<<(channels_out + 255) / 256, 256, 0, stream>>>(bias, bias_is_half, bias_f, channels_out);
  st = check_launch();
  if (st != B200_OK) return st;

  DcnFusedParams p{};
  p.x_nhwc = x_nhwc, p.w_r = w_r, p.bias = bias_f, p.offset = offset, p.mask = mask, p.out = output;
  p.scale_off = scale_off, p.scale_mask = scale_mask, p.out_mul = scale_i * scale_w, p.out_div = scale_o, p.out_inv = 1.f / scale_o;
  p.B = batch, p.C = channels, p.H = height, p.W = width, p.Co = channels_out, p.kh = kernel_h, p.kw = kernel_w;
  p.pad_h = pad_h, p.pad_w = pad_w, p.stride_h = stride_h, p.stride_w = stride_w, p.dil_h = dilation_h,
  p.dil_w = dilation_w, p.Ho = Ho, p.Wo = Wo;
  p.kb_per_tap = channels / kBK;
  p.num_kb = kk * p.kb_per_tap;
  const int mh = channels_out / 128;
  const int bn = mh == 1 ? FusedCfg<1>::kBN : (mh == 2 ? FusedCfg<2>::kBN : FusedCfg<4>::kBN);
  p.tiles_per_img = (Ho * Wo + bn - 1) / bn;
  p.num_tiles = p.tiles_per_img * batch;
  switch (mh) {
    case 1: return launch_fused<1, true>(p, stream);
    case 2: return launch_fused<2, true>(p, stream);
    case 4: return launch_fused<4, true>(p, stream);
    default: return B200_ERR_UNSUPPORTED;
  }
}

}  // namespace b200
