// Microbenchmark 2 (round 2): what bounds a bilinear-tap gather on one B200 SM, per access shape?
//
// Every mode reads pseudo-random "pieces" out of a buffer that is either L2-resident (64 MB) or L1-resident (64 KB),
// `LPP` lanes per piece, `VB` bytes per lane. Reported: GB/s, bytes/clk/SM, pieces/clk/SM and distinct 128-byte
// lines/clk/SM (the L1 tag-stage currency), so the request-rate limit can be told apart from the byte-rate limit.
//
//   A  64 B piece   4 lanes x LDG.128   FP16 tap, value layout [S, heads, 32 ch]            (round-1 kernel)
//   B  32 B piece   4 lanes x LDG.64    INT8 tap, same layout                               (round-1 INT8 kernel)
//   C  32 B piece   2 lanes x LDG.128   INT8 tap, 16 channels per lane
//   D  128 B piece  8 lanes x LDG.128   64 B-aligned: INT8 2x2 footprint in a column-pair layout (may straddle 2 lines)
//   E  128 B piece  8 lanes x LDG.128   128 B-aligned: FP16 column pair (one full line per request)
//   F  64 B piece   8 lanes x LDG.64    FP16 tap spread over 8 lanes (tensor-core fragment order)
//   G  cp.async.bulk (TMA, 1-D) of 64 / 128 / 256 B pieces into shared memory, mbarrier completion
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int LPP, int VB, int ALIGN, int PIECE>
__global__ void __launch_bounds__(256) gather(const char* __restrict__ buf, uint32_t slots, int iters, uint4* out) {
  // slots = number of ALIGN-sized slots a piece may start at (power of two)
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPP;
  uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) / LPP;
  seed = seed * 2654435761u + 12345u;
  uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      seed = seed * 1664525u + 1013904223u;
      const uint32_t slot = (seed >> 8) & (slots - 1);
      const char* p = buf + (size_t)slot * ALIGN + sub * VB;
      if (VB == 16) v[u] = __ldg(reinterpret_cast<const uint4*>(p));
      else { const uint2 t = __ldg(reinterpret_cast<const uint2*>(p)); v[u] = make_uint4(t.x, t.y, 0, 0); }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc.x ^= v[u].x; acc.y += v[u].y; acc.z ^= v[u].z; acc.w += v[u].w; }
  }
  if (acc.x == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// TMA 1-D bulk copies: every lane issues U copies of SZ bytes per round into its own smem slots; one mbarrier per warp.
template <int SZ, int U>
__global__ void __launch_bounds__(256) bulk(const char* __restrict__ buf, uint32_t slots, int iters, uint4* out) {
  extern __shared__ __align__(128) char smem[];
  __shared__ __align__(8) unsigned long long bars[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  char* mine = smem + (size_t)(warp * 32 + lane) * U * SZ;
  const uint32_t bar = smem_u32(&bars[warp]);
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  uint32_t acc = 0, parity = 0;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32 * U * SZ) : "memory");
    __syncwarp();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      seed = seed * 1664525u + 1013904223u;
      const uint32_t slot = (seed >> 8) & (slots - 1);
      const char* src = buf + (size_t)slot * SZ;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(mine + u * SZ)),
                   "l"(src), "r"(SZ), "r"(bar)
                   : "memory");
    }
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    parity ^= 1;
    acc += *reinterpret_cast<const uint32_t*>(mine);
  }
  if (acc == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = make_uint4(acc, 0, 0, 0);
}

static int g_sms, g_clk;

template <typename K>
void time_it(const char* name, K launch, double bytes_total, double pieces_total, double lines_per_piece) {
  launch(20);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  launch(1000);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  const double clks = ms * 1e-3 * g_clk * 1e3;
  printf("%-58s %8.1f GB/s %6.1f B/clk/SM %5.2f pieces/clk/SM %5.2f lines/clk/SM %s\n", name, bytes_total / ms / 1e6,
         bytes_total / clks / g_sms, pieces_total / clks / g_sms, pieces_total * lines_per_piece / clks / g_sms,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int LPP, int VB, int ALIGN, int PIECE>
void run_gather(const char* name, const char* buf, size_t bytes, uint4* out, int bps, double lines_per_piece) {
  const int grid = g_sms * bps;
  const uint32_t slots = (uint32_t)((bytes - PIECE) / ALIGN);
  uint32_t p2 = 1; while (p2 * 2 <= slots) p2 *= 2;
  const double pieces = (double)grid * 256 / LPP * 8 * 1000;
  time_it(name, [&](int iters) { gather<LPP, VB, ALIGN, PIECE><<<grid, 256>>>(buf, p2, iters, out); }, pieces * PIECE, pieces,
          lines_per_piece);
}

template <int SZ, int U>
void run_bulk(const char* name, const char* buf, size_t bytes, uint4* out, int bps) {
  const int grid = g_sms * bps;
  const size_t smem = (size_t)256 * U * SZ;
  cudaFuncSetAttribute(bulk<SZ, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  uint32_t slots = (uint32_t)(bytes / SZ), p2 = 1; while (p2 * 2 <= slots) p2 *= 2;
  const double pieces = (double)grid * 256 * U * 1000;
  time_it(name, [&](int iters) { bulk<SZ, U><<<grid, 256, smem>>>(buf, p2, iters, out); }, pieces * SZ, pieces, SZ > 128 ? 2.0 : 1.0);
}

int main() {
  int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  g_sms = p.multiProcessorCount; cudaDeviceGetAttribute(&g_clk, cudaDevAttrClockRate, dev); g_clk /= 1000;
  char* buf; cudaMalloc(&buf, 64ull << 20); cudaMemset(buf, 1, 64ull << 20);
  uint4* out; cudaMalloc(&out, (size_t)g_sms * 8 * 256 * sizeof(uint4));
  printf("SMs %d, max clock %d MHz (B/clk figures use the max clock)\n", g_sms, g_clk);
  for (int bps : {4, 8}) {
    for (size_t bytes : {(size_t)64 << 20, (size_t)64 << 10}) {
      printf("-- %d blocks x 256 threads per SM, buffer %zu KB (%s)\n", bps, bytes >> 10, bytes > (1 << 20) ? "L2" : "L1");
      run_gather<4, 16, 64, 64>("A  64B piece,  4 lanes x LDG.128 (fp16 tap)", buf, bytes, out, bps, 1.0);
      run_gather<4, 8, 32, 32>("B  32B piece,  4 lanes x LDG.64  (int8 tap, r1 kernel)", buf, bytes, out, bps, 1.0);
      run_gather<2, 16, 32, 32>("C  32B piece,  2 lanes x LDG.128 (int8 tap, 16 ch/lane)", buf, bytes, out, bps, 1.0);
      run_gather<8, 16, 64, 128>("D 128B piece,  8 lanes x LDG.128, 64B-aligned (int8 2x2)", buf, bytes, out, bps, 1.5);
      run_gather<8, 16, 128, 128>("E 128B piece,  8 lanes x LDG.128, 128B-aligned (fp16 pair)", buf, bytes, out, bps, 1.0);
      run_gather<8, 8, 64, 64>("F  64B piece,  8 lanes x LDG.64  (fp16 tap over 8 lanes)", buf, bytes, out, bps, 1.0);
      run_gather<16, 16, 128, 256>("H 256B piece, 16 lanes x LDG.128, 128B-aligned (fp16 2x2)", buf, bytes, out, bps, 2.0);
    }
  }
  for (int bps : {1, 2}) {
    printf("-- TMA 1-D bulk copies into shared memory, %d blocks x 256 threads per SM, 64 MB buffer (L2)\n", bps);
    run_bulk<64, 4>("G  cp.async.bulk  64B x 4 per lane per round", buf, 64ull << 20, out, bps);
    run_bulk<128, 2>("G  cp.async.bulk 128B x 2 per lane per round", buf, 64ull << 20, out, bps);
    run_bulk<256, 2>("G  cp.async.bulk 256B x 2 per lane per round", buf, 64ull << 20, out, bps);
  }
  return 0;
}
