// Microbenchmark: how fast can one SM gather 64-byte segments (one FP16 MSDA tap) through L1 / shared memory?
// Each lane loads 16 B; 4 consecutive lanes form one 64-byte tap at a pseudo-random 64B-aligned position.
//   mode 0: LDG.128 from a buffer of `bytes` (L2-resident if <= ~100 MB, L1-resident if <= ~128 KB per SM)
//   mode 1: LDS.128 from shared memory, random taps (bank conflicts as they fall)
//   mode 2: LDS.128 from shared memory, conflict-free pattern (each quarter-warp covers all 32 banks)
//   mode 3: LDG.128, but taps of a quarter-warp pair share a 128-byte line (adjacent 64B halves)
// Prints achieved GB/s and B/clk/SM.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <int MODE>
__global__ void __launch_bounds__(256) gather(const uint4* __restrict__ buf, uint32_t taps_in_buf, int iters, uint4* out) {
  extern __shared__ uint4 smem[];
  const int lane = threadIdx.x & 31;
  const int sub = lane & 3, grp = lane >> 2;
  uint32_t s_taps = 0;
  if (MODE == 1 || MODE == 2) {
    s_taps = taps_in_buf;  // taps that fit the smem buffer
    for (uint32_t i = threadIdx.x; i < s_taps * 4; i += blockDim.x) smem[i] = buf[i];
    __syncthreads();
  }
  uint4 acc = make_uint4(0, 0, 0, 0);
  uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;  // same for the 4 lanes of a tap
  seed = seed * 2654435761u + 12345u;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      seed = seed * 1664525u + 1013904223u;  // LCG: 1 IMAD; use the high bits
      const uint32_t rnd = seed >> 8;
      uint32_t tap;  // all sizes are powers of two: masks, no modulo
      if (MODE == 0) tap = rnd & (taps_in_buf - 1);
      else if (MODE == 1) tap = rnd & (s_taps - 1);
      else if (MODE == 2) tap = ((rnd & (s_taps / 2 - 1)) * 2) | (grp & 1);  // even groups -> even taps, odd -> odd: no conflicts
      else { uint32_t pair = __shfl_sync(0xffffffffu, rnd, lane & ~7); tap = ((pair & (taps_in_buf / 2 - 1)) * 2) | (grp & 1); }
      if (MODE == 1 || MODE == 2) v[u] = smem[tap * 4 + sub];
      else v[u] = __ldg(buf + (size_t)tap * 4 + sub);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc.x ^= v[u].x; acc.y += v[u].y; acc.z ^= v[u].z; acc.w += v[u].w; }
  }
  if (acc.x == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, const uint4* buf, size_t bytes, size_t smem_bytes, uint4* out, int blocks_per_sm) {
  int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  const int sms = p.multiProcessorCount;
  const int iters = 2000;
  const uint32_t taps = (uint32_t)(((MODE == 1 || MODE == 2) ? smem_bytes : bytes) / 64);
  if (smem_bytes) cudaFuncSetAttribute(gather<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  const int grid = sms * blocks_per_sm;
  gather<MODE><<<grid, 256, smem_bytes>>>(buf, taps, 50, out);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  gather<MODE><<<grid, 256, smem_bytes>>>(buf, taps, iters, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  const double total = (double)grid * 256 * iters * 8 * 16;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  printf("%-44s %8.1f GB/s  %6.1f B/clk/SM (at %d MHz max)  %s\n", name, total / ms / 1e6, total / (ms * 1e-3) / sms / (clk * 1e3),
         clk / 1000, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  const size_t big = 64ull << 20;
  uint4* buf; cudaMalloc(&buf, 128ull << 20); cudaMemset(buf, 1, big);
  uint4* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(uint4));
  for (int bps : {2, 4, 8}) {
    printf("-- %d blocks of 256 threads per SM\n", bps);
    run<0>("LDG.128 random 64B taps, 64 MB (L2)", buf, big, 0, out, bps);
    run<0>("LDG.128 random 64B taps, 128 MB (L2+HBM)", buf, 128ull << 20, 0, out, bps);
    run<0>("LDG.128 random 64B taps, 16 MB (L2)", buf, 16 << 20, 0, out, bps);
    run<0>("LDG.128 random 64B taps, 64 KB (L1)", buf, 64 << 10, 0, out, bps);
    run<3>("LDG.128 tap pairs sharing a 128B line, 64 MB", buf, big, 0, out, bps);
    run<3>("LDG.128 tap pairs sharing a 128B line, 64 KB", buf, 64 << 10, 0, out, bps);
    if (bps * 32 * 1024 <= 220 * 1024) {
      run<1>("LDS.128 random 64B taps (32 KB smem)", buf, big, 32 << 10, out, bps);
      run<2>("LDS.128 conflict-free 64B taps (32 KB smem)", buf, big, 32 << 10, out, bps);
    }
  }
  return 0;
}
