// peer_reduce.cu — reduce-scatter of the BEV accumulator inside a camera group over NVLink peer memory.
//
// Multi-GPU spatial cross-attention (reference: slots = (queries * bev_mask).sum(0), spatial_cross_attention.py:270, which
// the reference runs on one GPU): the ranks of a camera group hold partial sums P_r[rows, width] (fp32, their own
// cameras only) for the SAME query tile. Rank `me` owns rows [row0, row0 + rows): it pulls that slice of every peer's
// partial straight out of the peer's HBM (ld.global over the NVLink aperture, torch symmetric memory supplies the
// mapped pointers), adds its own and writes the final rows. No NCCL call, no staging copy, no wire-format conversion
// kernel; one launch per step and rank.
//
// Handshake (no host involvement): a step counter per (rank, peer). Block 0 publishes "my partial of step `epoch` is
// final" into every peer's flag row with a system-scope release store — the gather kernels that produced the partial
// precede this kernel in stream order — and every block waits (acquire) until all peers have published theirs.
// Partials are double-buffered by step parity; this kernel zeroes the NEXT step's buffer after the handshake: a peer
// that has published step k has finished its step k-1 kernel, so nobody reads that buffer any more.
#include "common.cuh"

namespace b200 {

constexpr int kMaxPeers = 8;

struct PeerReduceParams {
  const float *part[kMaxPeers];  // peer r's partial buffer of this step (index `me` = the local one)
  uint32_t *flags[kMaxPeers];    // peer r's flag row: uint32[kMaxPeers], slot j is written by peer j
  int n, me;
  uint32_t epoch;
  long long elem0, elems;  // owned slice, in floats (multiples of 4)
  void *out;
  int out_half;
  float *zero;  // local buffer of the next step (may be null)
  long long zero_elems;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data must come from the peer's memory, not from a stale line of this SM's L1 (peer addresses bypass the local L2)
__device__ __forceinline__ float4 ld_peer(const float *p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) peer_reduce_kernel(const PeerReduceParams p) {
  const int t = threadIdx.x;
  if (t < p.n && t != p.me) {
    if (blockIdx.x == 0) {
      __threadfence_system();
      st_release_sys(p.flags[t] + p.me, p.epoch);
    }
    const uint32_t *mine = p.flags[p.me] + t;
    while (static_cast<int32_t>(ld_acquire_sys(mine) - p.epoch) < 0) __nanosleep(64);
  }
  __syncthreads();

  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long first = static_cast<long long>(blockIdx.x) * blockDim.x + t;
  for (long long i = first; i < p.elems / 4; i += stride) {
    const long long e = p.elem0 + i * 4;
    float4 acc = *reinterpret_cast<const float4 *>(p.part[p.me] + e);
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r) {
      if (r < p.n && r != p.me) {
        const float4 v = ld_peer(p.part[r] + e);
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
      }
    }
    if (p.out_half) {
      uint2 h = make_uint2(f2_to_h2(acc.x, acc.y), f2_to_h2(acc.z, acc.w));
      *reinterpret_cast<uint2 *>(static_cast<__half *>(p.out) + i * 4) = h;
    } else {
      *reinterpret_cast<float4 *>(static_cast<float *>(p.out) + i * 4) = acc;
    }
  }
  if (p.zero != nullptr)
    for (long long i = first; i < p.zero_elems / 4; i += stride)
      *reinterpret_cast<float4 *>(p.zero + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_sca_peer_reduce(const void *const *partials, void *const *flags, int group_size, int my_index,
                                    unsigned int epoch, long long first_elem, long long num_elems, void *out,
                                    int out_is_half, float *zero_next, long long zero_elems, void *stream) {
  if (!partials || !flags || !out || group_size < 1 || group_size > kMaxPeers || my_index < 0 || my_index >= group_size)
    return B200_ERR_BAD_PARAM;
  if (first_elem < 0 || num_elems < 0 || (first_elem % 4) || (num_elems % 4) || (zero_elems % 4) || zero_elems < 0)
    return B200_ERR_BAD_PARAM;
  PeerReduceParams p{};
  for (int r = 0; r < group_size; ++r) {
    if (!partials[r] || !flags[r] || reinterpret_cast<uintptr_t>(partials[r]) % 16) return B200_ERR_BAD_PARAM;
    p.part[r] = static_cast<const float *>(partials[r]);
    p.flags[r] = static_cast<uint32_t *>(flags[r]);
  }
  if (reinterpret_cast<uintptr_t>(out) % 16 || reinterpret_cast<uintptr_t>(zero_next) % 16) return B200_ERR_BAD_PARAM;
  p.n = group_size, p.me = my_index, p.epoch = epoch, p.elem0 = first_elem, p.elems = num_elems;
  p.out = out, p.out_half = out_is_half, p.zero = zero_next, p.zero_elems = zero_next ? zero_elems : 0;
  const long long work = (num_elems > p.zero_elems ? num_elems : p.zero_elems) / 4;
  long long blocks = (work + 255) / 256;
  // every block spins in the handshake: keep the grid within one wave so that no block waits behind a spinning one
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  peer_reduce_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch();
}
