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
  const float *part[2][kMaxPeers];  // [step parity][peer r]: r's partial buffer of that step (index `me` = the local one)
  uint32_t *flags[kMaxPeers];       // peer r's flag row: uint32[kMaxPeers + 2], slot j is written by peer j
  int n, me;
  uint32_t epoch;  // host-supplied step number, or 0: read it from the device (flags[me][kMaxPeers] + 1)
  long long elem0, elems;  // owned slice, in floats (multiples of 4)
  void *out;
  int out_half;
  float *zero[2];  // [step parity]: local buffer of the NEXT step to zero-fill (may be null)
  long long zero_elems;
  float *staging;  // overlapped step: the peers' contribution to the owned slice, [elems] floats, local memory
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

// MODE 0: the whole exchange in one launch (handshake, out = own + peers, zero the next buffer).
// MODE 1 (overlapped step, first half): handshake, then staging = sum of the PEERS' partials of the owned slice; runs on a
//         side stream while this rank still samples its own rows. MODE 2 (second half): out = own partial + staging,
//         zero the next buffer; no handshake (stream order after both).
template <int MODE>
__global__ void __launch_bounds__(256) peer_reduce_kernel(const PeerReduceParams p) {
  const int t = threadIdx.x;
  // Step number: from the host, or (CUDA-graph replays: identical launch parameters every step) from this rank's own
  // counter in device memory, which the last block of the previous handshake launch advanced.
  uint32_t *ctr = p.flags[p.me] + kMaxPeers;  // [0] completed steps, [1] blocks done in this launch
  const uint32_t epoch = p.epoch ? p.epoch : *reinterpret_cast<volatile uint32_t *>(ctr) + (MODE == 2 ? 0u : 1u);
  const int par = static_cast<int>((epoch - 1u) & 1u);
  if (MODE != 2) {
    if (t < p.n && t != p.me) {
      if (blockIdx.x == 0) {
        __threadfence_system();
        st_release_sys(p.flags[t] + p.me, epoch);
      }
      const uint32_t *mine = p.flags[p.me] + t;
      while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) __nanosleep(32);
    }
    __syncthreads();
  }

  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long first = static_cast<long long>(blockIdx.x) * blockDim.x + t;
  const long long n4 = p.elems / 4;
  // NVLink round trips are ~3 us: keep kUnroll x (n - 1) independent 16-byte peer loads in flight per thread
  constexpr int kUnroll = 4;
  for (long long i0 = first; i0 < n4; i0 += stride * kUnroll) {
    float4 acc[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = i0 + u * stride;
      acc[u] = (MODE != 1 && i < n4) ? *reinterpret_cast<const float4 *>(p.part[par][p.me] + p.elem0 + i * 4)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (MODE == 2) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long i = i0 + u * stride;
        if (i < n4) {
          const float4 v = *reinterpret_cast<const float4 *>(p.staging + i * 4);
          acc[u].x += v.x, acc[u].y += v.y, acc[u].z += v.z, acc[u].w += v.w;
        }
      }
    } else {
#pragma unroll 1
      for (int r = 0; r < p.n; ++r) {
        if (r == p.me) continue;
        const float *src = p.part[par][r] + p.elem0;
        float4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const long long i = i0 + u * stride;
          v[u] = i < n4 ? ld_peer(src + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) acc[u].x += v[u].x, acc[u].y += v[u].y, acc[u].z += v[u].z, acc[u].w += v[u].w;
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i >= n4) break;
      const float4 a = acc[u];
      if (MODE != 1 && p.out_half) {
        const uint2 h = make_uint2(f2_to_h2(a.x, a.y), f2_to_h2(a.z, a.w));
        *reinterpret_cast<uint2 *>(static_cast<__half *>(p.out) + i * 4) = h;
      } else {
        *reinterpret_cast<float4 *>((MODE == 1 ? p.staging : static_cast<float *>(p.out)) + i * 4) = a;
      }
    }
  }
  if (MODE != 1) {
    float *zero = p.zero[par];
    if (zero != nullptr)
      for (long long i = first; i < p.zero_elems / 4; i += stride)
        *reinterpret_cast<float4 *>(zero + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (MODE != 2 && p.epoch == 0u) {  // device-side step counter: the last block to finish advances it
    __syncthreads();
    if (t == 0) {
      __threadfence();
      if (atomicAdd(ctr + 1, 1u) == gridDim.x - 1) {
        ctr[1] = 0u;
        __threadfence();
        *reinterpret_cast<volatile uint32_t *>(ctr) = epoch;
      }
    }
  }
}

}  // namespace b200

using namespace b200;

static int launch_peer_reduce(PeerReduceParams &p, cudaStream_t stream, int mode = 0) {
  const long long work = (mode != 1 && p.zero_elems > p.elems ? p.zero_elems : p.elems) / 4;
  long long blocks = (work + 255) / 256 / 4;  // 4 x 16 bytes per thread and peer in flight
  // every block spins in the handshake: keep the grid within one wave so that no block waits behind a spinning one
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  if (mode == 1)
    peer_reduce_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
  else if (mode == 2)
    peer_reduce_kernel<2><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
  else
    peer_reduce_kernel<0><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
  return check_launch();
}

extern "C" int b200_sca_peer_reduce(const void *const *partials, void *const *flags, int group_size, int my_index,
                                    unsigned int epoch, long long first_elem, long long num_elems, void *out,
                                    int out_is_half, float *zero_next, long long zero_elems, void *stream) {
  if (!partials || !flags || !out || group_size < 1 || group_size > kMaxPeers || my_index < 0 || my_index >= group_size)
    return B200_ERR_BAD_PARAM;
  if (epoch == 0u) return B200_ERR_BAD_PARAM;  // steps are numbered from 1
  if (first_elem < 0 || num_elems < 0 || (first_elem % 4) || (num_elems % 4) || (zero_elems % 4) || zero_elems < 0)
    return B200_ERR_BAD_PARAM;
  PeerReduceParams p{};
  for (int r = 0; r < group_size; ++r) {
    if (!partials[r] || !flags[r] || reinterpret_cast<uintptr_t>(partials[r]) % 16) return B200_ERR_BAD_PARAM;
    p.part[0][r] = p.part[1][r] = static_cast<const float *>(partials[r]);
    p.flags[r] = static_cast<uint32_t *>(flags[r]);
  }
  if (reinterpret_cast<uintptr_t>(out) % 16 || reinterpret_cast<uintptr_t>(zero_next) % 16) return B200_ERR_BAD_PARAM;
  p.n = group_size, p.me = my_index, p.epoch = epoch, p.elem0 = first_elem, p.elems = num_elems;
  p.out = out, p.out_half = out_is_half, p.zero[0] = p.zero[1] = zero_next, p.zero_elems = zero_next ? zero_elems : 0;
  return launch_peer_reduce(p, static_cast<cudaStream_t>(stream));
}

extern "C" int b200_sca_peer_reduce_auto(const void *const *partials_even, const void *const *partials_odd,
                                         void *const *flags, int group_size, int my_index, long long first_elem,
                                         long long num_elems, void *out, int out_is_half, long long zero_elems,
                                         void *stream) {
  if (!partials_even || !partials_odd || !flags || !out || group_size < 1 || group_size > kMaxPeers || my_index < 0 ||
      my_index >= group_size)
    return B200_ERR_BAD_PARAM;
  if (first_elem < 0 || num_elems < 0 || (first_elem % 4) || (num_elems % 4) || (zero_elems % 4) || zero_elems < 0)
    return B200_ERR_BAD_PARAM;
  PeerReduceParams p{};
  for (int r = 0; r < group_size; ++r) {
    if (!partials_even[r] || !partials_odd[r] || !flags[r] || reinterpret_cast<uintptr_t>(partials_even[r]) % 16 ||
        reinterpret_cast<uintptr_t>(partials_odd[r]) % 16)
      return B200_ERR_BAD_PARAM;
    p.part[0][r] = static_cast<const float *>(partials_even[r]);  // steps 1, 3, 5, ... ((step - 1) & 1 == 0)
    p.part[1][r] = static_cast<const float *>(partials_odd[r]);
    p.flags[r] = static_cast<uint32_t *>(flags[r]);
  }
  if (reinterpret_cast<uintptr_t>(out) % 16) return B200_ERR_BAD_PARAM;
  p.n = group_size, p.me = my_index, p.epoch = 0u, p.elem0 = first_elem, p.elems = num_elems;
  p.out = out, p.out_half = out_is_half, p.zero_elems = zero_elems;
  // after reducing the buffer of parity k, zero the OTHER buffer: it is the next step's accumulator
  p.zero[0] = zero_elems ? const_cast<float *>(p.part[1][my_index]) : nullptr;
  p.zero[1] = zero_elems ? const_cast<float *>(p.part[0][my_index]) : nullptr;
  return launch_peer_reduce(p, static_cast<cudaStream_t>(stream));
}

// Overlapped step (see peer_reduce_kernel MODE 1 / 2): `b200_sca_peer_pull_auto` on a side stream right after the launch
// that samples the rows the PEERS own, `b200_sca_peer_add_auto` on the main stream after the launch that samples this
// rank's own rows and after the pull. Same buffers / flags / device-side step counter as b200_sca_peer_reduce_auto.
extern "C" int b200_sca_peer_pull_auto(const void *const *partials_even, const void *const *partials_odd,
                                       void *const *flags, int group_size, int my_index, long long first_elem,
                                       long long num_elems, float *staging, void *stream) {
  if (!partials_even || !partials_odd || !flags || !staging || group_size < 2 || group_size > kMaxPeers || my_index < 0 ||
      my_index >= group_size)
    return B200_ERR_BAD_PARAM;
  if (first_elem < 0 || num_elems < 0 || (first_elem % 4) || (num_elems % 4) || reinterpret_cast<uintptr_t>(staging) % 16)
    return B200_ERR_BAD_PARAM;
  PeerReduceParams p{};
  for (int r = 0; r < group_size; ++r) {
    if (!partials_even[r] || !partials_odd[r] || !flags[r]) return B200_ERR_BAD_PARAM;
    p.part[0][r] = static_cast<const float *>(partials_even[r]);
    p.part[1][r] = static_cast<const float *>(partials_odd[r]);
    p.flags[r] = static_cast<uint32_t *>(flags[r]);
  }
  p.n = group_size, p.me = my_index, p.epoch = 0u, p.elem0 = first_elem, p.elems = num_elems, p.staging = staging;
  return launch_peer_reduce(p, static_cast<cudaStream_t>(stream), 1);
}

extern "C" int b200_sca_peer_add_auto(const void *partial_even, const void *partial_odd, void *flags_local, int my_index,
                                      long long first_elem, long long num_elems, const float *staging, void *out,
                                      int out_is_half, long long zero_elems, void *stream) {
  if (!partial_even || !partial_odd || !flags_local || !staging || !out || my_index < 0 || my_index >= kMaxPeers)
    return B200_ERR_BAD_PARAM;
  if (first_elem < 0 || num_elems < 0 || (first_elem % 4) || (num_elems % 4) || (zero_elems % 4) || zero_elems < 0 ||
      reinterpret_cast<uintptr_t>(out) % 16 || reinterpret_cast<uintptr_t>(staging) % 16)
    return B200_ERR_BAD_PARAM;
  PeerReduceParams p{};
  p.part[0][my_index] = static_cast<const float *>(partial_even);
  p.part[1][my_index] = static_cast<const float *>(partial_odd);
  p.flags[my_index] = static_cast<uint32_t *>(flags_local);
  p.n = 1, p.me = my_index, p.epoch = 0u, p.elem0 = first_elem, p.elems = num_elems;
  p.staging = const_cast<float *>(staging), p.out = out, p.out_half = out_is_half, p.zero_elems = zero_elems;
  p.zero[0] = zero_elems ? const_cast<float *>(p.part[1][my_index]) : nullptr;
  p.zero[1] = zero_elems ? const_cast<float *>(p.part[0][my_index]) : nullptr;
  return launch_peer_reduce(p, static_cast<cudaStream_t>(stream), 2);
}
