"""Host-buffer entry for the MSDA plugin op: ``HostMSDA`` takes the five plugin inputs in (pinned) HOST memory and
returns the output in pinned host memory — what a caller outside the GPU process boundary pays for one call.

The op is independent per batch element (one camera of SpatialCrossAttention: ``value[b]``, ``reference_points[b]``,
``sampling_offsets[b]``, ``attention_weights[b]`` -> ``out[b]``), so the call is pipelined per camera over three CUDA
streams: while camera ``b`` is being sampled, camera ``b+1`` is on its way in over PCIe and camera ``b-1`` on its way
out (PCIe is full duplex), with ``depth`` device-side slots allocated once (nothing is allocated on the call path, so
the stream-ordered caching allocator never has to wait for another stream). Back-to-back calls overlap the same way across call
boundaries. A single-stream call pays H2D + kernel + D2H back to back (11.7 ms at BEVFormer-base shapes, FP16); the
pipelined call is bound by the larger of the two copy directions (H2D: 467 MB per call).

No CPU fallback: the kernels are the sm_100a ones behind ``multi_scale_deformable_attn``.
"""
import ctypes
import math

import torch

from .functions.multi_scale_deformable_attn import multi_scale_deformable_attn_out

_cudart = None


def empty_pinned(shape, dtype=torch.float16):
    """Page-locked host tensor straight from ``cudaHostAlloc``. Measured on the B200 box (scripts/micro/numa_pin.py):
    55 GB/s host-to-device from such a buffer wherever the calling thread runs, against 40-49 GB/s from
    ``tensor.pin_memory()``, whose staging copy lets the first-touch NUMA node of the pages vary. Use it for the
    staging buffers handed to ``HostMSDA``. The allocation is owned by the returned tensor's storage: ``cudaFreeHost``
    runs when the last tensor / view over it is garbage-collected."""
    global _cudart
    if _cudart is None:
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                _cudart = ctypes.CDLL(name)
                break
            except OSError:
                continue
        else:
            raise ImportError("libcudart not found: empty_pinned needs the CUDA runtime")
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    nbytes = max(1, math.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    ptr = ctypes.c_void_p()
    err = _cudart.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0))
    if err != 0 or not ptr.value:
        raise MemoryError(f"cudaHostAlloc({nbytes}) failed with error {err}")
    address, cudart = ptr.value, _cudart

    class _PinnedBytes(ctypes.c_char * nbytes):  # torch.frombuffer holds the exporting object; its death frees the pages
        def __del__(self):
            cudart.cudaFreeHost(ctypes.c_void_p(address))

    return torch.frombuffer(_PinnedBytes.from_address(address), dtype=dtype, count=math.prod(shape)).view(shape)


class HostMSDA:
    def __init__(self, device=None, depth=3, op=multi_scale_deformable_attn_out):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth, self.op = int(depth), op
        with torch.cuda.device(self.device):
            self.s_in, self.s_k, self.s_out = (torch.cuda.Stream() for _ in range(3))
        self._slots, self._key = None, None
        self._shapes_dev, self._shapes_src = None, None
        self._turn = 0

    def _prepare(self, value, ref, off, logits):
        key = (value.shape[1:], ref.shape[1:], off.shape[1:], logits.shape[1:], value.dtype)
        if key == self._key:
            return
        dev, dt = self.device, value.dtype
        self._slots = []
        for _ in range(self.depth):
            slot = {
                "v": torch.empty((1, *value.shape[1:]), dtype=dt, device=dev),
                "r": torch.empty((1, *ref.shape[1:]), dtype=dt, device=dev),
                "o": torch.empty((1, *off.shape[1:]), dtype=dt, device=dev),
                "w": torch.empty((1, *logits.shape[1:]), dtype=dt, device=dev),
                "out": torch.empty(1, off.shape[1], value.shape[2], value.shape[3], dtype=dt, device=dev),
                "in_done": torch.cuda.Event(), "k_done": torch.cuda.Event(), "out_done": torch.cuda.Event(),
            }  # fmt: skip
            self._slots.append(slot)
        self._key = key

    def __call__(self, value, spatial_shapes, reference_points, sampling_offsets, attention_weights, out=None):
        """All tensors on the host (pin them: pageable memory makes the copies synchronous). Returns ``out`` (pinned
        host tensor [bs, num_query, heads, channels]); the work is asynchronous — call ``synchronize()`` before reading."""
        for t in (value, reference_points, sampling_offsets, attention_weights):
            if t.is_cuda:
                raise RuntimeError("HostMSDA takes host tensors; use multi_scale_deformable_attn for device tensors")
        bs = value.shape[0]
        if out is None:
            out = torch.empty(bs, sampling_offsets.shape[1], value.shape[2], value.shape[3], dtype=value.dtype).pin_memory()
        with torch.cuda.device(self.device):
            self._prepare(value, reference_points, sampling_offsets, attention_weights)
            if self._shapes_src is not spatial_shapes:  # tiny, converted once per distinct shapes tensor
                self._shapes_dev = spatial_shapes.to(device=self.device, dtype=torch.int32)
                self._shapes_src = spatial_shapes
                torch.cuda.current_stream().synchronize()
            for b in range(bs):
                slot = self._slots[self._turn % self.depth]
                self._turn += 1
                with torch.cuda.stream(self.s_in):
                    self.s_in.wait_event(slot["k_done"])  # the kernel that last read this slot's inputs has finished
                    slot["v"].copy_(value[b : b + 1], non_blocking=True)
                    slot["r"].copy_(reference_points[b : b + 1], non_blocking=True)
                    slot["o"].copy_(sampling_offsets[b : b + 1], non_blocking=True)
                    slot["w"].copy_(attention_weights[b : b + 1], non_blocking=True)
                    slot["in_done"].record(self.s_in)
                with torch.cuda.stream(self.s_k):
                    self.s_k.wait_event(slot["in_done"])
                    self.s_k.wait_event(slot["out_done"])  # the previous result of this slot has left the device
                    self.op(slot["v"], self._shapes_dev, slot["r"], slot["o"], slot["w"], slot["out"])
                    slot["k_done"].record(self.s_k)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(slot["k_done"])
                    out[b : b + 1].copy_(slot["out"], non_blocking=True)
                    slot["out_done"].record(self.s_out)
        return out

    def synchronize(self):
        self.s_in.synchronize()
        self.s_k.synchronize()
        self.s_out.synchronize()
