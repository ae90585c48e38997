"""oracle/msda.py — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Python drivers for the MSDA oracles:

* ``msda_f32`` / ``msda_i8_dequant`` / ``msda_i8_refemu`` — ctypes wrappers over oracle/msda_oracle.c (the literal
  per-output restatement of the plugin kernels, reference file:line cited there).
* ``msda_torch_port`` — the reference's *CPU* implementation restated: per-level ``F.grid_sample`` + weighted sum,
  following det2trt/models/utils/trt_ops.py:4-85, fed through the plugin-signature adapter of
  det2trt/models/functions/multi_scale_deformable_attn.py:58-92 (locations = ref + off/(W,H); softmax over L*P).
  This is the function bench.py times as ``cpu_baseline`` / ``--impl reference`` (kind "port": the reference's Python
  file cannot travel to the GPU box).
* ``RefKernels`` — ctypes wrapper over oracle/_ref/libref_kernels.so (the reference's own CUDA kernels, GPU only).
"""
import ctypes
import os

import numpy as np

from . import REF_LIB, lib

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i8p = ctypes.POINTER(ctypes.c_int8)

INDEX_DTYPE = np.dtype([("in_range", "<i4"), ("h_low", "<i4"), ("w_low", "<i4"), ("tap_mask", "<i4")])


def _dims(value, shapes, ref, off, logits):
    B, S, M, C = value.shape
    L = shapes.shape[0]
    Q = off.shape[1]
    G = ref.shape[-1] // 2
    NP = logits.shape[-1]
    assert NP % L == 0
    P = NP // L
    assert ref.shape[:2] == (B, Q) and off.shape == (B, Q, M, NP * 2) and logits.shape == (B, Q, M, NP)
    assert int((shapes[:, 0].astype(np.int64) * shapes[:, 1]).sum()) == S
    return B, S, M, C, L, Q, P, G


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def msda_f32(value, shapes, ref, off, logits, return_index=False):
    """FP32 plugin semantics on numpy arrays (any float dtype in; evaluated in float32). Returns out[B,Q,M,C]."""
    value, ref, off, logits = (_c(x, np.float32) for x in (value, ref, off, logits))
    shapes = _c(shapes, np.int32)
    B, S, M, C, L, Q, P, G = _dims(value, shapes, ref, off, logits)
    out = np.empty((B, Q, M, C), np.float32)
    idx = np.empty((B, Q, M, L * P), INDEX_DTYPE) if return_index else None
    lib().oracle_msda_f32(
        value.ctypes.data_as(_f32p), shapes.ctypes.data_as(_i32p), ref.ctypes.data_as(_f32p),
        off.ctypes.data_as(_f32p), logits.ctypes.data_as(_f32p),
        B, S, M, C, L, Q, P, G, out.ctypes.data_as(_f32p),
        idx.ctypes.data_as(ctypes.c_void_p) if return_index else None,
    )  # fmt: skip
    return (out, idx) if return_index else out


def _i8_call(fn, value, sv, shapes, ref, off, so, logits, sw, sout, want_real):
    value, off, logits = (_c(x, np.int8) for x in (value, off, logits))
    ref = _c(ref, np.float32)
    shapes = _c(shapes, np.int32)
    B, S, M, C, L, Q, P, G = _dims(value, shapes, ref, off, logits)
    out = np.empty((B, Q, M, C), np.int8)
    args = [
        value.ctypes.data_as(_i8p), ctypes.c_float(sv), shapes.ctypes.data_as(_i32p), ref.ctypes.data_as(_f32p),
        off.ctypes.data_as(_i8p), ctypes.c_float(so), logits.ctypes.data_as(_i8p), ctypes.c_float(sw),
        B, S, M, C, L, Q, P, G, out.ctypes.data_as(_i8p), ctypes.c_float(sout),
    ]  # fmt: skip
    real = None
    if want_real is not None:
        real = np.empty((B, Q, M, C), np.float32) if want_real else None
        args.append(real.ctypes.data_as(_f32p) if want_real else None)
    fn(*args)
    return out, real


def msda_i8_dequant(value, sv, shapes, ref, off, so, logits, sw, sout, return_real=False):
    """INT8 'in-register dequant' definition: fp32 formulas on dequantised inputs, T2int8(result/scale_out)."""
    out, real = _i8_call(lib().oracle_msda_i8_dequant, value, sv, shapes, ref, off, so, logits, sw, sout, return_real)
    return (out, real) if return_real else out


def msda_i8_refemu(value, sv, shapes, ref, off, so, logits, sw, sout):
    """Emulation of the reference's quantised-intermediate INT8 kernel (float ref points)."""
    assert (logits.shape[-1] // shapes.shape[0]) % 4 == 0, "reference INT8 kernel needs P % 4 == 0"
    out, _ = _i8_call(lib().oracle_msda_i8_refemu, value, sv, shapes, ref, off, so, logits, sw, sout, None)
    return out


def num_threads():
    return int(lib().oracle_num_threads())


# ----------------------------------------------------------------------------------------------------------------
# The reference's CPU path, restated (torch; multi-threaded through ATen's grid_sample)
# ----------------------------------------------------------------------------------------------------------------
def plugin_inputs_to_locations(shapes, ref, off, logits):
    """Adapter of functions/multi_scale_deformable_attn.py:58-92: plugin inputs -> (normalised sampling locations
    [B,Q,M,L,P,2], softmaxed attention weights [B,Q,M,L,P])."""
    import torch

    B, Q, M, _ = off.shape
    L = shapes.shape[0]
    G = ref.shape[-1] // 2
    P = logits.shape[-1] // L
    wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).to(off.dtype)  # (W, H) per level
    o = off.reshape(B, Q, M, L, P // G, G, 2) / wh.view(1, 1, 1, L, 1, 1, 2)
    loc = ref.reshape(B, Q, 1, 1, 1, G, 2) + o
    att = logits.reshape(B, Q, M, L * P).softmax(-1)
    return loc.reshape(B, Q, M, L, P, 2), att.reshape(B, Q, M, L, P)


def msda_locations_torch(value, shapes, loc, att):
    """trt_ops.py:4-85 restated: for every level, bilinear grid_sample (zeros padding, align_corners=False) of the
    [B*M, C, H, W] view of that level at 2*loc-1, times the attention weights, summed over points and levels."""
    import torch
    import torch.nn.functional as F

    B, S, M, C = value.shape
    Q = loc.shape[1]
    out = value.new_zeros(B * M, C, Q)
    start = 0
    for lvl, (H, W) in enumerate(shapes.tolist()):
        v = value[:, start : start + H * W].permute(0, 2, 3, 1).reshape(B * M, C, H, W)
        g = (2 * loc[:, :, :, lvl] - 1).permute(0, 2, 1, 3, 4).reshape(B * M, Q, -1, 2)
        s = F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False)  # [B*M, C, Q, P]
        a = att[:, :, :, lvl].permute(0, 2, 1, 3).reshape(B * M, 1, Q, -1)
        out += (s * a).sum(-1)
        start += H * W
    return out.view(B, M, C, Q).permute(0, 3, 1, 2).contiguous()  # [B, Q, M, C]


def msda_torch_port(value, shapes, ref, off, logits):
    """Plugin signature -> [B,Q,M,C] through the reference's CPU algorithm (float32 torch tensors on CPU)."""
    loc, att = plugin_inputs_to_locations(shapes, ref, off, logits)
    return msda_locations_torch(value, shapes, loc, att)


# ----------------------------------------------------------------------------------------------------------------
# oracle/_ref : the reference's own CUDA kernels (GPU box only)
# ----------------------------------------------------------------------------------------------------------------
class RefKernels:
    """Calls the reference's launchers (oracle/ref_shim.cu) on torch CUDA tensors."""

    def __init__(self):
        if not os.path.exists(REF_LIB):
            raise FileNotFoundError(f"{REF_LIB} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = ctypes.CDLL(REF_LIB)

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr())

    @staticmethod
    def _stream():
        import torch

        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _msda_dims(self, value, shapes, ref, off, logits):
        B, S, M, C = value.shape
        L = shapes.shape[0]
        Q = off.shape[1]
        G = ref.shape[-1] // 2
        P = logits.shape[-1] // L
        return B, S, M, C, L, Q, P, G

    def msda(self, value, shapes, ref, off, logits, variant="f32"):
        import torch

        assert value.is_cuda and shapes.dtype == torch.int32
        d = self._msda_dims(value, shapes, ref, off, logits)
        out = torch.empty(d[0], d[5], d[2], d[3], dtype=value.dtype, device=value.device)
        fn = {"f32": self.lib.ref_msda_f32, "f16": self.lib.ref_msda_f16, "f16_h2": self.lib.ref_msda_f16_h2}[variant]
        fn(self._p(value), self._p(shapes), self._p(ref), self._p(off), self._p(logits), *d, self._p(out),
           self._stream())  # fmt: skip
        return out

    def msda_i8(self, value, sv, shapes, ref, off, so, logits, sw, sout):
        import torch

        d = self._msda_dims(value, shapes, ref, off, logits)
        out = torch.empty(d[0], d[5], d[2], d[3], dtype=torch.int8, device=value.device)
        fn = self.lib.ref_msda_i8_h2ref if ref.dtype == torch.float16 else self.lib.ref_msda_i8_f32ref
        fn(self._p(value), ctypes.c_float(sv), self._p(shapes), self._p(ref), self._p(off), ctypes.c_float(so),
           self._p(logits), ctypes.c_float(sw), *d, self._p(out), ctypes.c_float(sout), self._stream())  # fmt: skip
        return out
