"""oracle/rotate.py — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

ctypes driver for oracle/rotate_oracle.c (the reference RotateTRT FP32 kernel restated), numpy helpers for the FP16 /
INT8 forms of the same arithmetic, and a torch restatement of the reference binding's forward
(det2trt/models/functions/rotate.py:12-84)."""
import ctypes
import math
import os

import numpy as np

from . import REF_LIB, lib

_f32p = ctypes.POINTER(ctypes.c_float)
_f = ctypes.c_float


def rotate(img, angle, center, interp, trig=None, return_index=False):
    """img [C,H,W]; angle degrees; center (x, y); interp 0 bilinear / 1 nearest. float32 evaluation of the kernel's
    formulas. trig = (cos, sin) of the rotation as float32 to bypass libm (see rotate_oracle.c header)."""
    img = np.ascontiguousarray(img, np.float32)
    C, H, W = img.shape
    out = np.empty_like(img)
    xy = np.empty((H, W, 2), np.float32) if return_index else None
    c, s = (float(trig[0]), float(trig[1])) if trig is not None else (0.0, 0.0)
    lib().oracle_rotate_f32(img.ctypes.data_as(_f32p), _f(float(angle)), _f(float(center[0])), _f(float(center[1])),
                            C, H, W, int(interp), int(trig is not None), _f(c), _f(s), out.ctypes.data_as(_f32p),
                            xy.ctypes.data_as(_f32p) if return_index else None)  # fmt: skip
    return (out, xy) if return_index else out


def source_indices(angle, center, H, W, trig=None):
    xy = np.empty((H, W, 2), np.float32)
    c, s = (float(trig[0]), float(trig[1])) if trig is not None else (0.0, 0.0)
    lib().oracle_rotate_f32(None, _f(float(angle)), _f(float(center[0])), _f(float(center[1])), 0, H, W, 0,
                            int(trig is not None), _f(c), _f(s), None, xy.ctypes.data_as(_f32p))  # fmt: skip
    return xy


def rotate_f16(img_h, angle_h, center_h, interp):
    """FP16 form: inputs already rounded to float16, FP32 arithmetic, one rounding of the output to float16."""
    out = rotate(img_h.astype(np.float32), float(np.float16(angle_h)), np.asarray(center_h, np.float16).astype(np.float32),
                 interp)  # fmt: skip
    return out.astype(np.float16)


def t2int8(x):
    """T2int8<float> (rotateKernel.cu:25-29): saturate to [-128, 127], then round half away from zero."""
    x = np.clip(np.asarray(x, np.float32), -128.0, 127.0)
    return np.trunc(x + np.where(x > 0, np.float32(0.5), np.float32(-0.5))).astype(np.int8)


def rotate_i8_dequant(img_q, scale_i, angle, center, interp, scale_o):
    """INT8 form: dequantise (q * scale_i), FP32 kernel arithmetic, requantise once with T2int8(v / scale_o)."""
    real = img_q.astype(np.float32) * np.float32(scale_i)
    out = rotate(real, angle, center, interp)
    return t2int8(out * (np.float32(1.0) / np.float32(scale_o)))


def rotate_torch_port(img, angle, center, interp):
    """The reference binding's forward (rotate.py:12-84): theta from cos/sin of -angle, base grid by linspace,
    rescaled theta, bmm, aten.grid_sampler(img, grid, interp, zeros, align_corners=False)."""
    import torch

    oh, ow = img.shape[-2:]
    cx = center[0] - center[0].new_tensor(ow * 0.5)
    cy = center[1] - center[1].new_tensor(oh * 0.5)
    a = -angle * math.pi / 180
    cos, sin = torch.cos(a), torch.sin(a)
    theta = torch.stack([cos, sin, -cx * cos - cy * sin + cx, -sin, cos, cx * sin - cy * cos + cy]).view(1, 2, 3)
    base = torch.empty(1, oh, ow, 3, dtype=theta.dtype, device=theta.device)
    base[..., 0] = torch.linspace(-ow * 0.5 + 0.5, ow * 0.5 - 0.5, steps=ow, device=theta.device).expand(1, oh, ow)
    base[..., 1] = torch.linspace(-oh * 0.5 + 0.5, oh * 0.5 - 0.5, steps=oh, device=theta.device).unsqueeze(-1).expand(1, oh, ow)  # fmt: skip
    base[..., 2].fill_(1)
    rt = 2 * theta.transpose(1, 2)
    rt[..., 0] /= ow
    rt[..., 1] /= oh
    grid = base.view(1, oh * ow, 3).bmm(rt).view(1, oh, ow, 2)
    return torch.grid_sampler(img.unsqueeze(0).to(grid.dtype), grid, interp, 0, False).squeeze(0).to(img.dtype)


class RefRotate:
    """Calls the reference's own launchers (oracle/ref_shim.cu: ref_rotate / ref_rotate_int8) on torch CUDA tensors."""

    def __init__(self):
        if not os.path.exists(REF_LIB):
            raise FileNotFoundError(f"{REF_LIB} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = ctypes.CDLL(REF_LIB)
        if not hasattr(self.lib, "ref_rotate"):
            raise FileNotFoundError(f"{REF_LIB} predates ref_rotate: rebuild with `make -C oracle ref`")

    def rotate(self, img, angle, center, interp, variant="f32"):
        """variant f32 / f16 ([C,H,W] kLINEAR) / f16_h2 (img already packed kCHW2 [ceil(C/2),H,W,2], C given by
        img.shape[0]*2)."""
        import torch

        dt = {"f32": 0, "f16": 1, "f16_h2": 2}[variant]
        C = img.shape[0] * (2 if dt == 2 else 1)
        dims = (ctypes.c_int * 3)(C, img.shape[1], img.shape[2])
        out = torch.empty_like(img)
        self.lib.ref_rotate(dt, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(img.data_ptr()),
                            ctypes.c_void_p(angle.data_ptr()), ctypes.c_void_p(center.data_ptr()), dims, int(interp),
                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))  # fmt: skip
        return out
