"""oracle/dcn.py — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

DCNv2 (modulated deformable convolution) oracle: deformable im2col in C (oracle/dcn_oracle.c, reference kernel
restated) + GEMM/bias in numpy, following ModulatedDeformConvForwardCUDAKernel<float>
(modulatedDeformableConv2dKernel.cu:695-760)."""
import ctypes

import numpy as np

from . import lib

_f32p = ctypes.POINTER(ctypes.c_float)


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def out_size(H, W, kh, kw, stride, padding, dilation):
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    return (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1, (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1


def im2col(x, offset, mask, kh, kw, stride=1, padding=0, dilation=1, deform_groups=1):
    """x [N,C,H,W], offset [N,dg*2*kh*kw,Ho,Wo], mask [N,dg*kh*kw,Ho,Wo] -> columns [N, C*kh*kw, Ho*Wo] (float32)."""
    x, offset, mask = (np.ascontiguousarray(a, np.float32) for a in (x, offset, mask))
    N, C, H, W = x.shape
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    Ho, Wo = out_size(H, W, kh, kw, stride, padding, dilation)
    assert offset.shape == (N, deform_groups * 2 * kh * kw, Ho, Wo) and mask.shape == (N, deform_groups * kh * kw, Ho, Wo)
    col = np.empty((N, C * kh * kw, Ho * Wo), np.float32)
    for n in range(N):
        lib().oracle_dcn_im2col_f32(x[n].ctypes.data_as(_f32p), offset[n].ctypes.data_as(_f32p),
                                    mask[n].ctypes.data_as(_f32p), col[n].ctypes.data_as(_f32p), C, H, W, kh, kw, ph,
                                    pw, sh, sw, dh, dw, deform_groups, Ho, Wo)  # fmt: skip
    return col, (Ho, Wo)


def modulated_deformable_conv2d(x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                                deform_groups=1):
    """Reference semantics in float32 storage with float64 GEMM accumulation. weight [Co, Ci/groups, kh, kw]."""
    weight = np.asarray(weight, np.float32)
    Co, Cig, kh, kw = weight.shape
    col, (Ho, Wo) = im2col(x, offset, mask, kh, kw, stride, padding, dilation, deform_groups)
    N = col.shape[0]
    k = Cig * kh * kw
    out = np.empty((N, Co, Ho * Wo), np.float32)
    for g in range(groups):
        wg = weight[g * (Co // groups) : (g + 1) * (Co // groups)].reshape(Co // groups, k).astype(np.float64)
        cg = col[:, g * k : (g + 1) * k].astype(np.float64)
        out[:, g * (Co // groups) : (g + 1) * (Co // groups)] = np.einsum("mk,nkp->nmp", wg, cg).astype(np.float32)
    if bias is not None:
        out += np.asarray(bias, np.float32)[None, :, None]
    return out.reshape(N, Co, Ho, Wo)
