"""oracle/grid_sampler.py — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

ctypes driver for oracle/grid_sampler_oracle.c (the reference GridSampler2DTRT FP32 kernel restated) and a torch
restatement of the reference binding's forward (det2trt/models/functions/grid_sampler.py:28-32)."""
import ctypes

import numpy as np

from . import lib

_f32p = ctypes.POINTER(ctypes.c_float)
INDEX_DTYPE = np.dtype([("ix", "<i4"), ("iy", "<i4")])


def grid_sample_2d(inp, grid, interp, padding, align, return_index=False):
    """inp [N,C,Hi,Wi], grid [N,2,Ho,Wo] in [-10,10]; interp 0/1/2 = bilinear/nearest/bicubic; padding 0/1/2 =
    zeros/border/reflection. float32 evaluation of the kernel's formulas."""
    inp = np.ascontiguousarray(inp, np.float32)
    grid = np.ascontiguousarray(grid, np.float32)
    N, C, Hi, Wi = inp.shape
    _, two, Ho, Wo = grid.shape
    assert two == 2 and grid.shape[0] == N
    out = np.empty((N, C, Ho, Wo), np.float32)
    idx = np.empty((N, Ho, Wo), INDEX_DTYPE) if return_index else None
    lib().oracle_grid_sample_2d_f32(inp.ctypes.data_as(_f32p), grid.ctypes.data_as(_f32p), out.ctypes.data_as(_f32p),
                                    N, C, Hi, Wi, Ho, Wo, int(interp), int(padding), int(bool(align)),
                                    idx.ctypes.data_as(ctypes.c_void_p) if return_index else None)  # fmt: skip
    return (out, idx) if return_index else out


def grid_sampler_torch_port(inp, grid, interp, padding, align):
    """The reference binding's forward: aten.grid_sampler(input, grid.permute(0,2,3,1)/10, mode, pad, align) for 4-D,
    permute(0,2,3,4,1) for 5-D (grid_sampler.py:28-32, :84-88)."""
    import torch

    perm = (0, 2, 3, 1) if grid.dim() == 4 else (0, 2, 3, 4, 1)
    return torch.ops.aten.grid_sampler(inp, grid.permute(*perm) / 10, interp, padding, bool(align))
