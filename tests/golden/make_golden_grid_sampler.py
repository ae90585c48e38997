"""Generates tests/golden/grid_sampler_*.npz by running the REFERENCE's own Python binding, unmodified:
det2trt/models/functions/grid_sampler.py (pure torch: `grid_sampler(input, grid, mode, pad, align)` ->
aten.grid_sampler(input, grid.permute(0,2,3,1)/10, …), :7-36, :144-223). Build container only.

Inputs follow the reference op test (det2trt/models/utils/test_trt_ops/test_grid_sampler.py:23-36): input ~ N(0,1),
grid spanning 1.5x the image on every side ([-15, 15]) plus noise, at reduced size so the fixtures stay small.
"""
import importlib.util
import itertools
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.helpers import make_grid_sampler_inputs  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_gs", f"{REF}/det2trt/models/functions/grid_sampler.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for interp, pad, align in itertools.product(["bilinear", "nearest", "bicubic"], ["zeros", "border", "reflection"],
                                                [False, True]):  # fmt: skip
        inp, grid = make_grid_sampler_inputs(2, 6, 11, 13, 17, 19, seed=3)
        out[f"{interp}_{pad}_{int(align)}"] = ref.grid_sampler(inp, grid, interp, pad, align).numpy()
    inp3, grid3 = make_grid_sampler_inputs(1, 3, 5, 6, 7, 8, seed=4, depth=(4, 5))
    for interp, pad, align in itertools.product(["bilinear", "nearest"], ["zeros", "border", "reflection"],
                                                [False, True]):  # fmt: skip
        out[f"3d_{interp}_{pad}_{int(align)}"] = ref.grid_sampler(inp3, grid3, interp, pad, align).numpy()
    np.savez_compressed(os.path.join(HERE, "grid_sampler_ref.npz"), **out)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
