"""Shared test helpers (tests may use oracle/; the product package may not)."""
import glob
import hashlib
import os

import numpy as np
import torch

from bevformer_tensorrt_b200.workloads import MSDAConfig, make_msda_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def input_digest(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.contiguous().numpy().tobytes())
    return h.hexdigest()


def golden_msda_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "msda_*.npz")))


def load_golden_msda(path):
    """Returns (cfg, (value, shapes, ref, off, logits) float32 CPU tensors, expected out ndarray)."""
    z = np.load(path)
    B, Q, M, C, L, P, G, seed = (int(x) for x in z["meta"])
    shapes = tuple((int(h), int(w)) for h, w in z["shapes"])
    cfg = MSDAConfig(os.path.basename(path), B, Q, M, C, shapes, P, G)
    inputs = make_msda_inputs(cfg, str(z["dist"]), seed, torch.float32)
    if input_digest(*inputs) != str(z["digest"]):
        if "value" in z.files:  # self-contained fixture
            inputs = (torch.from_numpy(z["value"]), inputs[1], torch.from_numpy(z["ref"]),
                      torch.from_numpy(z["off"]), torch.from_numpy(z["logits"]))  # fmt: skip
        else:
            raise RuntimeError(f"{path}: seeded CPU generator no longer reproduces the golden inputs")
    return cfg, inputs, z["out"]


def make_grid_sampler_inputs(N, C, Hi, Wi, Ho, Wo, seed=0, depth=None, span=15.0):
    """input ~ N(0,1); grid = a regular sweep over [-span, span] (1.5x beyond the image, as the reference test's
    linspace(-15, 15), test_grid_sampler.py:29-36) plus N(0, 0.5) jitter, channel-first [N, 2|3, ...]."""
    g = torch.Generator().manual_seed(seed)
    if depth is None:
        inp = torch.randn(N, C, Hi, Wi, generator=g)
        ys, xs = torch.meshgrid(torch.linspace(-span, span, Ho), torch.linspace(-span, span, Wo), indexing="ij")
        grid = torch.stack([xs, ys], 0)[None].repeat(N, 1, 1, 1)
    else:
        Di, Do = depth
        inp = torch.randn(N, C, Di, Hi, Wi, generator=g)
        zs, ys, xs = torch.meshgrid(torch.linspace(-span, span, Do), torch.linspace(-span, span, Ho),
                                    torch.linspace(-span, span, Wo), indexing="ij")  # fmt: skip
        grid = torch.stack([xs, ys, zs], 0)[None].repeat(N, 1, 1, 1, 1)
    grid = grid + 0.5 * torch.randn(grid.shape, generator=g)
    return inp, grid.contiguous()
