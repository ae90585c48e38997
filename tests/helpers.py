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
