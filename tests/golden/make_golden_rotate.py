"""Generates tests/golden/rotate_ref.npz by running the REFERENCE's own Python binding, unmodified:
det2trt/models/functions/rotate.py (pure torch: theta -> affine grid by bmm -> aten.grid_sampler, :12-84, :99-124).
Build container only.

Inputs follow the reference op test (det2trt/models/utils/test_trt_ops/test_rotate.py:19-26): img ~ N(0,1), one angle
in degrees, a centre in pixels — at reduced size so the fixture stays small (tests/helpers.py ROTATE_CASES).
"""
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.helpers import ROTATE_CASES, make_rotate_inputs  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_rotate", f"{REF}/det2trt/models/functions/rotate.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for case in ROTATE_CASES:
        img, angle, center = make_rotate_inputs(case)
        for interp in ("bilinear", "nearest"):
            # the binding is called as the model calls it (transformer.py:298-302): angle a 0-d tensor, center [2]
            out[f"{case}_{interp}"] = ref.rotate(img, angle[0], center, interpolation=interp).numpy()
    np.savez_compressed(os.path.join(HERE, "rotate_ref.npz"), **out)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
