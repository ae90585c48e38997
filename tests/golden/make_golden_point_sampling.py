"""Generates tests/golden/point_sampling_ref.npz by executing the REFERENCE's own code, unmodified:
BEVFormerEncoderTRTP.get_reference_points_3d and .point_sampling_trt (det2trt/models/modules/encoder.py:168-259).
The class lives in a module that imports mmcv/mmdet (absent here), so the two method definitions are read out of the
reference file with ``ast`` at generation time and executed as plain functions — nothing is copied into this
repository. Build container only.

Inputs: the synthetic 6-camera ring of bevformer_tensorrt_b200.workloads (NuScenes-like intrinsics, 928x1600 images),
BEVFormer's pc_range, reduced BEV sizes plus one non-square one.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
os.environ.setdefault("B200_BEV_OPS_BUILDING", "1")  # workloads.py is pure torch; do not require the .so here
from tests.helpers import POINT_SAMPLING_CASES, make_point_sampling_inputs  # noqa: E402


def reference_methods():
    src = open(f"{REF}/det2trt/models/modules/encoder.py").read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "BEVFormerEncoderTRTP")
    fns = {}
    for n in cls.body:
        if isinstance(n, ast.FunctionDef) and n.name in ("get_reference_points_3d", "point_sampling_trt"):
            n.decorator_list = []
            ns = {"torch": torch, "np": np}
            exec(compile(ast.Module(body=[n], type_ignores=[]), f"{REF}/det2trt/models/modules/encoder.py", "exec"), ns)
            fns[n.name] = ns[n.name]
    return fns


def main():
    fns = reference_methods()
    out = {}
    for case, (H, W, D, img_hw, pc_range) in POINT_SAMPLING_CASES.items():
        lidar2img = make_point_sampling_inputs(case)
        ref_3d = fns["get_reference_points_3d"](H, W, pc_range[5] - pc_range[2], D, bs=1, device="cpu",
                                                dtype=torch.float)  # fmt: skip
        me = types.SimpleNamespace(num_points_in_pillar=D)
        cam, mask = fns["point_sampling_trt"](me, ref_3d, list(pc_range), lidar2img, list(img_hw))
        out[f"{case}_ref3d"] = ref_3d.numpy()
        out[f"{case}_cam"] = cam.contiguous().numpy()
        out[f"{case}_mask"] = mask.numpy()
    np.savez_compressed(os.path.join(HERE, "point_sampling_ref.npz"), **out)
    print("wrote", len(out), "arrays", {k: v.shape for k, v in out.items() if k.startswith("ring_small")})


if __name__ == "__main__":
    main()
