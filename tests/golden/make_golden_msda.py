"""Generates tests/golden/msda_*.npz by running the REFERENCE's own Python code (run in the build container only;
/root/reference does not exist on the GPU box, the committed .npz files travel instead).

What runs, unmodified, from /root/reference:
  * det2trt/models/functions/multi_scale_deformable_attn.py  — `_MultiScaleDeformableAttnFunction.forward`
    (the plugin-signature adapter, :58-123): reshapes offsets, loc = ref + off/(W,H), softmax over L*P.
  * det2trt/models/utils/trt_ops.py — `multi_scale_deformable_attn_pytorch` (:4-85), the pure-PyTorch CPU op.
The binding normally calls mmcv's compiled `_ext.ms_deform_attn_forward` (absent here: mmcv-full 1.5.0 is not
installed and not vendored). This script installs a stand-in `mmcv.utils.ext_loader` whose `ms_deform_attn_forward`
forwards to the reference's own pure-PyTorch op, so every arithmetic step in the golden vectors is reference code.

Usage:  python tests/golden/make_golden_msda.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from bevformer_tensorrt_b200.workloads import CONFIGS, MSDAConfig, make_msda_inputs  # noqa: E402


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_binding():
    trt_ops = _load(f"{REF}/det2trt/models/utils/trt_ops.py", "ref_trt_ops")

    class _Ext:
        @staticmethod
        def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                                   im2col_step):  # fmt: skip
            bs, _, num_heads, ch = value.shape
            _, nq, _, nl, npts, _ = sampling_locations.shape
            sizes = [int(h) * int(w) for h, w in spatial_shapes.tolist()]
            levels = list(value.split(sizes, dim=1))
            return trt_ops.multi_scale_deformable_attn_pytorch(
                levels, spatial_shapes, sampling_locations, attention_weights, num_heads, num_heads * ch, nl, npts,
                nq, bs,
            )  # fmt: skip

        ms_deform_attn_backward = None

    mmcv = types.ModuleType("mmcv")
    mmcv_utils = types.ModuleType("mmcv.utils")
    ext_loader = types.ModuleType("mmcv.utils.ext_loader")
    ext_loader.load_ext = lambda name, funcs: _Ext
    mmcv_utils.ext_loader = ext_loader
    mmcv.utils = mmcv_utils
    sys.modules.update({"mmcv": mmcv, "mmcv.utils": mmcv_utils, "mmcv.utils.ext_loader": ext_loader})
    binding = _load(f"{REF}/det2trt/models/functions/multi_scale_deformable_attn.py", "ref_msda_binding")
    return binding._MultiScaleDeformableAttnFunction


def input_digest(*tensors):
    import hashlib

    h = hashlib.sha256()
    for t in tensors:
        h.update(t.contiguous().numpy().tobytes())
    return h.hexdigest()


CASES = [
    # (file stem, config, distribution, seed)
    ("msda_cpu_plumbing_U", CONFIGS["cpu_plumbing"], "U", 0),
    ("msda_small_sca_U", CONFIGS["small_sca"], "U", 1),
    ("msda_small_sca_edge", CONFIGS["small_sca"], "edge", 2),
    ("msda_tsa_like_U", MSDAConfig("tsa_like", 2, 257, 8, 32, ((20, 20),), 4, 1), "U", 3),
    ("msda_odd_U", MSDAConfig("odd", 1, 61, 3, 20, ((7, 9), (4, 5)), 3, 1), "U", 4),
]


def main():
    fn = load_reference_binding()
    torch.set_num_threads(4)
    for stem, cfg, dist, seed in CASES:
        value, shapes, ref, off, logits = make_msda_inputs(cfg, dist, seed, torch.float32)
        out = fn.apply(value, shapes.long(), ref, off, logits)  # int64 shapes, as transformer.py:313 passes them
        out = out.reshape(cfg.batch, cfg.num_query, cfg.num_heads, cfg.channels)
        # Inputs are reproducible from (config, dist, seed) through the seeded CPU generator, so only the expected
        # output and a digest of the inputs are stored (small fixtures); the self-contained "odd" case keeps its inputs.
        digest = input_digest(value, shapes, ref, off, logits)
        extra = {}
        if value.numel() < 50000:
            extra = dict(value=value.numpy(), ref=ref.numpy(), off=off.numpy(), logits=logits.numpy())
        np.savez_compressed(
            os.path.join(HERE, stem + ".npz"),
            shapes=shapes.numpy(), out=out.numpy(), dist=np.array(dist), digest=np.array(digest),
            meta=np.array([cfg.batch, cfg.num_query, cfg.num_heads, cfg.channels, cfg.num_levels, cfg.num_points,
                           cfg.points_per_group, seed]),
            **extra,
        )  # fmt: skip
        print(stem, tuple(out.shape), "absmax", float(out.abs().max()))


if __name__ == "__main__":
    main()
