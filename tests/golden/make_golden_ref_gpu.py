#!/usr/bin/env python
"""Golden vectors from the reference's OWN CUDA kernels (run on a B200 box through gpurun; /root/reference is not needed at
run time: oracle/_ref/libref_kernels.so was compiled from the reference's .cu files where they lie and travels with the
snapshot).

    python tests/golden/make_golden_ref_gpu.py  [out_dir]        # default: gpurun_out/golden_ref_gpu

For each small seeded case it runs ms_deformable_im2col_cuda<float> (…Kernel.cu:1106-1128) and
ms_deformable_im2col_cuda_int8<float> / <__half2> (:1172-1218) and stores inputs' digest + outputs. The files are then
copied to tests/golden/ref_gpu_*.npz and checked on the CPU every round by tests/test_oracle_golden.py against
oracle/msda_oracle.c (FP32 restatement, INT8 quantised-intermediate emulation)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import input_digest  # noqa: E402
from bevformer_tensorrt_b200.workloads import CONFIGS, MSDAConfig, make_msda_inputs, quantize_per_tensor  # noqa: E402
from oracle import msda as omsda  # noqa: E402

CASES = [
    ("small_sca", CONFIGS["small_sca"], "U", 201),
    ("small_sca", CONFIGS["small_sca"], "edge", 202),
    ("tsa_like", MSDAConfig("tsa_like", 2, 300, 8, 32, ((30, 30),), 4, 1), "edge", 203),
    ("g2", MSDAConfig("g2", 2, 129, 8, 32, ((9, 11), (5, 6)), 8, 2), "U", 204),
]


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    rk = omsda.RefKernels()
    for name, cfg, dist, seed in CASES:
        value, shapes, ref, off, logits = make_msda_inputs(cfg, dist, seed, torch.float32)
        dev = [t.cuda() for t in (value, shapes, ref, off, logits)]
        out32 = rk.msda(*dev, variant="f32").cpu().numpy()
        vq, sv = quantize_per_tensor(value)
        oq, so = quantize_per_tensor(off)
        wq, sw = quantize_per_tensor(logits)
        sout = float(np.abs(out32).max()) / 127.0
        i8 = {}
        for rdt, tag in ((torch.float32, "f32ref"), (torch.float16, "h2ref")):
            i8[tag] = rk.msda_i8(vq.cuda(), sv, dev[1], ref.to(rdt).cuda(), oq.cuda(), so, wq.cuda(), sw, sout).cpu().numpy()
        torch.cuda.synchronize()
        meta = np.array([cfg.batch, cfg.num_query, cfg.num_heads, cfg.channels, cfg.num_levels, cfg.num_points,
                         cfg.points_per_group, seed], np.int64)  # fmt: skip
        np.savez_compressed(os.path.join(out_dir, f"ref_gpu_{name}_{dist}.npz"), meta=meta,
                            shapes=np.array(cfg.spatial_shapes, np.int32), dist=dist,
                            digest=input_digest(value, shapes, ref, off, logits), out_f32=out32,
                            scales=np.array([sv, so, sw, sout], np.float64), out_i8_f32ref=i8["f32ref"],
                            out_i8_h2ref=i8["h2ref"], device=torch.cuda.get_device_name(0))
        print(name, dist, "f32 amax", float(np.abs(out32).max()), "i8 agreement f32ref/h2ref",
              float((i8["f32ref"] == i8["h2ref"]).mean()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden_ref_gpu"))
