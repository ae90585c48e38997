#!/usr/bin/env python
"""Runs the MSDA plugin op at base shapes a few times (profiling target for ncu):
    python scripts/prof_msda.py <f16|i8> <U|G> [iters] [v2:0|1]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevformer_tensorrt_b200 as bt  # noqa: E402
from bevformer_tensorrt_b200 import _lib  # noqa: E402
from bevformer_tensorrt_b200.workloads import CONFIGS, make_msda_inputs, quantize_per_tensor  # noqa: E402

dtype, dist = sys.argv[1], sys.argv[2]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
bt.set_msda_v2(bool(int(sys.argv[4])) if len(sys.argv) > 4 else True)
cfg = CONFIGS["base_sca"]
v, sh, r, o, w = make_msda_inputs(cfg, dist, 0, torch.float32)
if dtype == "f16":
    a = [v.half().cuda(), sh.cuda(), r.half().cuda(), o.half().cuda(), w.half().cuda()]
    fn = lambda: bt.multi_scale_deformable_attn(*a)  # noqa: E731
else:
    vq, sv = quantize_per_tensor(v)
    oq, so = quantize_per_tensor(o)
    wq, sw = quantize_per_tensor(w)
    a = [vq.cuda(), sh.cuda(), r.half().cuda(), oq.cuda(), wq.cuda()]
    fn = lambda: bt.multi_scale_deformable_attn_int8(a[0], sv, a[1], a[2], a[3], so, a[4], sw, 1.6 / 127)  # noqa: E731
for _ in range(iters):
    fn()
torch.cuda.synchronize()
