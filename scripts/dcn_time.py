#!/usr/bin/env python
"""Device time of the DCNv2 plugin ops at the R101 stage-3 layer [6,256,58,100] 3x3 (same tensors as bench.py's
dcn_*_base legs): FP16 and INT8, CUDA-graph replay of 40 calls timed with events. Prints one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevformer_tensorrt_b200 as bt  # noqa: E402


def time_graph(fn, n=40, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n * 1e3)
    return best


g = torch.Generator(device="cuda").manual_seed(0)
xd = torch.randn(6, 256, 58, 100, device="cuda", generator=g)
off = torch.randn(6, 18, 58, 100, device="cuda", generator=g) * 2
mask = torch.sigmoid(torch.randn(6, 9, 58, 100, device="cuda", generator=g))
w = torch.nn.Parameter(torch.randn(256, 256, 3, 3, device="cuda", generator=g).half() / 48, requires_grad=False)
b = torch.randn(256, device="cuda", generator=g).half()
flops = 2 * 256 * 2304 * 5800 * 6
a16 = [xd.half(), off.half(), mask.half()]
xq = torch.randint(-127, 127, (6, 64, 58, 100, 4), dtype=torch.int8, device="cuda")
wq = torch.randint(-127, 127, (256, 64, 3, 3, 4), dtype=torch.int8, device="cuda")
oq = torch.randint(-127, 127, (6, 18, 58, 100), dtype=torch.int8, device="cuda")
mq = torch.randint(0, 127, (6, 9, 58, 100), dtype=torch.int8, device="cuda")
out = {}
for name, fn in (("dcn_f16_base", lambda: bt.modulated_deformable_conv2d(*a16, w, b, 1, 1, 1, 1, 1)),
                 ("dcn_i8_base", lambda: bt.modulated_deformable_conv2d_int8(xq, 0.02, oq, 0.03, mq, 1 / 127, wq, 0.001, b,
                                                                             0.05, 256, 1, 1, 1, 1, 1))):
    us = time_graph(fn)
    out[name] = {"kernel_us": us, "tflops": flops / (us * 1e-6) / 1e12}
print(json.dumps(out))
