#!/usr/bin/env python
"""Runs the encoder-prologue ops a few times at BEVFormer-base size (for ncu captures of rotate / point sampling)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevformer_tensorrt_b200 as bt  # noqa: E402
from bevformer_tensorrt_b200.workloads import camera_ring_lidar2img  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
prev = torch.randn(200 * 200, 1, 256, device="cuda", dtype=torch.float16)
ang, ctr = torch.tensor([2.3], device="cuda").half(), torch.tensor([100.0, 100.0], device="cuda").half()
l2i = camera_ring_lidar2img(6).cuda()
for _ in range(n):
    bt.rotate(prev.view(200, 200, 256).permute(2, 0, 1), ang, ctr, "bilinear")
    bt.rotate(prev.view(200, 200, 256).permute(2, 0, 1).contiguous(), ang, ctr, "bilinear")
    bt.bev_point_sampling(200, 200, (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), l2i, (928, 1600), 4, dtype=torch.float16)
torch.cuda.synchronize()
