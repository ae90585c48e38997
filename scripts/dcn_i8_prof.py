"""Profiling target: the INT8 DCN plugin op at the R101 stage-3 layer (same tensors as bench.py's dcn_i8_base leg)."""
import sys, torch
sys.path.insert(0, '.')
import bevformer_tensorrt_b200 as bt
g = torch.Generator(device="cuda").manual_seed(0)
xq = torch.randint(-127, 127, (6, 64, 58, 100, 4), dtype=torch.int8, device="cuda")
wq = torch.randint(-127, 127, (256, 64, 3, 3, 4), dtype=torch.int8, device="cuda")
oq = torch.randint(-127, 127, (6, 18, 58, 100), dtype=torch.int8, device="cuda")
mq = torch.randint(0, 127, (6, 9, 58, 100), dtype=torch.int8, device="cuda")
b = torch.randn(256, device="cuda", generator=g).half()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    bt.modulated_deformable_conv2d_int8(xq, 0.02, oq, 0.03, mq, 1 / 127, wq, 0.001, b, 0.05, 256, 1, 1, 1, 1, 1)
torch.cuda.synchronize()
