import sys, torch
sys.path.insert(0, '.')
import bevformer_tensorrt_b200 as bt
torch.manual_seed(0)
xd = torch.randn(6, 256, 58, 100, device="cuda").half()
off = (torch.randn(6, 18, 58, 100, device="cuda") * 2).half()
mask = torch.sigmoid(torch.randn(6, 9, 58, 100, device="cuda")).half()
w = (torch.randn(256, 256, 3, 3, device="cuda") / 48).half()
b = torch.randn(256, device="cuda").half()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    bt.modulated_deformable_conv2d(xd, off, mask, w, b, 1, 1, 1, 1, 1)
torch.cuda.synchronize()
