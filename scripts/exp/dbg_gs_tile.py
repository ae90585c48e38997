"""Debug / A-B aid: grid-sampler tile modes (0 generic, 1 bulk rows, 2 tensor 2-D) per dtype in separate processes
(CUDA errors are sticky), with timings at the prev-BEV warp shape."""
import subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CODE = r'''
import sys, torch
sys.path.insert(0, %r)
import bevformer_tensorrt_b200 as bt
from bevformer_tensorrt_b200.functions.grid_sampler import pack_chw
sys.path.insert(0, %r + "/tests")
from helpers import make_rotation_grid
kind, mode, shape = sys.argv[1], int(sys.argv[2]), sys.argv[3]
if shape == "small":
    N, C, Hi, Wi, Ho, Wo = 1, 64, 32, 64, 24, 64
    grid = make_rotation_grid(Ho, Wo, 5.0)
elif shape == "inner":  # windows never reach the image border
    N, C, Hi, Wi, Ho, Wo = 1, 64, 64, 128, 16, 32
    grid = make_rotation_grid(64, 128, 3.0)[:, :, 24:40, 48:80].contiguous()
else:
    N, C, Hi, Wi, Ho, Wo = 1, 256, 200, 200, 200, 200
    grid = make_rotation_grid(Ho, Wo, 2.9, shift=(1.5, -0.7))
g = torch.Generator().manual_seed(1)
inp = torch.randn(N, C, Hi, Wi, generator=g)
lib = bt._lib.load()
if kind == "f32": a = (inp.cuda(), grid.cuda()); run = lambda: bt.grid_sampler(a[0], a[1], "bilinear", "zeros", False)
elif kind == "f16": a = (inp.half().cuda(), grid.half().cuda()); run = lambda: bt.grid_sampler(a[0], a[1], "bilinear", "zeros", False)
elif kind == "chw2":
    a = (pack_chw(inp.half(), 2).cuda(), grid.half().permute(0, 2, 3, 1).unsqueeze(1).contiguous().cuda())
    run = lambda: bt.grid_sampler_chw2(a[0], a[1], C, "bilinear", "zeros", False)
else:
    xi = torch.randint(-127, 127, (N, C // 4, Hi, Wi, 4), dtype=torch.int8).cuda()
    gi = torch.zeros(N, 1, Ho, Wo, 4, dtype=torch.int8)
    gi[..., 0] = (grid[:, 0] * 12.7).round().clamp(-127, 127).to(torch.int8)[:, None][:, 0]
    gi[..., 1] = (grid[:, 1] * 12.7).round().clamp(-127, 127).to(torch.int8)[:, None][:, 0]
    a = (xi, gi.cuda()); run = lambda: bt.grid_sampler_int8(a[0], 0.03, a[1], 10 / 127, 0.03, C, "bilinear", "zeros", False)
lib.b200_grid_sample_set_tile_path(0); ref = run(); torch.cuda.synchronize()
lib.b200_grid_sample_set_tile_path(mode); out = run(); torch.cuda.synchronize()
for _ in range(5): run()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(50): run()
gr.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
print(kind, "mode", mode, shape, "equal:", torch.equal(ref, out), "us %%.2f" %% (e0.elapsed_time(e1) / 50 * 1e3))
''' % (ROOT, ROOT)
for shape, kinds in (("inner", ("f32",)), ("small", ("f32", "f16")), ("base", ("f32", "f16", "chw2", "i8"))):
    for kind in kinds:
        for mode in ((2,) if shape != "base" else (0, 1, 2)):
            r = subprocess.run([sys.executable, "-c", CODE, kind, str(mode), shape], capture_output=True, text=True)
            print(shape, kind, mode, "rc", r.returncode, (r.stdout.strip() or r.stderr.strip()[-200:]), flush=True)
            if r.returncode != 0 and shape != "base":
                break
