"""Debug aid: INT8 v2 (resident) vs round-1 kernel on small shapes; prints where they differ."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bevformer_tensorrt_b200 as bt
from bevformer_tensorrt_b200 import _lib
from bevformer_tensorrt_b200.workloads import MSDAConfig, make_msda_inputs, quantize_per_tensor

lib = _lib.load()
for name, cfg, cap in [("one_level", MSDAConfig("a", 1, 8, 1, 32, ((6, 6),), 16, 1), 131072),
                       ("two_level", MSDAConfig("b", 1, 8, 2, 32, ((6, 6), (3, 3)), 8, 1), 131072),
                       ("two_level_smallcap", MSDAConfig("b", 1, 8, 2, 32, ((6, 6), (3, 3)), 8, 1), 4096),
                       ("small_sca", MSDAConfig("small_sca", 2, 333, 8, 32, ((12, 20), (6, 10), (3, 5), (2, 3)), 8, 4), 131072)]:
    lib.b200_msda_set_i8_resident_bytes(cap)
    v, sh, r, o, w = make_msda_inputs(cfg, "U", 3, torch.float32)
    vq, sv = quantize_per_tensor(v); oq, so = quantize_per_tensor(o); wq, sw = quantize_per_tensor(w)
    args = (vq.cuda(), sv, sh.cuda(), r.half().cuda(), oq.cuda(), so, wq.cuda(), sw, 1.6 / 127)
    bt.set_msda_v2(False); a = bt.multi_scale_deformable_attn_int8(*args).cpu().numpy().astype(np.int32)
    bt.set_msda_v2(True); b = bt.multi_scale_deformable_attn_int8(*args).cpu().numpy().astype(np.int32)
    d = np.abs(a - b)
    print(name, "max diff", d.max(), "frac != ", (d > 1).mean(), "shape", a.shape)
    if d.max() > 1:
        bad = np.argwhere(d > 1)
        print("  first bad idx", bad[:6].tolist())
        i = tuple(bad[0][:3])
        print("  v1", a[i].tolist()); print("  v2", b[i].tolist())
        print("  bad items per (b,m):", [(bb, mm, int((d[bb, :, mm] > 1).any(-1).sum())) for bb in range(a.shape[0]) for mm in range(a.shape[2])][:16])
        print("  bad channels hist", (d > 1).reshape(-1, 32).sum(0).tolist())
