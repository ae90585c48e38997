import sys, torch, numpy as np
sys.path.insert(0, '.')
import bevformer_tensorrt_b200 as bt
from tests.helpers import make_dcn_inputs
from oracle import dcn as odcn
case = sys.argv[1] if len(sys.argv) > 1 else "fused_co128"
x, off, mask, w, b, kw = make_dcn_inputs(case, dtype=torch.float16)
want = odcn.modulated_deformable_conv2d(*(t.float().numpy() for t in (x, off, mask, w, b)), **kw)
args = [t.cuda() for t in (x, off, mask, w)]
for rep in range(3):
    for bias in (b.cuda(), None):
        out = bt.modulated_deformable_conv2d(*args, bias, kw["stride"], kw["padding"], kw["dilation"], 1, 1)
        torch.cuda.synchronize()
        ref = want if bias is not None else want - b.float().numpy()[None, :, None, None]
        print(case, rep, "bias" if bias is not None else "nobias", "err", np.abs(out.float().cpu().numpy() - ref).max())
