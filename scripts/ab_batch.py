#!/usr/bin/env python
"""A/B of the launch shapes of the MSDA plugin op at BEVFormer-base shapes (csrc/msda.cu UPW, csrc/msda_v2.cu UPW):
1 unit per warp (round-1 grid) against the batched launches with the visibility scan (2 / 4 units, neighbouring or
grid-strided). CUDA events per call, mean / min of N after warm-up, inputs resident in HBM (590 MB FP16 per call: larger
than L2). Every shape's output is compared bit for bit with the 1-unit launch. Results are appended to the JSON file as
they come, so a cut-off visit still leaves what was measured."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevformer_tensorrt_b200 as bt  # noqa: E402
from bevformer_tensorrt_b200.workloads import CONFIGS, make_msda_inputs, quantize_per_tensor  # noqa: E402

OUT = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ab_batch.json"
N = int(os.environ.get("AB_N", "40"))
res = {}


def dump():
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1)


def timeit(fn, n=N, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    per = sorted(a.elapsed_time(b) for a, b in evs)
    return {"ms": sum(per) / len(per), "min_ms": per[0], "med_ms": per[len(per) // 2]}


SHAPES = [(1, False), (2, False), (4, False), (2, True), (4, True)]


def tag(u, s):
    return f"{u}{'s' if s else ''}"


def main():
    cfg = CONFIGS["base_sca"]
    for dist in ("G", "U"):
        v, sh, r, o, w = make_msda_inputs(cfg, dist, 0, torch.float32)
        f16 = [v.half().cuda(), sh.cuda(), r.half().cuda(), o.half().cuda(), w.half().cuda()]
        vq, sv = quantize_per_tensor(v)
        oq, so = quantize_per_tensor(o)
        wq, sw = quantize_per_tensor(w)
        i8 = [vq.cuda(), sh.cuda(), r.half().cuda(), oq.cuda(), wq.cuda()]
        sout = 1.6 / 127
        legs = [("f16", lambda: bt.multi_scale_deformable_attn(*f16), SHAPES),
                ("i8", lambda: bt.multi_scale_deformable_attn_int8(i8[0], sv, i8[1], i8[2], i8[3], so, i8[4], sw, sout),
                 [(1, False), (2, False), (4, False)])]
        if os.environ.get("AB_F32", "1") == "1" and dist == "G":
            f32 = [v.cuda(), sh.cuda(), r.cuda(), o.cuda(), w.cuda()]
            legs.append(("f32", lambda: bt.multi_scale_deformable_attn(*f32), [(1, False), (4, False), (4, True)]))
        for name, fn, shapes in legs:
            base = None
            for u, s in shapes:
                key = f"{name}_{dist}_{tag(u, s)}"
                try:
                    bt.set_msda_batch_units(u, s)
                    got = fn()
                    torch.cuda.synchronize()
                    if base is None:
                        base = got
                    res[key] = dict(timeit(fn), identical_to_1=bool(torch.equal(got, base)))
                except Exception as e:  # keep going: the other legs are still worth having
                    res[key] = {"error": repr(e)}
                dump()
                print(key, res[key], flush=True)
            bt.set_msda_batch_units(1)
        del f16, i8
        torch.cuda.empty_cache()
    dump()


if __name__ == "__main__":
    main()
