#!/usr/bin/env python
"""A/B of the round-1 gather kernel (msda.cu) against the second-generation path (msda_v2.cu: pack + tensor-core / dp2a
gather) at BEVFormer-base shapes: CUDA events per call, mean of 30 after 5 warm-up, inputs resident in HBM."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevformer_tensorrt_b200 as bt  # noqa: E402
from bevformer_tensorrt_b200 import _lib  # noqa: E402
from bevformer_tensorrt_b200.workloads import CONFIGS, make_msda_inputs, quantize_per_tensor  # noqa: E402


def timeit(fn, n=30, warm=5):
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
    return sum(per) / len(per), per[0]


CAPS = [int(x) << 10 for x in os.environ.get("AB_CAPS", "128,32,8").split(",")]


def main():
    cfg = CONFIGS["base_sca"]
    _lib.load()
    out = {}
    for dist in ("U", "G"):
        host = make_msda_inputs(cfg, dist, 0, torch.float32)
        v, sh, r, o, w = host
        f16 = [v.half().cuda(), sh.cuda(), r.half().cuda(), o.half().cuda(), w.half().cuda()]
        vq, sv = quantize_per_tensor(v)
        oq, so = quantize_per_tensor(o)
        wq, sw = quantize_per_tensor(w)
        i8 = [vq.cuda(), sh.cuda(), r.half().cuda(), oq.cuda(), wq.cuda()]
        sout = 1.6 / 127

        def run16():
            return bt.multi_scale_deformable_attn(*f16)

        def run8():
            return bt.multi_scale_deformable_attn_int8(i8[0], sv, i8[1], i8[2], i8[3], so, i8[4], sw, sout)

        for tag, fn in (("f16", run16), ("i8", run8)):
            bt.set_msda_v2(False)
            bt.set_msda_f16_path(False)
            ref_out = fn().float()
            m, mn = timeit(fn)
            out[f"{tag}_{dist}_v1"] = {"ms": m, "min_ms": mn}
            bt.set_msda_v2(True)
            if tag == "i8":
                got = fn().float()
                m, mn = timeit(fn)
                out[f"{tag}_{dist}_v2"] = {"ms": m, "min_ms": mn, "max_abs_vs_v1": (got - ref_out).abs().max().item()}
            else:  # resident-tail kernel at several shared-memory capacities
                for cap in (CAPS if dist == "U" else CAPS[:1]):
                    bt.set_msda_f16_path(True, cap)
                    got = fn().float()
                    m, mn = timeit(fn)
                    out[f"{tag}_{dist}_res{cap >> 10}k"] = {"ms": m, "min_ms": mn,
                                                            "max_abs_vs_v1": (got - ref_out).abs().max().item()}
                bt.set_msda_f16_path(True, 128 << 10)
        del f16, i8
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
