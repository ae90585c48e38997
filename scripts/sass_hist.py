#!/usr/bin/env python
"""SASS instruction histogram of the library's kernels (cuobjdump -sass on the in-tree .so; runs on the build box).
    python scripts/sass_hist.py [regex of kernel names] > profiles/rNN_sass_histograms.txt
Per kernel: instruction count, the 14 most frequent mnemonics, and the Blackwell-specific ones the profiling recipe
names (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UBLKCP = TMA, HMMA = mma.sync, IDP = dp2a/dp4a)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bevformer_tensorrt_b200", "lib", "libb200_bev_ops.so")
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else "msda_gather_kernelI6__halfS1_Li32ELi2ELi0ELi0ELb0|msda_gather_kernelIaS|"
                 "msda_i8p_kernelI6__halfLb0|msda_pack|msda_res_kernelILb0|dcn_fused_kernelILi2ELb|dcn_generic|"
                 "grid_sample_2d_kernelILi1ELi0|peer_reduce|rotate_hwc_kernelILi1ELi0|point_sampling_kernelILi4")
SPECIAL = re.compile(r"^(UTC\w*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|HMMA|IDP|UTCBAR|SYNCS|FFMA2|FHFMA|REDG|RED)")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
name, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and name and pat.search(name):
        hist.setdefault(name, collections.Counter())[m.group(1)] += 1
for name, h in sorted(hist.items()):
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    print(f"== {dem[:150]}")
    print(f"   instructions: {sum(h.values())}")
    print("   top: " + ", ".join(f"{k} {v}" for k, v in h.most_common(14)))
    sp = collections.Counter()
    for k, v in h.items():
        m = SPECIAL.match(k)
        if m:
            sp[k.split('.')[0] if not k.startswith(('UTMALDG', 'LDTM', 'IDP')) else k] += v
    if sp:
        print("   blackwell / packed: " + ", ".join(f"{k} {v}" for k, v in sorted(sp.items())))
