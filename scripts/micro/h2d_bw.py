#!/usr/bin/env python
"""PCIe copy bandwidth on the box: H2D from default-pinned vs write-combined pinned host memory, alone and with a
concurrent D2H (what HostMSDA's pipeline does). Prints one JSON line."""
import ctypes
import json

import torch

rt = ctypes.CDLL("libcudart.so")
N = 256 << 20


def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags)) == 0
    return p


def timed(fn, n=8):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return N * n / (a.elapsed_time(b) * 1e-3) / 1e9


def main():
    torch.cuda.init()
    dev = torch.empty(N, dtype=torch.uint8, device="cuda")
    dev2 = torch.empty(N, dtype=torch.uint8, device="cuda")
    out_host = torch.empty(N, dtype=torch.uint8).pin_memory()
    s_out = torch.cuda.Stream()
    cur = torch.cuda.current_stream().cuda_stream
    res = {}
    for name, flags in (("pinned_default", 0), ("pinned_write_combined", 4), ("pinned_portable_mapped", 1 | 2)):
        p = host_alloc(N, flags)
        ctypes.memset(p, 1, N)

        def h2d():
            assert rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), p, ctypes.c_size_t(N), 1, ctypes.c_void_p(cur)) == 0

        res[name + "_h2d_GBps"] = timed(h2d)

        def both():
            with torch.cuda.stream(s_out):
                out_host.copy_(dev2, non_blocking=True)
            h2d()

        res[name + "_h2d_with_concurrent_d2h_GBps"] = timed(both)
        torch.cuda.synchronize()
        rt.cudaFreeHost(p)
    t = torch.empty(N, dtype=torch.uint8).pin_memory()
    res["torch_pin_memory_h2d_GBps"] = timed(lambda: dev.copy_(t, non_blocking=True))
    res["d2h_GBps"] = timed(lambda: out_host.copy_(dev2, non_blocking=True))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
