#!/usr/bin/env python
"""Does the NUMA node of the allocating thread decide pinned-memory H2D bandwidth on this box? Prints one JSON line."""
import ctypes
import json
import os

import torch

rt = ctypes.CDLL("libcudart.so")
N = 256 << 20


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


def parse_cpulist(s):
    out = set()
    for part in s.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            out.update(range(int(a), int(b) + 1))
        elif part:
            out.add(int(part))
    return out


def main():
    torch.cuda.init()
    props = torch.cuda.get_device_properties(0)
    res = {"cpus": os.cpu_count()}
    bdf = None
    try:
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        res["bdf"] = bdf
        res["numa_node"] = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
        local = parse_cpulist(open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read())
        res["local_cpus"] = len(local)
    except Exception as e:  # noqa: BLE001
        res["sysfs_error"] = str(e)[:200]
        local = set()
    try:
        res["nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    except Exception as e:  # noqa: BLE001
        res["nodes_error"] = str(e)[:100]
    dev = torch.empty(N, dtype=torch.uint8, device="cuda")
    everything = os.sched_getaffinity(0)
    res["affinity_cpus"] = len(everything)
    remote = everything - local if local else set()

    def trial(tag):
        t = torch.empty(N, dtype=torch.uint8).pin_memory()
        res[f"torch_pin_{tag}_GBps"] = timed(lambda: dev.copy_(t, non_blocking=True))
        p = ctypes.c_void_p()
        assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(N), ctypes.c_uint(0)) == 0
        ctypes.memset(p, 1, N)
        cur = torch.cuda.current_stream().cuda_stream
        res[f"raw_hostalloc_{tag}_GBps"] = timed(lambda: rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), p, ctypes.c_size_t(N), 1, ctypes.c_void_p(cur)))
        rt.cudaFreeHost(p)
        del t

    trial("default_affinity")
    if local and (local & everything):
        os.sched_setaffinity(0, local & everything)
        trial("gpu_local_cpus")
    if remote:
        os.sched_setaffinity(0, remote)
        trial("remote_cpus")
    os.sched_setaffinity(0, everything)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
