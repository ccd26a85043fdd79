#!/usr/bin/env python3
"""Where does pinned host memory land relative to the GPU?  h2d / d2h bandwidth of pinned buffers
allocated while the process is bound to each NUMA node's CPUs."""
import glob
import os
import time

import torch


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def bw(h, d, iters=6):
    res = {}
    for nm, src, dst in (("h2d", h, d), ("d2h", d, h)):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        res[nm] = round(iters * h.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
    return res


def main():
    torch.cuda.init()
    bus = torch.cuda.get_device_properties(0)
    pci = "%04x:%02x:%02x.0" % (bus.pci_domain_id, bus.pci_bus_id, bus.pci_device_id) if hasattr(bus, "pci_bus_id") else None
    print("gpu pci", pci, "cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    if pci:
        for f in ("numa_node", "local_cpulist"):
            try:
                print(" ", f, open("/sys/bus/pci/devices/%s/%s" % (pci, f)).read().strip())
            except OSError as e:
                print(" ", f, "unreadable", e)
    nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
    print("numa nodes:", [os.path.basename(n) for n in nodes])
    d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    allowed = os.sched_getaffinity(0)
    h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    print("default placement (cpu %s)" % (os.sched_getcpu() if hasattr(os, "sched_getcpu") else "?"), bw(h, d))
    del h
    for n in nodes:
        cpus = set(cpulist(open(n + "/cpulist").read())) & allowed
        if not cpus:
            print(os.path.basename(n), "no allowed cpus")
            continue
        os.sched_setaffinity(0, cpus)
        time.sleep(0.01)
        h = torch.empty(64 << 20, dtype=torch.uint8)
        h.fill_(1)
        h = h.pin_memory()
        print(os.path.basename(n), "cpus", len(cpus), bw(h, d))
        del h
        os.sched_setaffinity(0, allowed)


if __name__ == "__main__":
    main()
