#!/usr/bin/env python
"""Host <-> device copy bandwidth from pinned memory, CPU affinity and NUMA placement (context for the e2e leg)."""
import glob
import os
import time

import torch

print("cpus:", os.cpu_count(), "affinity:", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)))
for p in glob.glob("/sys/bus/pci/devices/*/numa_node"):
    try:
        cls = open(os.path.dirname(p) + "/class").read().strip()
        if cls.startswith("0x0302") or cls.startswith("0x0300"):
            print("gpu pci", os.path.basename(os.path.dirname(p)), "numa", open(p).read().strip())
    except OSError:
        pass
try:
    print("numa nodes:", [os.path.basename(x) for x in glob.glob("/sys/devices/system/node/node*")])
except OSError:
    pass
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, a, b in (("H2D", d, h), ("D2H", h, d)):
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.copy_(b, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("%s pinned 1 GiB: %.1f GB/s" % (name, n / dt / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(s1):
    d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2):
    h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("H2D + D2H concurrently, 1 GiB each: %.1f GB/s per direction" % (n / dt / 1e9))
