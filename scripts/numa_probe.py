#!/usr/bin/env python
"""Host->device copy rate of a pinned buffer by the NUMA node it was allocated on (which socket does the GPU hang off?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from edlib_b200 import workloads
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
print("nodes", nodes, "affinity", len(os.sched_getaffinity(0)))
d = torch.empty(150_000_000, dtype=torch.uint8, device=dev)
for node in [-1] + nodes:
    a = workloads.pinned_empty((150_000_000,), numa_node=node)
    t = torch.from_numpy(a)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d.copy_(t, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("pinned buffer allocated on node %2d: %.2f ms  %.1f GB/s" % (node, 1000 * dt, 0.15 / dt))
