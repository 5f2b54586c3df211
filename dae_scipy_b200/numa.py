"""Host-side placement for multi-GPU runs: keep a rank's threads and page-locked buffers on the NUMA node of its GPU.

With one process per GPU, every rank copies its results (1 GB per million pendulum systems and step) into page-locked host
memory.  The pages of a `cudaHostAlloc` allocation are placed on the node of the allocating thread, so eight ranks started
without affinity all write through one memory controller (round 1 measured 71 GB/s aggregate at N = 8, SCALE_r01.json).
`bind_to_device_node()` reads the GPU's PCI address from the CUDA runtime, its NUMA node and that node's CPU list from sysfs, and
pins the calling process there.  Linux only; a no-op (returning None) wherever the information is not available.
"""
from __future__ import annotations

import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def device_numa_node(device_index: int):
    """NUMA node of CUDA device `device_index` (as numbered by the runtime), or None."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_device_node(device_index: int):
    """Restrict the calling process to the CPUs of the GPU's NUMA node.  Returns {"node", "cpus"} or None."""
    node = device_numa_node(device_index)
    if node is None:
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed)}
    except Exception:
        return None
