"""Host-side placement for the one-process-per-GPU layout: bind the process to the CPUs (and thereby, by first touch, its pinned
staging buffers to the memory) of the NUMA node its GPU hangs off.  On an 8-GPU HGX box the end-to-end path moves ≈ 0.5 GB per
GPU and step across PCIe; with buffers on the far socket that traffic crosses the inter-socket link and the eight ranks share it."""
from __future__ import annotations

import os


def _cpus_from_sysfs(pci_bus_id: str) -> list[int]:
    node_path = f"/sys/bus/pci/devices/{pci_bus_id.lower()}/numa_node"
    with open(node_path) as fh:
        node = int(fh.read().strip())
    if node < 0:
        return []
    with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
        spec = fh.read().strip()
    cpus = []
    for part in spec.split(","):
        a, _, b = part.partition("-")
        cpus += list(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa(device_index: int) -> list[int] | None:
    """Restrict this process to the CPUs local to CUDA device `device_index` (sched_setaffinity).  Call it before allocating
    pinned host buffers.  Returns the CPU list, or None when the topology is not exposed (containers without sysfs NUMA data) —
    placement is an optimisation, never an error."""
    if not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        cpus = _cpus_from_sysfs(bdf)
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus or len(cpus) == len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None
