"""Keep a rank's host side next to its GPU.

The host-buffer path (`ggs_ctx_fitness_host`, `HostEvaluator`) streams the genomes of every
evaluation over PCIe from pinned host memory.  With one process per GPU on a two-socket box, a rank
that runs (and first-touches its pinned buffers) on the socket its GPU does NOT hang off pulls
every byte across the socket interconnect, and eight ranks doing so at once share that link.
`bind_to_device` pins the calling process to the CPUs of the GPU's NUMA node BEFORE the buffers
are allocated; it changes nothing on single-node machines and never raises."""
from __future__ import annotations

import os
from typing import Dict, List, Optional


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            cpus.extend(range(int(lo), int(hi) + 1))
        else:
            cpus.append(int(part))
    return cpus


def device_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of CUDA device `device_index` from sysfs, or None when the kernel does not say."""
    import torch

    p = torch.cuda.get_device_properties(device_index)
    bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    try:
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def bind_to_device(device_index: int) -> Dict[str, object]:
    """Restrict the calling process to the CPUs of the GPU's NUMA node (threads and pinned
    allocations made afterwards follow).  Returns what was done: {"node", "cpus", "bound"}."""
    info: Dict[str, object] = {"node": None, "cpus": 0, "bound": False}
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        node = device_numa_node(device_index)
        info["node"] = node
        if node is None or len(nodes) < 2:
            return info  # one node, or unknown: nothing to choose
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            local = set(_parse_cpulist(f.read()))
        allowed = os.sched_getaffinity(0) & local
        if not allowed:
            return info
        os.sched_setaffinity(0, allowed)
        info["cpus"] = len(allowed)
        info["bound"] = True
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    return info
