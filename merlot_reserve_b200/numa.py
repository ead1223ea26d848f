"""Host-side placement for one-process-per-GPU runs: pin the process (and therefore its page-locked staging buffers,
which Linux allocates first-touch on the local node) to the CPUs that share a NUMA node / PCIe root with its GPU.
Eight ranks that each push ~0.8 GB of fp32 patches per step over PCIe otherwise meet on the memory controller of
whichever socket the launcher happened to start them on (the 4-5 % end-to-end loss at 8 GPUs seen in round 1)."""
from __future__ import annotations

import os
from typing import List, Optional


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_local_cpus(device_index: int) -> Optional[List[int]]:
    """CPUs local to CUDA device `device_index` (sysfs `local_cpulist` of its PCI function); None when unknown."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index)
        pci = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
    except Exception:
        return None
    path = f"/sys/bus/pci/devices/{pci}/local_cpulist"
    try:
        with open(path) as f:
            cpus = _parse_cpulist(f.read())
    except OSError:
        return None
    allowed = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else set(cpus)
    cpus = [c for c in cpus if c in allowed]
    return cpus or None


def bind_to_gpu(device_index: int) -> Optional[List[int]]:
    """Restricts this process to the GPU-local CPUs.  Returns the CPU list it bound to, or None (left unchanged)."""
    cpus = gpu_local_cpus(device_index)
    if not cpus or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        os.sched_setaffinity(0, cpus)
    except OSError:
        return None
    return cpus
