"""Host-side placement for the host-buffer (libenv) path: its limiter is the 12 KiB/env D2H stream,
and with one process per GPU that stream wants its destination pages (and the thread that drains
them) on the NUMA node the GPU hangs off. Opt-in helper; nothing in the library calls it on its own."""
from __future__ import annotations

import os


def _parse_cpulist(txt):
    cpus = set()
    for part in txt.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index: int):
    """NUMA node of a CUDA device from sysfs (None when the platform does not say)."""
    try:
        import torch

        prop = torch.cuda.get_device_properties(device_index)
        bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        return node if node >= 0 else None
    except Exception:
        return None


def pin_to_gpu_numa_node(device_index: int):
    """Restrict this process to the CPUs of the GPU's NUMA node (so first-touch and cudaHostAlloc
    place host buffers there). Returns a dict describing what was done."""
    node = gpu_numa_node(device_index)
    info = {"device": device_index, "numa_node": node, "pinned": False}
    if node is None:
        return info
    try:
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(pinned=True, cpus=len(allowed))
    except Exception as e:  # noqa: BLE001
        info["error"] = repr(e)
    return info
