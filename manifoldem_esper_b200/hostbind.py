"""Keep a rank's host threads and page-locked buffers on the NUMA node of its GPU.

With one process per GPU the 500 MB image stack of every PD crosses PCIe from pinned host memory; if that memory sits on
the other socket the copy also crosses the inter-socket link, which all ranks then share.  Linux allocates pages on
the node of the CPU that first touches them, so restricting the process to the CPUs of the GPU's node before any
buffer is allocated is enough (no libnuma needed).  Every step is best effort: on a single-node host, inside a cpuset
that excludes those CPUs, or without sysfs access the process is left as it is.
"""
from __future__ import annotations

import os


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index: int):
    """NUMA node of a CUDA device from sysfs, or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        addr = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open('/sys/bus/pci/devices/%s/numa_node' % addr) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node.  Returns what was done (for logs / bench output)."""
    info = {'node': None, 'cpus': None, 'bound': False}
    try:
        node = gpu_numa_node(device_index)
        info['node'] = node
        if node is None or not hasattr(os, 'sched_setaffinity'):
            return info
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            node_cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        target = node_cpus & allowed
        if len(target) < 4 or target == allowed:        # nothing to gain, or too few CPUs for the host threads
            info['cpus'] = len(target)
            return info
        os.sched_setaffinity(0, target)
        info.update(cpus=len(target), bound=True)
    except Exception as e:                               # noqa: BLE001 -- best effort by design
        info['error'] = '%s: %s' % (type(e).__name__, e)
    return info
