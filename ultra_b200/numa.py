"""NUMA placement of a rank next to its GPU (one process per GPU).

The e2e path moves ~1 GB per call between pinned host memory and the GPU.  On a two-socket box (B200 HGX: GPU0-3 hang
off NUMA node 0, GPU4-7 off node 1) a rank whose pinned buffers were first-touched on the far socket pays the
inter-socket link on every copy, and eight unpinned ranks also crowd one socket's memory controllers (round 1: e2e
scaling efficiency 0.46 at 8 GPUs).  `pin_to_gpu_numa` restricts the calling process to the CPUs of the GPU's node
BEFORE it allocates pinned memory, so first-touch places the pages locally.  Pure sysfs reads; no-op when the topology
cannot be determined (single-node boxes, containers without sysfs)."""
import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(text):
    cpus = set()
    for part in (text or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index):
    """NUMA node of a CUDA device from its PCI address (sysfs), or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        addr = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None
    node = _read("/sys/bus/pci/devices/%s/numa_node" % addr)
    try:
        node = int(node)
    except (TypeError, ValueError):
        return None
    return node if node >= 0 else None


def pin_to_gpu_numa(device_index):
    """Restrict this process to the CPUs of the GPU's NUMA node (intersected with the current affinity).
    Returns {"node": n, "cpus": count} or {"node": None, "why": ...}."""
    node = gpu_numa_node(device_index)
    if node is None:
        return {"node": None, "why": "no NUMA node for the device in sysfs"}
    cpus = _parse_cpulist(_read("/sys/devices/system/node/node%d/cpulist" % node))
    try:
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if not target:
            return {"node": node, "why": "node CPUs outside the allowed set", "cpus": 0}
        os.sched_setaffinity(0, target)
        return {"node": node, "cpus": len(target)}
    except OSError as exc:
        return {"node": node, "why": repr(exc), "cpus": 0}
