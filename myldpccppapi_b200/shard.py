"""Batch sharding of independent codewords over the GPUs of one box (no collective).

The reference decodes a stream codeword by codeword (MyLdpc.cpp:694) / chunk by chunk
(MyLdpc.cpp:577-616); codewords never interact, so GPU g simply decodes the contiguous range
shard_range(ncw, g, G) and its bytes land in the matching slice of the caller's buffer.
"""
from __future__ import annotations

from typing import List, Tuple


def shard_range(ncw: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous [begin, end) of codewords for `rank`.  `align` keeps shard boundaries on a
    multiple of `align` codewords (use 8 when K is not a multiple of 8 so bytes stay whole)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    units = (ncw + align - 1) // align
    b = units * rank // world * align
    e = units * (rank + 1) // world * align
    return min(b, ncw), min(e, ncw)


def shard_ranges(ncw: int, world: int, align: int = 1) -> List[Tuple[int, int]]:
    return [shard_range(ncw, r, world, align) for r in range(world)]


def bind_to_gpu_numa_node(device: int) -> dict:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that pinned staging buffers
    allocated afterwards are node-local and host-to-device copies of several ranks do not cross the socket
    interconnect.  Host-side placement only; returns what it found (empty dict if the topology is not exposed)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = device
        if visible:
            ids = [x for x in visible.split(",") if x.strip() != ""]
            if device < len(ids) and ids[device].strip().isdigit():
                index = int(ids[device])
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return {"pci": bus, "numa_node": node}
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"pci": bus, "numa_node": node, "cpus": len(allowed)}
    except Exception as e:  # topology not exposed (containers, single-socket hosts): leave the affinity alone
        return {"error": type(e).__name__}
