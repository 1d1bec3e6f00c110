"""NUMA placement of a rank's host side (one process per GPU).

The end-to-end path moves ~94 MB per 4K frame between pinned host memory and the GPU.  With
eight ranks on one node the pinned framebuffer mirrors and staging buffers must live on the
NUMA node the rank's GPU hangs off, and the rank's threads must run there; otherwise every
rank's DMA crosses the socket interconnect and lands on one memory controller.  `bind_to_gpu`
pins the calling process to the CPUs of its GPU's NUMA node and makes that node the preferred
one for the allocations that follow (pinned allocations are backed by ordinary host pages, so
call it BEFORE the library allocates framebuffers).  Pure host logic: sysfs + libnuma via ctypes
where present; a box without the information leaves the process unbound and says so.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path


def parse_cpulist(text: str) -> list[int]:
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11]"""
    cpus: list[int] = []
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-", 1)
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_pci_bus_id(index: int) -> str | None:
    try:
        out = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip()
    except (OSError, subprocess.SubprocessError):
        return None
    if not out:
        return None
    # nvidia-smi prints an 8-digit domain (00000000:1B:00.0); sysfs uses 4 (0000:1b:00.0)
    dom, rest = out.split(":", 1)
    return f"{dom[-4:]}:{rest}".lower()


def gpu_numa_node(index: int, sysfs: Path = Path("/sys/bus/pci/devices")) -> int | None:
    bus = gpu_pci_bus_id(index)
    if bus is None:
        return None
    try:
        node = int((sysfs / bus / "numa_node").read_text().strip())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def node_cpus(node: int, sysfs: Path = Path("/sys/devices/system/node")) -> list[int]:
    try:
        return parse_cpulist((sysfs / f"node{node}" / "cpulist").read_text())
    except OSError:
        return []


def split_evenly(cpus: list[int], parts: int, which: int) -> list[int]:
    """`which`-th of `parts` contiguous, balanced slices (every slice non-empty if len >= parts)"""
    if parts <= 1 or len(cpus) < parts:
        return list(cpus)
    base, extra = divmod(len(cpus), parts)
    start = which * base + min(which, extra)
    return cpus[start:start + base + (1 if which < extra else 0)]


def bind_to_gpu(local_rank: int, ranks_on_node: int | None = None) -> str:
    """CPU affinity + preferred memory node for the calling process; returns what was done."""
    node = gpu_numa_node(local_rank)
    if node is None:
        return "unbound (no NUMA information for the GPU)"
    cpus = node_cpus(node)
    allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
    if not allowed:
        return f"unbound (NUMA node {node} has no CPU this process may use)"
    what = f"cpus of NUMA node {node} ({len(allowed)})"
    try:
        os.sched_setaffinity(0, allowed)
    except OSError as e:
        return f"unbound (sched_setaffinity: {e})"
    try:
        libnuma = ctypes.CDLL("libnuma.so.1")
        if libnuma.numa_available() >= 0:
            libnuma.numa_set_preferred(ctypes.c_int(node))
            what += ", preferred memory node set (libnuma)"
    except OSError:
        what += ", memory follows first touch (no libnuma)"
    return what
