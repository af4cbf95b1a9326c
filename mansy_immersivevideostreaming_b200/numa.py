"""NUMA placement of pinned host buffers (Linux): a rank's pinned rollout ring belongs on the memory node its GPU hangs off,
otherwise every device-to-host copy crosses the socket interconnect (the host-buffer path is bound by
that link).  Pure host plumbing: sysfs + the ``set_mempolicy`` / ``move_pages`` system calls through ctypes (no libnuma).
Every function degrades to a no-op where the information or the permission is missing."""
from __future__ import annotations

import contextlib
import ctypes
import os
from typing import Dict, Iterator, List, Optional, Set

_SYS_SET_MEMPOLICY = 238          # x86_64
_SYS_GET_MEMPOLICY = 239
_SYS_MOVE_PAGES = 279
_MPOL_DEFAULT, _MPOL_PREFERRED, _MPOL_BIND = 0, 1, 2


def _libc():
    return ctypes.CDLL(None, use_errno=True)


def _parse_list(text: str) -> Set[int]:
    out: Set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.update(range(int(lo), int(hi or lo) + 1))
    return out


def online_nodes() -> List[int]:
    try:
        return sorted(_parse_list(open("/sys/devices/system/node/online").read()))
    except OSError:
        return [0]


def node_cpus(node: int) -> Set[int]:
    try:
        return _parse_list(open(f"/sys/devices/system/node/node{node}/cpulist").read())
    except OSError:
        return set()


def mems_allowed() -> Optional[str]:
    try:
        for line in open("/proc/self/status"):
            if line.startswith("Mems_allowed_list"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return None


def gpu_numa_node(device: int) -> Optional[int]:
    """Memory node of CUDA device ``device`` (sysfs ``numa_node`` of its PCI function) or None when unknown (-1)."""
    import torch
    p = torch.cuda.get_device_properties(device)
    try:
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
    except (OSError, AttributeError, ValueError):
        return None
    return node if node >= 0 else None


@contextlib.contextmanager
def memory_on_node(node: Optional[int]) -> Iterator[bool]:
    """Allocations (page faults / pinned allocations) of the calling thread prefer ``node`` inside the block.  Yields
    whether the policy was applied."""
    if node is None or node not in online_nodes() or len(online_nodes()) < 2:
        yield False
        return
    libc = _libc()
    mask = ctypes.c_ulong(1 << node)
    rc = libc.syscall(_SYS_SET_MEMPOLICY, _MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask) + 1))
    ok = rc == 0
    try:
        yield ok
    finally:
        if ok:
            libc.syscall(_SYS_SET_MEMPOLICY, _MPOL_DEFAULT, None, ctypes.c_ulong(0))


def pages_node(ptr: int, nbytes: int, samples: int = 16) -> Dict[int, int]:
    """Which nodes hold the buffer: histogram over ``samples`` pages (move_pages with a NULL target list queries)."""
    page = os.sysconf("SC_PAGE_SIZE")
    n_pages = max(1, nbytes // page)
    idx = sorted({int(i * (n_pages - 1) / max(1, samples - 1)) for i in range(samples)})
    pages = (ctypes.c_void_p * len(idx))(*[(ptr // page + i) * page for i in idx])
    status = (ctypes.c_int * len(idx))()
    rc = _libc().syscall(_SYS_MOVE_PAGES, 0, ctypes.c_ulong(len(idx)), pages, None, status, 0)
    out: Dict[int, int] = {}
    if rc == 0:
        for s in status:
            out[int(s)] = out.get(int(s), 0) + 1
    return out
