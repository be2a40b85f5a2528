"""Frame-range sharding across the GPUs of one box.

Every frame is independent (reference ``OcrRecogniser.predict`` keeps no state across calls, backend/tools/ocr.py:24-86),
so GPU r of R owns the contiguous frame range [r*N/R, (r+1)*N/R) and there is NO collective on the hot path.  The only
exchanges are (1) a one-time broadcast of the packed plans from rank 0 (NCCL over NVLink when the ranks hold GPUs, gloo
on CPU) and (2) gathering the per-rank result lines, merged by frame number because the reference's de-dup
(backend/main.py:774-818) assumes frame-ordered lines.  ``torch.distributed`` is plumbing only.
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence, Tuple

import numpy as np


def frame_range(rank: int, world: int, n_frames: int) -> Tuple[int, int]:
    """Contiguous, balanced, exhaustive: union over ranks = [0, n_frames), sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return rank * n_frames // world, (rank + 1) * n_frames // world


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device: int) -> Optional[int]:
    """NUMA node the GPU's PCIe root hangs off (/sys/bus/pci/devices/<bdf>/numa_node), or None when it cannot be told."""
    import os
    import subprocess
    bdf = None
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        bdf = None
    if bdf is None or not os.path.exists(f"/sys/bus/pci/devices/{bdf}/numa_node"):
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device)],
                                 capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0].strip().lower()
            bdf = out[-12:] if len(out) >= 12 else out          # 00000000:1b:00.0 -> 0000:1b:00.0
        except Exception:
            return None
    try:
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa_node(device: int) -> Optional[int]:
    """Restrict this process to the CPUs of the GPU's NUMA node BEFORE it allocates page-locked frame buffers: first-touch then
    places them in the memory next to the GPU's PCIe root, so that the eight concurrent host->device copies of an 8-GPU box do
    not cross the socket interconnect (the end-to-end limiter at N = 8).  Returns the node, or None (nothing
    changed) when the topology cannot be read or the node's CPUs are not available to this process."""
    import os
    node = gpu_numa_node(device)
    if node is None or os.environ.get("VSE_NO_NUMA_BIND"):
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read()) & set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def _dist():
    import torch.distributed as dist
    return dist


def is_distributed() -> bool:
    dist = _dist()
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_blobs(blobs: Optional[Sequence[bytes]], src: int = 0, device: Optional[str] = None) -> List[bytes]:
    """Rank ``src`` passes the packed plans; every rank returns the same bytes."""
    import torch
    dist = _dist()
    if not is_distributed():
        return list(blobs or [])
    dev = torch.device(device) if device else torch.device("cpu")
    if dist.get_rank() == src:
        sizes = torch.tensor([len(blobs)] + [len(b) for b in blobs] + [0] * (15 - len(blobs)), dtype=torch.int64, device=dev)
    else:
        sizes = torch.zeros(16, dtype=torch.int64, device=dev)
    dist.broadcast(sizes, src)
    n = int(sizes[0])
    out: List[bytes] = []
    for k in range(n):
        size = int(sizes[1 + k])
        if dist.get_rank() == src:
            t = torch.frombuffer(bytearray(blobs[k]), dtype=torch.uint8).to(dev)
        else:
            t = torch.empty(size, dtype=torch.uint8, device=dev)
        dist.broadcast(t, src)
        out.append(t.cpu().numpy().tobytes())
    return out


def gather_by_frame(local: Sequence[Tuple[int, Any]]) -> List[Tuple[int, Any]]:
    """local: [(global frame number, payload)] of this rank -> all ranks' items sorted by frame number (on every rank)."""
    dist = _dist()
    if not is_distributed():
        return sorted(local, key=lambda t: t[0])
    bucket: List[Any] = [None] * dist.get_world_size()
    dist.all_gather_object(bucket, list(local))
    merged = [item for part in bucket for item in part]
    merged.sort(key=lambda t: t[0])
    return merged


def max_over_ranks(value: float, device: Optional[str] = None) -> float:
    import torch
    dist = _dist()
    if not is_distributed():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=torch.device(device) if device else torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
