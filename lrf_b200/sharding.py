"""Multi-GPU plumbing: images are independent (no cross-image term anywhere in
lrf/compression/qmf.py:116-292), so a batch is cut into contiguous per-rank ranges with no collective on
the data path; the only communication is one all_gather of the per-image (bpp, PSNR) pairs at the end."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Images [lo, hi) of rank `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_stats(stats: torch.Tensor, total: int, rank: int, world: int) -> torch.Tensor:
    """all_gather of ragged per-rank (n_i, k) float tensors → (total, k) in image order on every rank."""
    if world == 1:
        return stats
    k = stats.shape[1]
    longest = (total + world - 1) // world
    padded = torch.zeros((longest, k), dtype=stats.dtype, device=stats.device)
    padded[: stats.shape[0]] = stats
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    parts = []
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        parts.append(bufs[r][: hi - lo])
    return torch.cat(parts)


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device_index: int) -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that page-locked buffers it allocates
    afterwards (first touch) and the threads that feed cudaMemcpyAsync sit next to the GPU's PCIe root port.  With one
    process per GPU this is what keeps the host-to-device copies of 8 ranks from crossing the socket interconnect.
    Best effort: returns what was done ({"numa_node": n, "cpus": k} or {"numa_node": None, "why": ...})."""
    import os

    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"numa_node": None, "why": "sysfs reports no NUMA node for " + bdf}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return {"numa_node": node, "why": "no allowed CPU on that node"}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus), "pci": bdf}
    except (OSError, AttributeError, ValueError) as e:  # no sysfs / old torch / restricted container
        return {"numa_node": None, "why": f"{type(e).__name__}: {e}"}
