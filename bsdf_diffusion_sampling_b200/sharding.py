"""Multi-GPU sharding of a query batch (one process per GPU, ``torch.distributed``).

Every query is independent (state = 2 floats + a running determinant product; weights are read-only),
so the batch is split into contiguous row blocks, one per rank, and each rank runs the single-GPU kernel
on its block -- there is NO collective on the hot path.  The Philox counter of a query is its GLOBAL row
index (``first_index`` = start of the shard), so results are bit-identical for every shard count.
The only communication is the optional final gather of ``wo [N,3]`` + ``pdf [N]`` (16 B/query), one
``all_gather_into_tensor`` over NCCL/NVLink (``gloo`` on CPU for the host-logic tests).

The reference has no multi-GPU code at all (device hard-coded to "cuda:0",
rendering/brdf_measured_disk.py:10); this module is the data-parallel layer SURVEY.md 8(e) adds.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin the calling process to the CPU cores of the NUMA node the GPU hangs off (PCI bus id ->
    /sys/bus/pci/devices/<id>/numa_node -> node cpulist).  Pinned host buffers allocated afterwards land on that
    node, so the host<->device copies of ``NeuralBSDFSampler.sample_host`` do not cross the inter-socket link when
    several ranks stream at once.  Returns the node, or None when the topology is not exposed (single node, VM)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:                                  # noqa: BLE001  (best effort: never fail the caller)
        return None


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of rank ``rank``: the first ``n % world`` ranks get one extra row."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    q, r = divmod(int(n), world)
    start = rank * q + min(rank, r)
    return start, start + q + (1 if rank < r else 0)


def shard_rows(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    a, b = shard_range(t.shape[0], rank, world)
    return t[a:b]


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather the row blocks produced by ``shard_range`` back into one [n_total, ...] tensor.

    Shards may differ by one row; they are padded to the largest block for a single
    ``all_gather_into_tensor`` and the padding rows are dropped afterwards."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    q, r = divmod(int(n_total), world)
    mx = q + (1 if r else 0)
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = local.new_empty((world * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad, group=group)
    pieces = []
    for k in range(world):
        a, b = shard_range(n_total, k, world)
        pieces.append(out[k * mx: k * mx + (b - a)])
    return torch.cat(pieces, 0)


class ShardedSampler:
    """Run a ``plugins.NeuralBSDFSampler`` over this rank's shard of a global query batch."""

    def __init__(self, sampler, rank: Optional[int] = None, world: Optional[int] = None, group=None):
        self.sampler = sampler
        self.group = group
        init = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank(group) if init else 0)
        self.world = world if world is not None else (dist.get_world_size(group) if init else 1)

    def local_range(self, n_total: int) -> Tuple[int, int]:
        return shard_range(n_total, self.rank, self.world)

    def sample_local(self, wi_local: torch.Tensor, n_total: int, seed: int, offset: int = 0, x0_local=None):
        """wi_local = this rank's rows of the global wi.  -> (wo_local, pdf_local)."""
        start, stop = self.local_range(n_total)
        if wi_local.shape[0] != stop - start:
            raise ValueError(f"rank {self.rank}: expected {stop - start} local rows, got {wi_local.shape[0]}")
        return self.sampler.sample(wi_local, x0=x0_local, seed=seed, offset=offset, first_index=start)

    def pdf_local(self, wi_local: torch.Tensor, wo_local: torch.Tensor):
        return self.sampler.pdf(wi_local, wo_local)

    def sample(self, wi_global: torch.Tensor, seed: int, offset: int = 0, gather: bool = True):
        """Convenience: slice the (replicated) global wi, sample the shard, optionally gather."""
        n = wi_global.shape[0]
        a, b = self.local_range(n)
        wo, pdf = self.sample_local(wi_global[a:b], n, seed, offset)
        if gather and self.world > 1:
            return gather_rows(wo, n, self.group), gather_rows(pdf, n, self.group)
        return wo, pdf
