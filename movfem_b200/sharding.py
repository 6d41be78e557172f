"""Multi-GPU sharding of the assembly path: plumbing over torch.distributed.

* frequency sharding (config 4): the mesh, pattern and cached K_e/M_e are replicated; rank r takes
  frequencies r+1, r+1+W, ... (1-based, round-robin).  Shards are independent: the only history the
  reference's sequential loop carries is the host's g_sigma and the stale GPML flags seen by
  element (1,1,1) (Q17), both inputs of movfem_assemble.  No collective on the data path; finished
  triplet values go to the MUMPS host rank (send/recv over NCCL on GPUs, gloo in the CPU tests).
* slab sharding (config 5): contiguous ranges of the reference's ``ie`` index; x-slabs own contiguous
  row ranges because DOFs are numbered in (ie,je,ke) first-encounter order (global_assembly.f90:237-296).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def frequency_shard(nf: int, rank: int, world: int) -> list[int]:
    """1-based frequency indices handled by ``rank`` (round-robin deal)."""
    return list(range(rank + 1, nf + 1, world))


def slab_partition(n_ie: int, rank: int, world: int) -> tuple[int, int]:
    """Inclusive 1-based ``ie`` range of the x-slab owned by ``rank`` (empty: hi < lo)."""
    base, rem = divmod(n_ie, world)
    lo = rank * base + min(rank, rem) + 1
    hi = lo + base + (1 if rank < rem else 0) - 1
    return lo, hi


def gather_to_root(results: dict[int, torch.Tensor], nf: int, root: int = 0) -> dict[int, torch.Tensor]:
    """Hand the per-frequency value arrays to ``root`` (the rank attached to ZMUMPS).  Sizes differ per
    frequency (zero stripping), so a length is sent first.  Point-to-point only."""
    rank, world = dist.get_rank(), dist.get_world_size()
    out = dict(results) if rank == root else {}
    for ifreq in range(1, nf + 1):
        owner = (ifreq - 1) % world
        if owner == root:
            continue
        if rank == owner:
            t = results[ifreq]
            n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
            dist.send(n, dst=root)
            dist.send(t, dst=root)
        elif rank == root:
            dev = next(iter(results.values())).device if results else torch.device("cpu")
            n = torch.zeros(1, dtype=torch.int64, device=dev)
            dist.recv(n, src=owner)
            t = torch.empty(int(n.item()), dtype=torch.float64, device=dev)
            dist.recv(t, src=owner)
            out[ifreq] = t
    return out
