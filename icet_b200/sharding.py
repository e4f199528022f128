"""Pair sharding across ranks (one process per GPU) and the final gather of poses / covariances.

Scan pairs are independent units (each reference `ICET` object is self-contained, include/icet.h:43), so the
path shards by contiguous pair range with no data-path collective; the only communication is one all_gather of
48 floats per pair (X 6 | pred_stds 6 | Q 36) after the local pairs are done.  Backend-agnostic
(`nccl` on GPUs, `gloo` in the CPU tests).
"""
from __future__ import annotations


def shard_range(total_pairs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous pair range [lo, hi) of `rank`: GPU g of G gets pairs [g*P/G, (g+1)*P/G) and therefore scans
    lo .. hi (one boundary scan is shared with the next rank)."""
    if world < 1 or not (0 <= rank < world) or total_pairs < 0:
        raise ValueError("bad shard arguments")
    lo = (total_pairs * rank) // world
    hi = (total_pairs * (rank + 1)) // world
    return lo, hi


def shard_scans(total_pairs: int, rank: int, world: int) -> tuple[int, int]:
    """(first_scan, nscans) a rank must hold to register its pair range."""
    lo, hi = shard_range(total_pairs, rank, world)
    return lo, (hi - lo + 1) if hi > lo else 0


def gather_results(local, world: int):
    """all_gather of the per-pair result rows (equal shard sizes) into [world * P_local, C] in global pair order.
    `local` is a [P_local, C] tensor on the backend's device; returns `local` itself when world == 1."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out
