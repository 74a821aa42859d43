"""Multi-GPU plumbing for ensembles: independent disorder realisations are the natural shard
(SURVEY.md section 8e). One process per GPU; rank g owns a contiguous range of realisations
and the matching disjoint pcg32 seeds; nothing is exchanged during the evolution, per-realisation
results (return codes, S, A, u_frame, mean f_frame) are gathered afterwards.

``torch.distributed`` is plumbing only (NCCL on the GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

import numpy as np


def shard_realisations(total: int, rank: int, world: int) -> tuple[int, int]:
    """(first, count) of the realisations rank `rank` integrates: contiguous, balanced to one."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def shard_seed(seed: int, first: int, stride: int) -> int:
    """Seed of a shard whose first realisation has global index `first`: realisation r of the
    whole ensemble uses initstates seed + r*stride + p (Line1d.h:148-151 per realisation;
    stride = the ensemble's ``seed_stride``, by default the number of blocks)."""
    return int(seed) + int(first) * int(stride)


def make_sharded(cls, total_realisations: int, rank: int, world: int, *, seed: int, **kw):
    """Construct this rank's shard of an ``Ensemble_*`` class."""
    first, count = shard_realisations(total_realisations, rank, world)
    if count == 0:
        raise ValueError(f"rank {rank} of {world} would own no realisation of {total_realisations}")
    size = int(np.prod(kw["shape"]))
    stride = int(kw.get("seed_stride", 0) or size)  # the rule of fqsb_create: 0 reads as size
    if "seed_forcing" in kw:
        # thermal ensembles: realisation r draws its random forces from
        # pcg32(seed_forcing + r * seed_forcing_stride); per-realisation schedules are sliced
        fstride = int(kw.get("seed_forcing_stride", 1) or 1)
        kw["seed_forcing"] = int(kw["seed_forcing"]) + first * fstride
        for key in ("dinc_init", "dinc"):
            arr = np.asarray(kw[key])
            if arr.ndim == len(kw["shape"]) + 1:
                kw[key] = arr[first:first + count]
    return cls(nrealisations=count, seed=shard_seed(seed, first, stride), **kw), first, count


def gather_per_realisation(local: np.ndarray, total: int, group=None) -> np.ndarray:
    """All ranks receive the [total] array assembled from every rank's [count] slice."""
    import torch
    import torch.distributed as dist

    local = np.ascontiguousarray(local)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local.copy()
    world = dist.get_world_size(group)
    device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    counts = [shard_realisations(total, r, world)[1] for r in range(world)]
    width = max(counts)
    pad = np.zeros(width, dtype=local.dtype)
    pad[: local.size] = local
    mine = torch.from_numpy(pad).to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return np.concatenate([p.cpu().numpy()[:c] for p, c in zip(parts, counts)])


def allgather_bytes(payload: bytes, group=None) -> list:
    """Every rank receives the list of all ranks' ``payload`` (rank order). Used once per slab-
    decomposed system to move the 64-byte CUDA IPC handles of the members' mailboxes."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [bytes(payload)]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, bytes(payload), group=group)
    return [bytes(b) for b in out]
