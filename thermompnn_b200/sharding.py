"""Multi-GPU plan for a batch of independent proteins: proteins never interact, so each rank runs
its own share and the per-protein [L,21] tables meet in ONE all_gather at the end.  Nothing here touches the GPU;
the collective is whatever backend the process group was created with (NCCL on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def protein_cost(length: int, k: int = 48) -> int:
    """Work of one protein ~ its edge count L * min(k, L)."""
    return int(length) * min(k, int(length))


def lpt_assign(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first bin packing: protein indices per rank (each list in ascending index order)."""
    order = sorted(range(len(lengths)), key=lambda i: (-protein_cost(lengths[i]), i))
    load = [0] * world
    bins: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        bins[r].append(i)
        load[r] += protein_cost(lengths[i])
    return [sorted(b) for b in bins]


def sharded_tables(pdbs: Sequence[dict], compute: Callable[[Sequence[dict]], List[np.ndarray]], device=None) -> List[np.ndarray]:
    """Every rank calls this with the SAME list of proteins and gets back ALL [L,21] tables, in input order.
    ``compute`` maps a list of proteins to their tables on this rank's GPU (``TransferModel.head_tables``)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lengths = [len(p["seq"]) for p in pdbs]
    plan = lpt_assign(lengths, world)
    mine = plan[rank]
    tables = compute([pdbs[i] for i in mine]) if mine else []
    if world == 1:
        return list(tables)
    # one collective: every rank's tables concatenated and padded to the largest share
    rows = [sum(lengths[i] for i in plan[r]) for r in range(world)]
    dev = device if device is not None else torch.device("cpu")
    send = torch.zeros((max(rows), 21), dtype=torch.float32, device=dev)
    if tables:
        send[:rows[rank]] = torch.from_numpy(np.concatenate(tables, 0)).to(dev)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    out: List[np.ndarray] = [None] * len(pdbs)       # type: ignore
    for r in range(world):
        buf = recv[r].cpu().numpy()
        o = 0
        for i in plan[r]:
            out[i] = buf[o:o + lengths[i]].copy()
            o += lengths[i]
    return out
