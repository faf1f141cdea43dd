"""Multi-GPU plan for a batch of independent proteins: proteins never interact, so each rank runs
its own share (longest-processing-time-first by edge count) and the per-protein [L,21] tables meet in ONE
all-gather at the end -- ``all_gather_into_tensor`` on a preallocated flat buffer, in place: every rank's engine writes
its tables straight into its own slot of that buffer (``TMPNN_OUT_DEVICE``), so between the head kernel and the
collective nothing touches the host.  The collective is whatever backend the process group was created with (NCCL over
NVLink on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

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


class ShardPlan:
    """Who computes which protein, and where each table sits in the flat [world, max_rows, 21] gather buffer."""

    def __init__(self, lengths: Sequence[int], world: int, rank: int):
        self.lengths = [int(l) for l in lengths]
        self.world, self.rank = world, rank
        self.plan = lpt_assign(self.lengths, world)
        self.rows = [sum(self.lengths[i] for i in b) for b in self.plan]
        self.max_rows = max(max(self.rows), 1)
        self.mine = self.plan[rank]

    def unpack(self, flat: np.ndarray) -> List[np.ndarray]:
        """flat [world, max_rows, 21] (host) -> the tables in input order."""
        out: List[Optional[np.ndarray]] = [None] * len(self.lengths)
        for r in range(self.world):
            o = 0
            for i in self.plan[r]:
                out[i] = flat[r, o:o + self.lengths[i]]
                o += self.lengths[i]
        return out   # type: ignore


def gather_in_place(flat: torch.Tensor, rank: int) -> None:
    """ONE collective: every rank contributes flat[rank] and receives all slots (no-op for a single process)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_gather_into_tensor(flat.view(-1), flat[rank].reshape(-1))


def sharded_tables(pdbs: Sequence[dict], compute: Callable[[Sequence[dict]], List[np.ndarray]], device=None) -> List[np.ndarray]:
    """Every rank calls this with the SAME list of proteins and gets back ALL [L,21] tables, in input order.
    ``compute`` maps a list of proteins to their tables on this rank (host arrays); this generic form stages them into
    the gather buffer.  The CUDA path proper is ``ShardedScan`` (device-resident end to end)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    plan = ShardPlan([len(p["seq"]) for p in pdbs], world, rank)
    tables = compute([pdbs[i] for i in plan.mine]) if plan.mine else []
    if world == 1:
        return list(tables)
    dev = device if device is not None else torch.device("cpu")
    flat = torch.zeros((world, plan.max_rows, 21), dtype=torch.float32, device=dev)
    if tables:
        flat[rank, :plan.rows[rank]] = torch.from_numpy(np.concatenate(tables, 0)).to(dev)
    gather_in_place(flat, rank)
    host = flat.cpu().numpy()
    return [t.copy() for t in plan.unpack(host)]


class ShardedScan:
    """Site-saturation tables of a batch of proteins on this process group's GPUs, device-resident end to end.

    Built once per batch shape (plan, packed inputs of this rank's share, the flat gather buffer and a pinned host
    mirror); ``run()`` then costs: one H2D of the packed inputs, the kernel chain writing into this rank's slot, ONE
    all_gather_into_tensor over NVLink, one D2H of the gathered tables.
    """

    def __init__(self, model, pdbs: Sequence[dict], device: Optional[torch.device] = None):
        from .featurize import featurize_batch
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.plan = ShardPlan([len(p["seq"]) for p in pdbs], self.world, self.rank)
        self.engine = model._get_engine()
        self.device = device if device is not None else torch.device("cuda", self.engine.device)
        self.packed = featurize_batch([pdbs[i] for i in self.plan.mine]) if self.plan.mine else None
        self.flat = torch.zeros((self.world, self.plan.max_rows, 21), dtype=torch.float32, device=self.device)
        self.host = torch.empty((self.world, self.plan.max_rows, 21), dtype=torch.float32).pin_memory()
        self._dev_in = None
        self._stream = torch.cuda.Stream(device=self.device)   # run(): kernels, collective and the D2H are ordered on this stream

    # -- inputs ---------------------------------------------------------------------------------------------------
    def host_inputs(self):
        """Pinned host copies of this rank's packed inputs (what a caller hands to ``run``)."""
        p = self.packed
        return {k: torch.from_numpy(np.ascontiguousarray(getattr(p, k))).pin_memory() for k in ("X", "S", "residue_idx", "chain_enc", "mask")}

    def device_inputs(self):
        if self._dev_in is None:
            p = self.packed
            self._dev_in = {k: torch.from_numpy(np.ascontiguousarray(getattr(p, k))).to(self.device)
                            for k in ("X", "S", "residue_idx", "chain_enc", "mask")}
        return self._dev_in

    # -- steps ----------------------------------------------------------------------------------------------------
    def compute_local(self, inputs=None, stream: int = 0, device_inputs: bool = False):
        """This rank's kernel chain, writing straight into its slot of the gather buffer (no synchronisation)."""
        from . import engine as E
        if self.packed is None:
            return
        slot = self.flat[self.rank]
        flags = E.OUT_DEVICE | E.NO_SYNC | (E.IN_DEVICE if device_inputs else 0)
        a = inputs if inputs is not None else (self.device_inputs() if device_inputs else self.host_inputs())
        arg = (lambda t: t.data_ptr()) if device_inputs else (lambda t: t.numpy())
        self.engine.run(self.packed.offsets, arg(a["X"]), arg(a["S"]), arg(a["residue_idx"]), arg(a["chain_enc"]), arg(a["mask"]),
                        slot.data_ptr(), flags=flags, stream=stream)

    def gather(self):
        gather_in_place(self.flat, self.rank)

    def run(self, inputs=None) -> List[np.ndarray]:
        """All tables, in input order, on the host (every rank gets all of them)."""
        st = self._stream
        with torch.cuda.stream(st):
            self.compute_local(inputs, stream=st.cuda_stream)
            self.gather()
            self.host.copy_(self.flat, non_blocking=True)
        st.synchronize()
        return [t.copy() for t in self.plan.unpack(self.host.numpy())]
