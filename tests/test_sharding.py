"""CPU: the N>1 path (plan + single all_gather) with a world_size-2 gloo group and a fake per-rank compute."""
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from thermompnn_b200.sharding import lpt_assign, protein_cost, sharded_tables
from thermompnn_b200.synth import synth


def _fake_compute(pdbs):
    # deterministic stand-in for TransferModel.head_tables: a function of the protein only
    return [np.outer(np.arange(len(p["seq"]), dtype=np.float32), np.ones(21, np.float32)) + float(sum(map(ord, p["seq"])) % 97)
            for p in pdbs]


def _worker(rank, world, port, lengths, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pdbs = [synth(L, seed=100 + i) for i, L in enumerate(lengths)]
    out = sharded_tables(pdbs, _fake_compute)
    ref = _fake_compute(pdbs)
    ok = all(np.array_equal(a, b) for a, b in zip(out, ref)) and len(out) == len(ref)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_lpt_balances_and_covers():
    lengths = [500, 40, 300, 299, 32, 480, 150, 72]
    plan = lpt_assign(lengths, 3)
    assert sorted(i for b in plan for i in b) == list(range(len(lengths)))
    load = [sum(protein_cost(lengths[i]) for i in b) for b in plan]
    assert max(load) <= 1.34 * sum(load) / 3                         # LPT bound 4/3 - 1/(3m)
    assert lpt_assign(lengths, 1) == [list(range(len(lengths)))]
    assert protein_cost(40) == 1600 and protein_cost(300) == 14400


@pytest.mark.parametrize("lengths", [[60, 20, 45, 33, 70], [25]])
def test_two_rank_gather_equals_single_process(lengths):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + len(lengths)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs: p.join(timeout=60)
    assert res == [(0, True), (1, True)]
