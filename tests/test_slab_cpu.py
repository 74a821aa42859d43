"""CPU tests of the slab-decomposition host logic (frictionqpotspringblock_b200/slab.py): the
StopList replay against the oracle's own stopping step, and the halo exchange pattern under
gloo with world sizes 2 and 3 (numpy arrays stand in for the device state)."""

import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from frictionqpotspringblock_b200 import slab
from oracle import oracle as orc


def test_first_stop_replays_the_reference_criterion():
    """Feed the residual history of an oracle minimisation to the host-side replay: it must stop
    at exactly the step the oracle's minimise() stops (detail.h:1764-1784)."""
    N = 200
    s = orc.Line1d.System_Cuspy_Laplace(
        m=1.0, eta=0.35, mu=1.0, k_interactions=1.0, k_frame=1.0 / N, dt=0.1, shape=[N], seed=4,
        distribution="random", parameters=[2.0], offset=-50)
    s.u_frame = 0.5
    t = orc.Line1d.System_Cuspy_Laplace(
        m=1.0, eta=0.35, mu=1.0, k_interactions=1.0, k_frame=1.0 / N, dt=0.1, shape=[N], seed=4,
        distribution="random", parameters=[2.0], offset=-50)
    t.u_frame = 0.5
    assert s.minimise() == 0
    nsteps = s.inc
    ring = slab.StopList(10)
    found, done = 0, 0
    while not found:
        k = 16
        log = np.zeros((k, slab.NLOG))
        for j in range(k):
            t.timeStep()
            log[j, 0] = np.sum(t.f ** 2)
            log[j, 1] = np.sum(t.f_frame ** 2)
        stop = slab.first_stop(log, ring, 1e-5)
        if stop:
            found = done + stop
        done += k
    assert found == nsteps
    with pytest.raises(RuntimeError, match="NaN entries found"):
        slab.first_stop(np.full((2, slab.NLOG), np.nan), slab.StopList(3), 1e-5)


def test_halo_plan_orders_receives_from_next_first():
    for world in (2, 3, 8):
        for rank in range(world):
            sends, recvs = slab.halo_plan(rank, world)
            assert sends == [((rank - 1) % world, "top"), ((rank + 1) % world, "bottom")]
            assert recvs == [((rank + 1) % world, "bottom_halo"), ((rank - 1) % world, "top_halo")]


def _worker(rank, world, port, rows_total, halo, unit, out):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from frictionqpotspringblock_b200 import slab as sl
    from frictionqpotspringblock_b200.distributed import shard_realisations

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, cnt = shard_realisations(rows_total, rank, world)
    local_rows = cnt + 2 * halo
    # 7 planes of the local state; plane q of global cell c holds 1000*q + c, halos start as -1
    state = -np.ones((7, local_rows * unit), dtype=np.int64)
    glob = (np.arange(lo * unit, (lo + cnt) * unit))
    for q in range(7):
        state[q, halo * unit:(halo + cnt) * unit] = 1000000 * q + glob

    def export_cells(first, count, tensor):
        tensor.copy_(torch.from_numpy(state[:, first:first + count].reshape(-1).copy()))

    def import_cells(first, count, tensor):
        state[:, first:first + count] = tensor.numpy().reshape(7, count)

    k = halo * unit
    own = (halo * unit, (halo + cnt) * unit)
    layout = {"top": (own[0], k), "bottom": (own[1] - k, k), "top_halo": (0, k),
              "bottom_halo": (own[1], k)}
    sl.exchange_halos(export_cells, import_cells, layout, rank, world)
    dist.barrier()
    np.save(f"{out}.{rank}.npy", state)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_under_gloo(world, tmp_path):
    rows_total, halo, unit = 23, 3, 5
    port = 29800 + (os.getpid() % 1000) + world
    out = str(tmp_path / "state")
    mp.spawn(_worker, args=(world, port, rows_total, halo, unit, out), nprocs=world, join=True)
    from frictionqpotspringblock_b200.distributed import shard_realisations

    n = rows_total * unit
    for rank in range(world):
        state = np.load(f"{out}.{rank}.npy")
        lo, cnt = shard_realisations(rows_total, rank, world)
        first = ((lo - halo) * unit) % n
        expect = (first + np.arange((cnt + 2 * halo) * unit)) % n
        for q in range(7):
            assert np.array_equal(state[q], 1000000 * q + expect), (rank, q)
