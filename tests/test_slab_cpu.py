"""CPU tests of the slab-decomposition host logic: the library's StopList replay
(fqsb_slab_first_stop, used per batch by fqsb_slab_minimise) against the oracle's own stopping
step, the row/seed plan of the members, and -- under gloo with world sizes 2 and 3 -- the one piece
of plumbing the caller provides: the all-gather of the members' 64-byte IPC handles."""

import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from frictionqpotspringblock_b200 import slab
from oracle import oracle as orc


def test_first_stop_replays_the_reference_criterion():
    """Feed the residual history of an oracle minimisation to the library's host-side replay: it
    must stop at exactly the step the oracle's minimise() stops (detail.h:1764-1784)."""
    N = 200
    kw = dict(m=1.0, eta=0.35, mu=1.0, k_interactions=1.0, k_frame=1.0 / N, dt=0.1, shape=[N],
              seed=4, distribution="random", parameters=[2.0], offset=-50)
    s = orc.Line1d.System_Cuspy_Laplace(**kw)
    s.u_frame = 0.5
    t = orc.Line1d.System_Cuspy_Laplace(**kw)
    t.u_frame = 0.5
    assert s.minimise() == 0
    nsteps = s.inc
    for niter_tol, k in ((10, 16), (10, 7)):
        t = orc.Line1d.System_Cuspy_Laplace(**kw)
        t.u_frame = 0.5
        ring = slab.StopList(niter_tol)
        found, done = 0, 0
        while not found:
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


def test_first_stop_takes_any_niter_tol():
    """The reference's StopList is unbounded (detail.h:1676-1689): 50 residuals below tol in
    non-increasing order stop at step 50, a single increase among them delays the stop."""
    k, n = 80, 50
    log = np.zeros((k, slab.NLOG))
    log[:, 0] = 1e-12 * np.exp(-0.01 * np.arange(k))
    log[:, 1] = 1.0
    assert slab.first_stop(log, slab.StopList(n), 1e-5) == n
    log[20, 0] = 1.5e-12  # residual rises once at step 21 (still below tol, above tol^2)
    assert slab.first_stop(log, slab.StopList(n), 1e-5) == 20 + n
    # all_less(tol^2) stops whatever the order
    log[:, 0] = 1e-22 * (1.0 + 0.5 * np.sin(np.arange(k)))
    assert slab.first_stop(log, slab.StopList(n), 1e-5) == n


def test_slab_plan_tiles_the_system():
    for shape, world, halo in (([23, 5], 3, 3), ([4096, 4096], 8, 32), ([1 << 20], 8, 64),
                               ([10], 2, 2)):
        n = int(np.prod(shape))
        unit = n // shape[0]
        covered = []
        for rank in range(world):
            p = slab.slab_plan(shape, rank, world, halo)
            assert p["local_shape"][0] == p["cnt"] + 2 * halo and p["seed_period"] == n
            local = (p["seed_first"] + np.arange((p["cnt"] + 2 * halo) * unit)) % n
            own = local[p["own"][0]:p["own"][1]]
            covered.append(own)
            # the top halo mirrors the rows just above the owned range (periodic)
            assert local[0] == ((p["lo"] - halo) * unit) % n
            assert p["halo_cells"] == halo * unit
        assert np.array_equal(np.concatenate(covered), np.arange(n))
    with pytest.raises(ValueError):
        slab.slab_plan([8], 0, 4, 3)  # members would own fewer rows than the halo


def test_unsupported_classes_are_refused():
    """LongRange couples every pair of blocks and the thermal classes draw from one ordered pcg32
    stream: a halo of rows cannot reproduce them, so SlabSystem refuses them up front."""
    for module, cls in (("Line1d", "System_Cuspy_LongRange"),
                        ("Line1d", "System_Cuspy_Laplace_RandomForcing"),
                        ("Line1d", "System_Cuspy_Quartic_RandomForcing")):
        with pytest.raises(RuntimeError, match="slab decomposition is not available"):
            slab.SlabSystem(module, cls, devices=[0], shape=[64])


def _worker(rank, world, port, out):
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from frictionqpotspringblock_b200.distributed import allgather_bytes

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    handle = bytes([rank]) * 64  # stands in for fqsb_slab_ipc_handle (needs a GPU)
    every = allgather_bytes(handle)
    plan = slab.slab_plan([23, 5], rank, world, 3)
    dist.barrier()
    np.save(f"{out}.{rank}.npy", np.frombuffer(b"".join(every), dtype=np.uint8))
    np.save(f"{out}.plan.{rank}.npy", np.array([plan["lo"], plan["cnt"], plan["seed_first"]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ipc_handle_allgather_under_gloo(world, tmp_path):
    port = 29800 + (os.getpid() % 1000) + world
    out = str(tmp_path / "handles")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    want = np.repeat(np.arange(world, dtype=np.uint8), 64)
    rows = 0
    for rank in range(world):
        assert np.array_equal(np.load(f"{out}.{rank}.npy"), want)  # rank order, on every rank
        lo, cnt, first = np.load(f"{out}.plan.{rank}.npy")
        assert lo == rows and first == ((lo - 3) % 23) * 5
        rows += cnt
    assert rows == 23
