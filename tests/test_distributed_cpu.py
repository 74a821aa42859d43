"""Host-side multi-GPU logic on CPU: world-size-2 gloo processes shard an ensemble, evolve their
realisations (here with the CPU oracle standing in for the device, which CI boxes lack) and
gather per-realisation results; the assembled arrays must equal the unsharded run."""

import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from frictionqpotspringblock_b200.distributed import shard_realisations, shard_seed


def test_shards_partition_the_ensemble():
    for total in (0, 1, 7, 16384, 16385):
        for world in (1, 2, 3, 8):
            covered = []
            for rank in range(world):
                first, count = shard_realisations(total, rank, world)
                covered += list(range(first, first + count))
            assert covered == list(range(total))
            counts = [shard_realisations(total, r, world)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1
    assert shard_seed(5, 3, 4096) == 5 + 3 * 4096
    with pytest.raises(ValueError):
        shard_realisations(4, 2, 2)


def _worker(rank, world, port, total, N, out):
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from frictionqpotspringblock_b200.distributed import gather_per_realisation
    from oracle import oracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_realisations(total, rank, world)
    S = np.empty(count, dtype=np.int64)
    uf = np.empty(count, dtype=np.float64)
    for k in range(count):
        s = orc.Line1d.System_Cuspy_Laplace(
            m=1.0, eta=0.35, mu=1.0, k_interactions=1.0, k_frame=1.0 / N, dt=0.1, shape=[N],
            seed=shard_seed(0, first + k, N), distribution="random", parameters=[2.0], offset=-50)
        s.u_frame = 0.5
        s.minimise()
        i_n = s.chunk.index_at_align
        s.eventDrivenStep(1e-3, False)
        s.eventDrivenStep(1e-3, True)
        s.minimise()
        S[k] = np.sum(s.chunk.index_at_align - i_n)
        uf[k] = s.u_frame
    S_all = gather_per_realisation(S, total)
    uf_all = gather_per_realisation(uf, total)
    dist.barrier()
    if rank == 0:
        np.savez(out, S=S_all, uf=uf_all)
    dist.destroy_process_group()


def test_world_size_2_gloo_gather_matches_unsharded(tmp_path):
    total, N = 5, 64
    out = str(tmp_path / "gathered.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, total, N, out), nprocs=2, join=True)
    got = np.load(out)

    from oracle import oracle as orc

    S = np.empty(total, dtype=np.int64)
    uf = np.empty(total)
    for r in range(total):
        s = orc.Line1d.System_Cuspy_Laplace(
            m=1.0, eta=0.35, mu=1.0, k_interactions=1.0, k_frame=1.0 / N, dt=0.1, shape=[N],
            seed=r * N, distribution="random", parameters=[2.0], offset=-50)
        s.u_frame = 0.5
        s.minimise()
        i_n = s.chunk.index_at_align
        s.eventDrivenStep(1e-3, False)
        s.eventDrivenStep(1e-3, True)
        s.minimise()
        S[r] = np.sum(s.chunk.index_at_align - i_n)
        uf[r] = s.u_frame
    assert np.array_equal(got["S"], S)
    assert np.array_equal(got["uf"], uf)


def test_make_sharded_passes_disjoint_seeds_and_forcing():
    """make_sharded hands every rank its realisations' disorder seeds and, for thermal ensembles,
    their forcing seeds and schedule slices (host logic only: a recording stand-in for the class)"""
    from frictionqpotspringblock_b200.distributed import make_sharded

    class Recorder:
        def __init__(self, **kw):
            self.kw = kw

    N, total = 8, 5
    dinc_init = np.arange(total * N).reshape(total, N)
    seen = []
    for rank in range(2):
        obj, first, count = make_sharded(Recorder, total, rank, 2, seed=7, shape=[N],
                                         seed_forcing=100, dinc_init=dinc_init,
                                         dinc=np.ones(N, dtype=int))
        assert obj.kw["nrealisations"] == count and obj.kw["seed"] == 7 + first * N
        assert obj.kw["seed_forcing"] == 100 + first
        assert np.array_equal(obj.kw["dinc_init"], dinc_init[first:first + count])
        assert obj.kw["dinc"].shape == (N,)  # a shared schedule is passed through
        seen += list(range(first, first + count))
    assert seen == list(range(total))


def test_make_sharded_follows_a_non_default_seed_stride():
    """Realisation r of an Ensemble uses initstates seed + r*seed_stride + p (include/fqsb.h): a
    shard must start at seed + first*seed_stride whatever the stride, or ranks would reuse
    initstates of other ranks. A rank without realisations is an error (nrealisations = 0 would be
    read as 1 by fqsb_create)."""
    from frictionqpotspringblock_b200.distributed import make_sharded

    class Recorder:
        def __init__(self, **kw):
            self.kw = kw

    N, total, stride, seed = 8, 7, 1000, 11
    unsharded = [seed + r * stride for r in range(total)]
    got = []
    for rank in range(3):
        obj, first, count = make_sharded(Recorder, total, rank, 3, seed=seed, shape=[N],
                                         seed_stride=stride)
        assert obj.kw["seed_stride"] == stride
        got += [obj.kw["seed"] + k * stride for k in range(count)]
    assert got == unsharded
    with pytest.raises(ValueError):
        make_sharded(Recorder, 2, 2, 3, seed=0, shape=[N])
