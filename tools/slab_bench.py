"""Slab-decomposed large systems (BASELINE configs #3 and #5) through fqsb_slab_*.

    python tools/slab_bench.py [--members G] [--halo H]           one process drives G members
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/slab_bench.py
                                                                  one process per GPU (CUDA IPC)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F  # noqa: E402
from frictionqpotspringblock_b200.slab import SlabSystem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=0)
ap.add_argument("--halo", type=int, default=32)
ap.add_argument("--cases", default="config5_verlet,config5_nopassing,config3_quartic")
ap.add_argument("--steps", type=int, default=512)
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if dist is not None:
        dist.barrier()


def run(name, module, cls, shape, halo, T, **extra):
    n = int(np.prod(shape))
    kw = dict(mu=1.0, k_frame=1.0 / n, shape=shape, seed=0, distribution="random",
              parameters=[2.0], offset=-50, **extra)
    if "Nopassing" not in cls:
        kw.update(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, dt=0.1)
    if world > 1:
        from frictionqpotspringblock_b200.distributed import allgather_bytes

        s = SlabSystem(module, cls, halo=halo, rank=rank, world=world, device=local,
                       allgather=allgather_bytes, **kw)
        G = world
    else:
        G = args.members or F.device_count()
        ngpu = F.device_count()
        s = SlabSystem(module, cls, halo=halo, devices=[g % ngpu for g in range(G)], **kw)
    s.u_frame = 1.0
    out = {"case": name, "shape": shape, "members": G, "processes": world, "halo": halo,
           "gpus": min(G, F.device_count())}
    if "Nopassing" not in cls:
        s.timeSteps(3 * s.batch)
        barrier()
        t0 = time.perf_counter()
        s.timeSteps(T)
        barrier()
        dt = time.perf_counter() - t0
        out.update(us_per_step=1e6 * dt / T, block_updates_per_s=n * T / dt)
    barrier()
    t0 = time.perf_counter()
    ret = s.minimise(max_iter=3000, max_iter_is_error=False)
    barrier()
    dt = time.perf_counter() - t0
    steps = s.last_minimise_steps
    out.update(minimise_ret=int(ret), minimise_steps=int(steps), minimise_s=dt,
               minimise_us_per_step=1e6 * dt / max(1, steps),
               minimise_block_updates_per_s=n * steps / dt, residual=s.residual,
               u_frame=s.u_frame, info=s.info())
    if rank == 0:
        print(json.dumps(out), flush=True)
    del s


cases = args.cases.split(",")
if "config5_verlet" in cases:
    run("config5_verlet", "Line2d", "System_Cuspy_Laplace", [4096, 4096], args.halo, args.steps,
        k_interactions=1.0)
if "config5_nopassing" in cases:
    run("config5_nopassing", "Line2d", "System_Cuspy_Laplace_Nopassing", [4096, 4096],
        args.halo + 1, 0, k_interactions=1.0)
if "config3_quartic" in cases:
    run("config3_quartic", "Line1d", "System_Cuspy_Quartic", [1 << 20], 64, 2 * args.steps,
        a1=1.0, a2=1.0)
if "config3_semismooth" in cases:
    run("config3_semismooth", "Line1d", "System_SemiSmooth_Laplace", [1 << 20], 64,
        2 * args.steps, k_interactions=1.0, kappa=1.0)
if dist is not None:
    dist.destroy_process_group()
