"""Slab-decomposed large systems over the ranks of a torchrun job (BASELINE configs #3 and #5).

    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/slab_bench.py
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from frictionqpotspringblock_b200.slab import SlabSystem  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(name, module, cls, shape, halo, T, **extra):
    n = int(np.prod(shape))
    kw = dict(mu=1.0, k_frame=1.0 / n, shape=shape, seed=0, distribution="random",
              parameters=[2.0], offset=-50, **extra)
    if "Nopassing" not in cls:
        kw.update(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, dt=0.1)
    s = SlabSystem(module, cls, halo=halo, device=local, **kw)
    s.u_frame = 1.0
    out = {"case": name, "shape": shape, "gpus": world, "halo": halo}
    if "Nopassing" not in cls:
        s.timeSteps(2 * s.batch)
        barrier()
        t0 = time.perf_counter()
        s.timeSteps(T)
        barrier()
        dt = time.perf_counter() - t0
        out.update(us_per_step=1e6 * dt / T, block_updates_per_s=n * T / dt)
    barrier()
    t0 = time.perf_counter()
    inc0 = s.inc
    steps0 = s.sys.step_count
    ret = s.minimise(max_iter=3000, max_iter_is_error=False)
    barrier()
    dt = time.perf_counter() - t0
    steps = s.sys.step_count - steps0
    out.update(minimise_ret=int(ret), minimise_steps=int(steps), minimise_s=dt,
               minimise_block_updates_per_s=n * steps / dt, residual=s.residual)
    if rank == 0:
        print(json.dumps(out), flush=True)
    del s


halo = int(sys.argv[1]) if len(sys.argv) > 1 else 32
run("config5_verlet", "Line2d", "System_Cuspy_Laplace", [4096, 4096], halo, 256,
    k_interactions=1.0)
run("config5_nopassing", "Line2d", "System_Cuspy_Laplace_Nopassing", [4096, 4096], halo, 0,
    k_interactions=1.0)
run("config3_quartic", "Line1d", "System_Cuspy_Quartic", [1 << 20], 64, 1024, a1=1.0, a2=1.0)
run("config3_semismooth", "Line1d", "System_SemiSmooth_Laplace", [1 << 20], 64, 1024,
    k_interactions=1.0, kappa=1.0)
if world > 1:
    dist.destroy_process_group()
