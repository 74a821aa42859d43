"""Thermal systems (External = RandomNormalForcing): device time per step of flowSteps for the
reference's example (one line of N = 1000) and for an ensemble."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F  # noqa: E402


def run(N, R, steps):
    kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_interactions=1.0, k_frame=1.0 / N,
              dt=0.1, shape=[N], seed=0, distribution="random", parameters=[2.0], offset=-50)
    rng = np.random.default_rng(0)
    f = dict(mean=0.0, stddev=0.05, seed_forcing=0, dinc_init=rng.integers(0, 100, N),
             dinc=100 * np.ones(N, dtype=np.int64))
    if R == 1:
        s = F.Line1d.System_Cuspy_Laplace_RandomForcing(**kw, **f)
    else:
        s = F.Line1d.Ensemble_Cuspy_Laplace_RandomForcing(nrealisations=R, **kw, **f)
    s.flowSteps(steps, 5e-2)
    t0 = time.perf_counter()
    s.flowSteps(steps, 5e-2)
    w = time.perf_counter() - t0
    dev = s.last_kernel_seconds
    print(f"thermal N={N} R={R}: {s.last_kernel} {dev / steps * 1e6:.2f} us/step device, "
          f"{w / steps * 1e6:.2f} wall, {N * R * steps / dev:.3e} block-updates/s, "
          f"T={np.mean(s.temperature):.3e}", flush=True)


run(1000, 1, 2000)
run(4096, 1024, 200)
