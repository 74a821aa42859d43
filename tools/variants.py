"""Times the resident-kernel configurations (blocks/thread x threads x wells-in-smem) on one GPU.

    python tools/variants.py [R] [T]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
N = 4096
kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_interactions=1.0, k_frame=1.0 / N,
          dt=0.1, shape=[N], distribution="random", parameters=[2.0], offset=-50, seed=0)
names = {0: "default", 4: "(8,512)"}
for variant in (0,):
    ens = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, kernel=1 + 16 * variant, **kw)
    t0 = time.perf_counter()
    ens.minimise()
    tmin = time.perf_counter() - t0
    steps_min = ens.step_count
    kmin = ens.last_kernel_seconds
    ens.eventDrivenStep(1e-3, False)
    ens.eventDrivenStep(1e-3, True)
    ens.timeSteps(T)
    best = 1e9
    for _ in range(3):
        ens.timeSteps(T)
        best = min(best, ens.last_kernel_seconds)
    t0 = time.perf_counter()
    ret = ens.minimise()
    steps2 = ens.step_count
    print(f"variant {variant} {names[variant]:16s} fixed: {R * N * T / best:.3e} upd/s "
          f"({best / T * 148 / R * 1e6 * 1:.3f} us/step/CTA)   "
          f"minimise: {steps_min * N / kmin:.3e} upd/s (kernel {kmin:.3f}s wall {tmin:.3f}s, "
          f"{steps_min / R:.0f} steps/realisation)", flush=True)
    del ens
