"""Config #3 (one line of N = 2^20 blocks): temporally blocked kernel (K2b) against the
one-step-per-launch streaming kernel. Prints device time per step for fixed-step calls and for
the stop modes, for several batch lengths (steps per launch)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, a1=1.0, a2=1.0, k_frame=1.0 / N,
          dt=0.1, shape=[N], distribution="random", parameters=[2.0], offset=-50, seed=0)
T = 2048
for label, kernel in [("stream", 2), ("blocked k=16", 3 | (16 << 8)), ("blocked k=32", 3 | (32 << 8)),
                      ("blocked k=48", 3 | (48 << 8)), ("blocked k=64", 3 | (64 << 8))]:
    if R == 1:
        s = F.Line1d.System_Cuspy_Quartic(kernel=kernel, **kw)
        s.u_frame = 0.5
    else:
        s = F.Line1d.Ensemble_Cuspy_Quartic(kernel=kernel, nrealisations=R, **kw)
        s.u_frame = np.full(R, 0.5)
    s.timeSteps(256)
    t0 = time.perf_counter()
    s.timeSteps(T)
    w = time.perf_counter() - t0
    k_fixed = s.last_kernel_seconds / T
    line = (f"{label:13s} {s.last_kernel:10s} fixed: {k_fixed * 1e6:7.2f} us/step device, "
            f"{w / T * 1e6:7.2f} wall, {N * R / k_fixed:.3e} upd/s")
    t0 = time.perf_counter()
    s.minimise(tol=1e-300, max_iter=T, max_iter_is_error=False)
    w = time.perf_counter() - t0
    line += f" | stop: {s.last_kernel_seconds / T * 1e6:7.2f} us/step device, {w / T * 1e6:7.2f} wall"
    t0 = time.perf_counter()
    r = s.minimise()
    w = time.perf_counter() - t0
    inc = int(np.max(s.inc))
    line += f" | minimise(): ret {np.max(r)}, inc {inc}, {s.last_kernel_launches} launches, {w * 1e3:.2f} ms"
    print(line, flush=True)
