"""Config #3 (N = 2^20, Cuspy_Quartic): temporally blocked kernel for several tile geometries
(kernel bits 8..15 = steps per launch, bits 16..31 = owned blocks per tile, 0 = planner)."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F  # noqa: E402

N = int(os.environ.get("FQSB_N", 1 << 20))
kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, a1=1.0, a2=1.0, k_frame=1.0 / N,
          dt=0.1, shape=[N], distribution="random", parameters=[2.0], offset=-50, seed=0)
T = 2048
ks = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else (32, 48, 64)
owns = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else (0,)
contracted = len(sys.argv) > 3 and sys.argv[3] == "fma"
for k in ks:
    for own in owns:
        s = F.Line1d.System_Cuspy_Quartic(kernel=3 | (k << 8) | (own << 16), contracted=contracted, **kw)
        s.u_frame = 0.5
        s.timeSteps(256)
        s.timeSteps(T)
        fixed = s.last_kernel_seconds / T
        s.minimise(tol=1e-300, max_iter=T, max_iter_is_error=False)
        stop = s.last_kernel_seconds / T
        print(f"k={k} own={own:5d}: fixed {fixed * 1e6:6.2f} us/step, stop {stop * 1e6:6.2f} us/step",
              flush=True)
        del s
