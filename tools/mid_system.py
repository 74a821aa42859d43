"""Step latency of one line of N blocks through the resident kernel (fixed and minimise mode)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F
for N in (2560, 2561, 3000, 3072, 3584, 4000, 4095, 4096):
    kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_interactions=1.0, k_frame=1.0 / N,
              dt=0.1, shape=[N], distribution="random", parameters=[2.0], offset=-50, seed=0)
    s = F.Line1d.System_Cuspy_Laplace(**kw)
    s.u_frame = 0.5
    s.timeSteps(2000)
    s.timeSteps(20000)
    fixed = s.last_kernel_seconds / 20000
    s.minimise(tol=1e-300, max_iter=20000, max_iter_is_error=False)
    stop = s.last_kernel_seconds / 20000
    print(f"N={N:5d}: fixed {fixed*1e6:.3f} us/step, minimise-mode {stop*1e6:.3f} us/step", flush=True)
