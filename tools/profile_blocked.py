"""ncu target for the temporally blocked kernel (K2b) on config #3: one Cuspy_Quartic line of
N = 2^20 blocks, k steps per launch (argv[1], default 64); fixed-step launches, then stop-mode
launches."""
import sys

import numpy as np

sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = 1 << 20
kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, a1=1.0, a2=1.0, k_frame=1.0 / N,
          dt=0.1, shape=[N], distribution="random", parameters=[2.0], offset=-50, seed=0)
s = F.Line1d.System_Cuspy_Quartic(kernel=3 | (K << 8), **kw)
s.u_frame = 0.5
s.timeSteps(4 * K)
s.timeSteps(4 * K)
print("fixed", s.last_kernel, N * 4 * K / s.last_kernel_seconds)
s.minimise(tol=1e-300, max_iter=4 * K, max_iter_is_error=False)
print("stop", s.last_kernel, N * 4 * K / s.last_kernel_seconds)
