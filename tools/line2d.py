"""Times the 2-D streaming kernel (BASELINE config #5 geometry: 4096 x 4096 interface)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = int(sys.argv[3]) if len(sys.argv) > 3 else 50
n = rows * cols
kw = dict(eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_interactions=1.0, k_frame=1.0 / n,
          shape=[rows, cols], distribution="random", parameters=[2.0], offset=-50, seed=0)
s = F.Line2d.System_Cuspy_Laplace(m=1.0, dt=0.1, **kw)
s.u_frame = 1.0
s.timeSteps(5)
s.timeSteps(T)
sec = s._get_scalars  # noqa
sec = s.last_kernel_seconds / T
print(f"{s.last_kernel}: {rows}x{cols}: {sec * 1e6:.1f} us/step, {n / sec:.3e} block-updates/s, "
      f"{64 * n / sec / 1e9:.0f} GB/s algorithmic")
import time
t0 = time.perf_counter()
ret = s.minimise(max_iter=2000, max_iter_is_error=False)
dt = time.perf_counter() - t0
print(f"minimise (stop mode): ret={ret} steps={s.inc - 5 - T} {dt:.3f}s -> "
      f"{(s.inc - 5 - T) * n / dt:.3e} block-updates/s")
del s
kw.pop("eta")
s = F.Line2d.System_Cuspy_Laplace_Nopassing(**kw)
s.u_frame = 1.0
t0 = time.perf_counter()
ret = s.minimise(max_iter=400, max_iter_is_error=False)
dt = time.perf_counter() - t0
print(f"{s.last_kernel}: no-passing sweeps: ret={ret} launches={s.last_kernel_launches} "
      f"{s.last_kernel_seconds / max(1, s.last_kernel_launches) * 1e6:.1f} us/sweep "
      f"({n * s.last_kernel_launches / s.last_kernel_seconds:.3e} block-updates/s)")
# steady state: sweeps at the fixed point move (almost) no block between wells, so they show the
# streaming rate of the kernel itself (32 B per block-update)
s.minimise(tol=1e-300, max_iter=60, max_iter_is_error=False)
per = s.last_kernel_seconds / max(1, s.last_kernel_launches)
print(f"{s.last_kernel}: sweeps at the fixed point: {s.last_kernel_launches} launches, "
      f"{per * 1e6:.1f} us/sweep = {n / per:.3e} block-updates/s = {32 * n / per / 1e9:.0f} GB/s")
