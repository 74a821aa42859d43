"""Streaming path, one large line: fixed-step launches vs stop-mode launches (finalise + polling)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F
N = 1 << 20
kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, a1=1.0, a2=1.0, k_frame=1.0 / N,
          dt=0.1, shape=[N], distribution="random", parameters=[2.0], offset=-50, seed=0)
s = F.Line1d.System_Cuspy_Quartic(**kw)
s.u_frame = 0.5
s.timeSteps(200)
for T in (2000,):
    t0 = time.perf_counter(); s.timeSteps(T); w = time.perf_counter() - t0
    print(f"fixed: kernel {s.last_kernel_seconds/T*1e6:.2f} us/step, wall {w/T*1e6:.2f} us/step")
    t0 = time.perf_counter(); s.minimise(tol=1e-300, max_iter=T, max_iter_is_error=False); w = time.perf_counter() - t0
    print(f"stop : kernel {s.last_kernel_seconds/T*1e6:.2f} us/step, wall {w/T*1e6:.2f} us/step")
    t0 = time.perf_counter(); r = s.minimise(); w = time.perf_counter() - t0
    print(f"minimise(): ret {r}, {s.last_kernel_launches} launches, wall {w*1e3:.2f} ms, {w/s.last_kernel_launches*1e6:.2f} us/launch")
