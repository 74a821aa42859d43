"""Timing probe: one 2^20 Quartic line as a plain handle and as slabs of 1 / 2 members (one GPU)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F  # noqa: E402
from frictionqpotspringblock_b200.slab import SlabSystem  # noqa: E402

N = 1 << 20
kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, dt=0.1, a1=1.0, a2=1.0, k_frame=1.0 / N,
          shape=[N], seed=0, distribution="random", parameters=[2.0], offset=-50)
T = 2048


def timed(label, s):
    s.u_frame = 1.0
    s.timeSteps(256)
    t0 = time.perf_counter()
    s.timeSteps(T)
    dt = time.perf_counter() - t0
    print(f"{label}: {1e6 * dt / T:.2f} us/step", flush=True)


timed("plain handle (blocked)", F.Line1d.System_Cuspy_Quartic(**kw))
for members in (1, 2, 4):
    for kernel, batch in ((None, None), (None, 32), (2, None)):
        s = SlabSystem("Line1d", "System_Cuspy_Quartic", halo=64, devices=[0] * members,
                       kernel=kernel, batch=batch, **kw)
        timed(f"slab members={members} kernel={kernel} batch={s.batch}", s)
        del s
