"""Latency of ONE small line (BASELINE config #1: N = 1000) per resident-kernel configuration."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F
from tests import protocol

for variant, name in ((0, "default"), (2, "(2,512)"), (3, "(2,1024)")):
    kw = dict(protocol.BASE)
    s = F.Line1d.System_Cuspy_Laplace(k_interactions=1.0, kernel=1 + 16 * variant, **kw)
    t0 = time.perf_counter()
    u, f, S = protocol.run(s, 120)
    dt = time.perf_counter() - t0
    print(f"{name:10s} 120 protocol steps: {dt:.3f} s, {s.inc} Verlet steps, "
          f"{dt / s.inc * 1e6:.3f} us/step, S sum {S.sum()}", flush=True)
