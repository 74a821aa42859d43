"""Throughput of flowSteps in sliding motion (every block changes wells every ~10 steps at
v_frame = 1): resident kernel on an ensemble, blocked kernel on one 2^20 line."""
import sys

import numpy as np

sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F  # noqa: E402

R, N, T = 2368, 4096, 500
kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_interactions=1.0, k_frame=1.0 / N,
          dt=0.1, shape=[N], distribution="random", parameters=[2.0], offset=-50, seed=0)
ens = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, **kw)
ens.timeSteps(T)
ens.timeSteps(T)
quiet = ens.last_kernel_seconds
for v_frame in (0.1, 1.0):
    ens.u_frame = ens.u_frame + 0.35 * N
    ens.flowSteps(T, v_frame)
    ens.mark_indices()
    ens.flowSteps(T, v_frame)
    sec = ens.last_kernel_seconds
    S, A = ens.avalanche_since_mark()
    print(f"resident v_frame={v_frame}: {R*N*T/sec:.3e} upd/s, hops/update {np.sum(S)/(R*N*T):.4f} "
          f"(quiescent {R*N*T/quiet:.3e})", flush=True)
del ens
N = 1 << 20
kw.update(shape=[N], k_frame=1.0 / N)
kw.pop("k_interactions")
s = F.Line1d.System_Cuspy_Quartic(a1=1.0, a2=1.0, **kw)
s.timeSteps(1024)
s.timeSteps(1024)
quiet = s.last_kernel_seconds / 1024
s.u_frame = 0.35 * N
s.flowSteps(1024, 1.0)
s.mark_indices()
s.flowSteps(1024, 1.0)
S, A = s.avalanche_since_mark()
print(f"blocked 2^20 v_frame=1: {1e6*s.last_kernel_seconds/1024:.2f} us/step, hops/update "
      f"{S/(N*1024):.4f} (quiescent {1e6*quiet:.2f} us/step)")
