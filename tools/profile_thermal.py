"""ncu target: the thermal resident kernel (K2t) on the reference's example (N = 1000) and on
an ensemble of 592 x N = 4096."""
import sys

import numpy as np

sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F  # noqa: E402

for N, R, steps in ((1000, 1, 2000), (4096, 592, 200)):
    kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_interactions=1.0, k_frame=1.0 / N,
              dt=0.1, shape=[N], seed=0, distribution="random", parameters=[2.0], offset=-50)
    rng = np.random.default_rng(0)
    f = dict(mean=0.0, stddev=0.05, seed_forcing=0, dinc_init=rng.integers(0, 100, N),
             dinc=100 * np.ones(N, dtype=np.int64))
    s = F.Line1d.Ensemble_Cuspy_Laplace_RandomForcing(nrealisations=R, **kw, **f)
    s.flowSteps(steps, 5e-2)
    s.flowSteps(steps, 5e-2)
    print(N, R, s.last_kernel, N * R * steps / s.last_kernel_seconds)
