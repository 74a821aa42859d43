"""Small deterministic workload for ncu captures: launches, in order,
k_resident (minimise), k_resident (timeSteps 200) x2, k_stream_step x 10."""
import sys

import numpy as np

sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 592
N = 4096
kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_interactions=1.0, k_frame=1.0 / N,
          dt=0.1, shape=[N], distribution="random", parameters=[2.0], offset=-50, seed=0)
ens = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, **kw)
ens.minimise()
ens.eventDrivenStep(1e-3, False)
ens.eventDrivenStep(1e-3, True)
ens.timeSteps(200)
ens.timeSteps(200)
print("resident", R * N * 200 / ens.last_kernel_seconds)
del ens
ens = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=4 * R, kernel=2, **kw)
ens.u_frame = np.full(4 * R, 1.0)
ens.timeSteps(10)
ens.timeSteps(10)
print("stream", 4 * R * N * 10 / ens.last_kernel_seconds)
