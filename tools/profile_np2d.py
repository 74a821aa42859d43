import sys
import numpy as np
sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F
n = 4096 * 4096
s = F.Line2d.System_Cuspy_Laplace_Nopassing(mu=1.0, k_interactions=1.0, k_frame=1.0 / n, shape=[4096, 4096],
                                            distribution="random", parameters=[2.0], offset=-50, seed=0)
s.u_frame = 1.0
s.minimise(max_iter=40, max_iter_is_error=False)
print(s.last_kernel, s.last_kernel_seconds / s.last_kernel_launches)
d = F.Line2d.System_Cuspy_Laplace(m=1.0, eta=0.35, dt=0.1, mu=1.0, k_interactions=1.0, k_frame=1.0 / n, shape=[4096, 4096],
                                  distribution="random", parameters=[2.0], offset=-50, seed=0)
d.u_frame = 1.0
d.timeSteps(12)
