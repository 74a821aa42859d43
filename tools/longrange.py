"""Times the LongRange DMMA GEMM step (BASELINE config #4: N=8192 x 1024 realisations, alpha=1.5)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
T = int(sys.argv[3]) if len(sys.argv) > 3 else 5
kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_interactions=1.0, alpha=1.5,
          k_frame=1.0 / N, dt=0.1, shape=[N], distribution="random", parameters=[2.0], offset=-50,
          seed=0)
ens = F.Line1d.Ensemble_Cuspy_LongRange(nrealisations=R, **kw)
ens.u_frame = np.full(R, 1.0)
ens.timeSteps(2)
ens.timeSteps(T)
sec = ens.last_kernel_seconds / T
flop = 2.0 * N * N * R
print(f"{ens.last_kernel}: N={N} R={R}: {sec * 1e3:.3f} ms/step, {flop / sec / 1e12:.2f} TFLOP/s "
      f"(2 N^2 R), {N * R / sec:.3e} block-updates/s")
