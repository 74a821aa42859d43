"""Per-step cost of the resident kernel in stop mode vs fixed mode (forced step count)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
T = 500
N = 4096
kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_interactions=1.0, k_frame=1.0 / N,
          dt=0.1, shape=[N], distribution="random", parameters=[2.0], offset=-50, seed=0)
ens = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, **kw)
ens.minimise()
ens.eventDrivenStep(1e-3, False)
ens.eventDrivenStep(1e-3, True)
ens.timeSteps(T)
ens.timeSteps(T)
fixed = ens.last_kernel_seconds
# tol so small that the criterion never fires: exactly T steps in MODE_MINIMISE
ens.minimise(tol=1e-300, max_iter=T, max_iter_is_error=False)
ens.minimise(tol=1e-300, max_iter=T, max_iter_is_error=False)
stop = ens.last_kernel_seconds
ens.timeStepsUntilEvent(tol=1e-300, max_iter=T)
print(f"fixed {R*N*T/fixed:.3e} upd/s ({fixed/T*148/R*1e6:.3f} us/step/CTA)  "
      f"minimise-mode {R*N*T/stop:.3e} upd/s ({stop/T*148/R*1e6:.3f} us/step/CTA)  ratio {stop/fixed:.2f}")
