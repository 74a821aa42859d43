"""ncu target: no-passing sweeps of a 4096 x 4096 interface AT the fixed point (no well changes:
the streaming rate of the sweep kernel itself)."""
import sys

sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F  # noqa: E402

n = 4096 * 4096
s = F.Line2d.System_Cuspy_Laplace_Nopassing(mu=1.0, k_interactions=1.0, k_frame=1.0 / n,
                                            shape=[4096, 4096], distribution="random",
                                            parameters=[2.0], offset=-50, seed=0)
s.u_frame = 1.0
s.minimise(max_iter=400, max_iter_is_error=False)
s.minimise(tol=1e-300, max_iter=12, max_iter_is_error=False)
print(s.last_kernel, s.last_kernel_seconds / s.last_kernel_launches)
