"""CPU check of the equivalence the device relies on for Particles.System_SemiSmooth /
System_Smooth (Particles.h:233-311): a line with k_interactions = 0 integrates exactly like the
interaction-free system (the interaction term is an exact +-0)."""

import numpy as np
import pytest

from oracle import oracle as orc

PHYS = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, dt=0.1, seed=3, distribution="random",
            parameters=[2.0], offset=-50)


@pytest.mark.parametrize("cls,line,extra", [
    ("System_Cuspy", "System_Cuspy_Laplace", dict()),
    ("System_SemiSmooth", "System_SemiSmooth_Laplace", dict(kappa=0.9)),
    ("System_Smooth", "System_Smooth_Laplace", dict()),
])
def test_line_with_zero_coupling_equals_particles(cls, line, extra):
    N = 200
    kw = dict(shape=[N], k_frame=1.0 / N, **extra, **PHYS)
    a = getattr(orc.Particles, cls)(**kw)
    b = getattr(orc.Line1d, line)(k_interactions=0.0, **kw)
    for s in (a, b):
        s.u_frame = 30.0
        s.timeSteps(300)
    for name in ("u", "v", "a", "f", "f_potential", "f_frame", "f_damping"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert np.all(b.f_interactions == 0.0)
    assert np.array_equal(a.chunk.index_at_align, b.chunk.index_at_align)
    if cls != "System_Smooth":
        assert a.minimise() == b.minimise() == 0
        assert a.inc == b.inc
        assert np.array_equal(a.u, b.u)
