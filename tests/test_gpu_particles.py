"""GPU parity of the particle systems (no interactions; /root/reference/include/
FrictionQPotSpringBlock/Particles.h:93-311): the reference's own tests/test_Particles.py cases
that are not already covered by the Line1d ports, and every class against the oracle."""

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import assert_same_state, product

pytestmark = pytest.mark.gpu

PHYS = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, dt=0.1, seed=3, distribution="random",
            parameters=[2.0], offset=-50)


def test_uniform_init_all_particle_classes():
    """tests/test_Particles.py:25-75 (Test_Uniform.test_init)"""
    F = product()
    N = 5
    par = dict(m=1, eta=0.37, mu=0.81, k_frame=0.23, dt=1.0, shape=[N], seed=0,
               distribution="delta", parameters=[1.0], offset=-49.5, nchunk=100)
    rand = dict(mean=0, stddev=1, seed_forcing=0, dinc_init=np.ones(N, dtype=int),
                dinc=np.ones(N, dtype=int))
    systems = [
        F.Particles.System_Cuspy(**par),
        F.Particles.System_SemiSmooth(kappa=1, **par),
        F.Particles.System_Smooth(**par),
        F.Particles.System_Cuspy_RandomForcing(**par, **rand),
    ]
    for system in systems:  # by construction u = 0 is a local minimum of all potentials
        assert system.residual < 1e-5
        for name in ("f", "f_potential", "f_frame", "f_interactions", "f_damping"):
            assert np.allclose(getattr(system, name), 0.0), name
        assert np.all(system.chunk.index_at_align + 1 == np.argmax(system.chunk.data[0, :] > 0))
        assert np.all(system.chunk.right_of_align > 0)
        assert np.all(system.chunk.left_of_align <= 0)


def test_semismooth_event_driven_step():
    """tests/test_Particles.py:317-371 (Test_System_SemiSmooth.test_eventDrivenStep)"""
    F = product()
    N, mu, kappa = 3, 1, 0.1
    system = F.Particles.System_SemiSmooth(
        m=1.0, eta=1.0, mu=mu, kappa=kappa, k_frame=0.1, dt=1.0, shape=[N], seed=0,
        distribution="delta", parameters=[1.0], offset=-49.5, nchunk=100)
    assert system.residual < 1e-5
    u0 = system.u.copy()
    uf0 = system.u_frame
    left = system.chunk.left_of_align
    right = system.chunk.right_of_align
    mid = 0.5 * (left + right)
    upper = (mu * mid + kappa * right) / (mu + kappa)
    lower = (mu * mid + kappa * left) / (mu + kappa)
    eps = 0.001
    assert np.isclose(system.maxUniformDisplacement(), np.min(upper - system.u))
    system.eventDrivenStep(eps=eps, kick=False)
    assert system.residual < 1e-5
    assert np.allclose(system.u, upper - 0.5 * eps)
    assert np.isclose(system.maxUniformDisplacement(), 0.5 * eps)
    system.eventDrivenStep(eps=eps, kick=True)
    assert system.residual > 1e-5
    assert np.allclose(system.u, upper + 0.5 * eps)
    assert abs(system.maxUniformDisplacement()) < 1e-7
    system.u = u0
    system.u_frame = uf0
    assert np.isclose(system.maxUniformDisplacement(-1), np.min(system.u - lower))
    system.eventDrivenStep(eps=eps, kick=False, direction=-1)
    assert system.residual < 1e-5
    assert np.allclose(system.u, lower + 0.5 * eps)
    assert np.isclose(system.maxUniformDisplacement(-1), 0.5 * eps)
    system.eventDrivenStep(eps=eps, kick=True, direction=-1)
    assert system.residual > 1e-5
    assert np.allclose(system.u, lower - 0.5 * eps)
    assert abs(system.maxUniformDisplacement()) < 1e-7


@pytest.mark.parametrize("cls,extra,exact", [
    ("System_Cuspy", dict(), True),
    ("System_SemiSmooth", dict(kappa=0.9), True),
    ("System_Smooth", dict(), False),  # sin(): 1e-12, not bit-exact
])
@pytest.mark.parametrize("kernel", [1, 2], ids=["resident", "stream"])
def test_particle_dynamics_match_oracle(cls, extra, exact, kernel):
    """The oracle integrates Interactions = void; the device runs the (semi-)smooth particles on
    the Laplace kernels with k = 0 -- same u, v, a, forces and well indices."""
    F = product()
    N = 300
    kw = dict(shape=[N], k_frame=1.0 / N, **extra, **PHYS)
    o = getattr(orc.Particles, cls)(**kw)
    p = getattr(F.Particles, cls)(kernel=kernel, **kw)
    for s in (o, p):
        s.u_frame = 30.0  # a frame far ahead: every particle crosses several wells
        s.timeSteps(400)
    assert_same_state(o, p, exact=exact)
    if cls != "System_Smooth":  # detail.h:420: "Operation not possible."
        for s in (o, p):
            assert s.minimise() == 0
            s.eventDrivenStep(1e-3, False)
            s.eventDrivenStep(1e-3, True)
            s.timeSteps(50)
        assert_same_state(o, p, exact=exact)
