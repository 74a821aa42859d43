"""The reference's own unit tests (/root/reference/tests/test_Line1d.py, test_Line2d.py) run
against the CUDA product. Only the imports changed: ``FrictionQPotSpringBlock`` is the B200
package, and the two python-prrng helpers the reference uses (``pcg32_array(...).weibull`` and
``lower_bound``, test_Line1d.py:283-285,320) come from the test-side oracle. Thermal
(RandomForcing) tests are out of scope."""

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

np.random.seed(20261017)


@pytest.fixture(scope="module")
def FrictionQPotSpringBlock():
    import frictionqpotspringblock_b200

    return frictionqpotspringblock_b200


def _uniform_par(N=5):
    eta, mu, k_interactions, k_frame = (float(np.random.random(1)[0]) for _ in range(4))
    return dict(m=1, eta=eta, mu=mu, k_interactions=k_interactions, k_frame=k_frame, dt=1.0,
                shape=[N], seed=0, distribution="delta", parameters=[1.0], offset=-49.5,
                nchunk=100)


def test_uniform_init(FrictionQPotSpringBlock):
    """test_Line1d.py:25-86"""
    par = _uniform_par()
    M = FrictionQPotSpringBlock.Line1d
    systems = [
        M.System_Cuspy_Laplace(**par),
        M.System_SemiSmooth_Laplace(kappa=1, **par),
        M.System_Smooth_Laplace(**par),
        M.System_Cuspy_LongRange(alpha=1, **par),
    ]
    par.pop("m")
    systems += [M.System_Cuspy_Laplace_Nopassing(**par)]
    par["m"] = 1
    par.pop("k_interactions")
    systems += [M.System_Cuspy_Quartic(a1=1, a2=1, **par)]
    systems += [M.System_Cuspy_QuarticGradient(k2=1, k4=1, **par)]
    for system in systems:
        assert system.residual < 1e-5
        assert np.allclose(system.f, 0.0)
        assert np.allclose(system.f_potential, 0.0)
        assert np.allclose(system.f_frame, 0.0)
        assert np.allclose(system.f_interactions, 0.0)
        assert np.allclose(system.f_damping, 0.0)
        assert np.all(system.chunk.index_at_align + 1 == np.argmax(system.chunk.data[0, :] > 0))
        assert np.all(system.chunk.right_of_align > 0)
        assert np.all(system.chunk.left_of_align <= 0)


def test_forces(FrictionQPotSpringBlock):
    """test_Line1d.py:89-176"""
    N = 5
    par = _uniform_par(N)
    eta, mu, k_interactions, k_frame = par["eta"], par["mu"], par["k_interactions"], par["k_frame"]
    system = FrictionQPotSpringBlock.Line1d.System_Cuspy_Laplace(**par)
    assert system.residual < 1e-5

    du = np.zeros(N)
    dv = np.zeros(N)
    du[0] = float(np.random.random(1)[0])
    dv[2] = float(np.random.random(1)[0])
    system.u += du
    system.v += dv
    umin = np.floor(du[0] + 0.5)

    f_potential = mu * np.array([umin - du[0], 0, 0, 0, 0])
    f_interactions = k_interactions * np.array([-2 * du[0], du[0], 0, 0, du[0]])
    f_frame = k_frame * np.array([-du[0], 0, 0, 0, 0])
    f_damping = eta * np.array([0, 0, -dv[2], 0, 0])

    assert np.all(system.chunk.right_of_align > system.u)
    assert np.all(system.chunk.left_of_align <= system.u)
    assert np.allclose(system.f_potential, f_potential)
    assert np.allclose(system.f_frame, f_frame)
    assert np.allclose(system.f_interactions, f_interactions)
    assert np.allclose(system.f_damping, f_damping)
    assert np.allclose(system.f, f_potential + f_frame + f_interactions + f_damping)

    du = np.zeros(N)
    dv = np.zeros(N)
    du[1] = 2.0 * float(np.random.random(1)[0])
    dv[3] = 2.0 * float(np.random.random(1)[0])
    system.u += du
    system.v += dv
    u = system.u
    v = system.v

    f_potential = mu * np.array([np.floor(u[0] + 0.5) - u[0], np.floor(u[1] + 0.5) - u[1], 0, 0, 0])
    f_interactions = k_interactions * np.array([
        u[-1] - 2 * u[0] + u[1],
        u[0] - 2 * u[1] + u[2],
        u[1] - 2 * u[2] + u[3],
        0,
        u[-2] - 2 * u[-1] + u[0],
    ])
    f_frame = k_frame * np.array([-u[0], -u[1], 0, 0, 0])
    f_damping = eta * np.array([0, 0, -v[2], -v[3], 0])

    assert np.all(system.chunk.right_of_align > system.u)
    assert np.all(system.chunk.left_of_align <= system.u)
    assert np.allclose(system.f_potential, f_potential)
    assert np.allclose(system.f_frame, f_frame)
    assert np.allclose(system.f_interactions, f_interactions)
    assert np.allclose(system.f_damping, f_damping)
    assert np.allclose(system.f, f_potential + f_frame + f_interactions + f_damping)


def _small(**extra):
    return dict(m=1.0, eta=1.0, mu=1.0, k_frame=0.1, dt=1.0, shape=[3], seed=0,
                distribution="delta", parameters=[1.0], offset=-49.5, nchunk=100, **extra)


def _check_event_driven(system):
    """test_Line1d.py:178-219, 520-563"""
    N = 3
    assert system.residual < 1e-5
    for k, (kick, pos) in enumerate([(False, 0.5 - 0.1), (True, 0.5 + 0.1), (False, 1.5 - 0.1),
                                     (True, 1.5 + 0.1)]):
        i_n = system.chunk.index_at_align
        system.eventDrivenStep(0.2, kick)
        if k == 0:
            assert system.residual < 1e-5
        assert np.allclose(system.u, pos * np.ones(N))
        if kick:
            assert not np.all(system.chunk.index_at_align == i_n)
        else:
            assert np.all(system.chunk.index_at_align == i_n)
        assert system.u_frame == pytest.approx(pos * (1.0 + 0.1) / 0.1, abs=1e-7)


def test_eventDrivenStep(FrictionQPotSpringBlock):
    _check_event_driven(
        FrictionQPotSpringBlock.Line1d.System_Cuspy_Laplace(**_small(k_interactions=1.0)))


def test_longrange_eventDrivenStep(FrictionQPotSpringBlock):
    _check_event_driven(
        FrictionQPotSpringBlock.Line1d.System_Cuspy_LongRange(**_small(k_interactions=1.0, alpha=1)))


def test_trigger(FrictionQPotSpringBlock):
    """test_Line1d.py:221-248"""
    N = 3
    system = FrictionQPotSpringBlock.Line1d.System_Cuspy_Laplace(**_small(k_interactions=1.0))
    i_n = system.chunk.index_at_align.copy()
    system.trigger(0, 0.2)
    u = np.zeros(N)
    u[0] = 0.5 + 0.1
    assert np.allclose(system.u, u)
    ret = system.minimise_truncate(i_n=i_n, A_truncate=1)
    assert np.sum(system.chunk.index_at_align != i_n) >= 1
    assert ret > 0
    assert system.residual > 1e-5


def test_trigger_leaves_forces_stale(FrictionQPotSpringBlock):
    """quirk Q1 (detail.h:1975-1976): trigger() returns before updated_u()."""
    system = FrictionQPotSpringBlock.Line1d.System_Cuspy_Laplace(**_small(k_interactions=1.0))
    o = orc.Line1d.System_Cuspy_Laplace(**_small(k_interactions=1.0))
    for s in (system, o):
        s.trigger(0, 0.2)
    assert np.array_equal(system.f, o.f)
    assert np.array_equal(system.f_potential, o.f_potential)
    assert np.array_equal(system.chunk.index_at_align, o.chunk.index_at_align)
    for s in (system, o):
        s.refresh()
    assert np.array_equal(system.f, o.f)
    assert np.array_equal(system.chunk.index_at_align, o.chunk.index_at_align)


def test_advanceToFixedForce(FrictionQPotSpringBlock):
    """test_Line1d.py:250-275"""
    system = FrictionQPotSpringBlock.Line1d.System_Cuspy_Laplace(**_small(k_interactions=1.0))
    assert system.residual < 1e-5
    system.advanceToFixedForce(0.1)
    assert np.mean(system.f_frame) == pytest.approx(0.1, abs=1e-7)
    assert system.residual < 1e-5
    system.advanceToFixedForce(0.0)
    assert np.mean(system.f_frame) == pytest.approx(0.0, abs=1e-7)
    assert np.allclose(system.u, 0.0)
    assert np.allclose(system.u_frame, 0.0)


def test_chunked(FrictionQPotSpringBlock):
    """test_Line1d.py:277-341: the regenerated landscape against a full cumsum, forward,
    backward and after restore()."""
    N = 3
    seed = 1697500000
    initstate = seed + np.arange(N)
    init_offset = 50.0

    r = np.stack([orc.pcg32_draws(int(i), 20000) for i in initstate])
    yref = np.cumsum(1e-3 + 1.1 * (-np.log(1.0 - r)) ** (1.0 / 2.0), axis=1) - init_offset

    mu = float(np.random.random(1)[0])
    system = FrictionQPotSpringBlock.Line1d.System_Cuspy_Laplace(
        m=1.0, eta=1.0, mu=mu, k_interactions=1.0, k_frame=0.1, dt=1.0, shape=[N], seed=seed,
        distribution="weibull", parameters=[2.0, 1.1, 1e-3], offset=-init_offset, nchunk=100)

    du = 10.0 * np.ones(N)
    du[0] = 5.0
    du[1] = 7.0

    u = np.copy(system.u)
    start = np.copy(system.chunk.start)
    state = np.copy(system.chunk.state_at(start))
    value = np.copy(system.chunk.data[..., 0])

    for repeat in range(3):
        if repeat >= 1:
            system.chunk.restore(state=state, value=value, index=start)
            system.u = u

        for i in list(range(0, 1500, 7)) + list(range(0, 1500, 11))[::-1]:
            system.u = i * du
            j = orc.lower_bound(yref, system.u)
            rr = np.arange(N)
            assert np.all(system.chunk.index_at_align == j)
            assert np.allclose(yref[rr, j], system.chunk.left_of_align)
            assert np.allclose(yref[rr, j + 1], system.chunk.right_of_align)
            umin = 0.5 * (yref[rr, j] + yref[rr, j + 1])
            assert np.allclose(mu * (umin - system.u), system.f_potential)

            if (repeat == 0 and i == 497) or (repeat == 1 and i == 1001):
                u = np.copy(system.u)
                start = np.copy(system.chunk.start)
                state = np.copy(system.chunk.state_at(start))
                value = np.copy(system.chunk.data[..., 0])


def test_semismooth_eventDrivenStep(FrictionQPotSpringBlock):
    """test_Line1d.py:345-399"""
    mu, kappa = 1, 0.1
    system = FrictionQPotSpringBlock.Line1d.System_SemiSmooth_Laplace(
        **{**_small(k_interactions=1.0, kappa=kappa), "mu": mu})
    assert system.residual < 1e-5
    u0 = system.u.copy()
    uf0 = system.u_frame
    left = system.chunk.left_of_align
    right = system.chunk.right_of_align
    mid = 0.5 * (left + right)
    upper = (mu * mid + kappa * right) / (mu + kappa)
    lower = (mu * mid + kappa * left) / (mu + kappa)
    eps = 0.001

    assert system.maxUniformDisplacement() == pytest.approx(np.min(upper - system.u), abs=1e-7)
    system.eventDrivenStep(eps=eps, kick=False)
    assert system.residual < 1e-5
    assert np.allclose(system.u, upper - 0.5 * eps)
    assert system.maxUniformDisplacement() == pytest.approx(0.5 * eps, abs=1e-7)

    system.eventDrivenStep(eps=eps, kick=True)
    assert system.residual > 1e-5
    assert np.allclose(system.u, upper + 0.5 * eps)
    assert system.maxUniformDisplacement() == pytest.approx(0, abs=1e-7)

    system.u = u0
    system.u_frame = uf0

    assert system.maxUniformDisplacement(-1) == pytest.approx(np.min(system.u - lower), abs=1e-7)
    system.eventDrivenStep(eps=eps, kick=False, direction=-1)
    assert system.residual < 1e-5
    assert np.allclose(system.u, lower + 0.5 * eps)
    assert system.maxUniformDisplacement(-1) == pytest.approx(0.5 * eps, abs=1e-7)

    system.eventDrivenStep(eps=eps, kick=True, direction=-1)
    assert system.residual > 1e-5
    assert np.allclose(system.u, lower - 0.5 * eps)
    assert system.maxUniformDisplacement() == pytest.approx(0, abs=1e-7)


def test_quartic_interactions(FrictionQPotSpringBlock):
    """test_Line1d.py:402-438"""
    N = 10
    a1 = float(np.random.random(1)[0])
    a2 = float(np.random.random(1)[0])
    system = FrictionQPotSpringBlock.Line1d.System_Cuspy_Quartic(
        m=1, eta=1, mu=1, a1=a1, a2=a2, k_frame=0.1, dt=1, shape=[N], seed=0,
        distribution="delta", parameters=[1.0], offset=-49.5, nchunk=100)
    assert system.residual < 1e-5
    du = float(np.random.random(1)[0])
    u0 = np.array([du, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    laplace = np.array([-2 * du, du, 0, 0, 0, 0, 0, 0, 0, du])
    du_p = np.array([-du, 0, 0, 0, 0, 0, 0, 0, 0, du])
    du_n = np.array([-du, du, 0, 0, 0, 0, 0, 0, 0, 0])
    f0 = a1 * laplace + a2 * (du_p**3 + du_n**3)
    for i in range(N):
        u = np.roll(u0, i)
        system.u = u
        assert np.allclose(system.f_interactions, np.roll(f0, i))
        assert np.allclose(system.u, u)


def test_quarticgradient_interactions(FrictionQPotSpringBlock):
    """test_Line1d.py:441-476"""
    N = 10
    k2 = float(np.random.random(1)[0])
    k4 = float(np.random.random(1)[0])
    du = float(np.random.random(1)[0])
    system = FrictionQPotSpringBlock.Line1d.System_Cuspy_QuarticGradient(
        m=1, eta=1, mu=1, k2=k2, k4=k4, k_frame=0.1, dt=1, shape=[N], seed=0,
        distribution="delta", parameters=[1.0], offset=-49.5, nchunk=100)
    assert system.residual < 1e-5
    u0 = np.array([du, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    laplace = np.array([-2 * du, du, 0, 0, 0, 0, 0, 0, 0, du])
    gradient = np.array([0, 0.5 * du, 0, 0, 0, 0, 0, 0, 0, 0.5 * du])
    f0 = k2 * laplace + k4 * laplace * gradient**2
    for i in range(N):
        u = np.roll(u0, i)
        system.u = u
        assert np.allclose(system.f_interactions, np.roll(f0, i))
        assert np.allclose(system.u, u)


def test_longrange_interactions(FrictionQPotSpringBlock):
    """test_Line1d.py:479-518"""
    N = 10
    k_interactions = 0.12
    system = FrictionQPotSpringBlock.Line1d.System_Cuspy_LongRange(
        m=1, eta=1, mu=1, k_interactions=k_interactions, k_frame=0.1, dt=1, alpha=1, shape=[N],
        seed=0, distribution="delta", parameters=[1.0], offset=-49.5, nchunk=100)
    dp = np.arange(N)
    dn = np.arange(N)[::-1] + 1
    d = np.where(dp < dn, dp, dn)
    x = np.zeros_like(system.u)
    x[0] = 1
    system.u = x
    f = np.zeros_like(x)
    for j in range(1, N):
        f[j] = k_interactions * (x[0] - x[j]) / (d[j] ** 2)
    f[0] = -np.sum(f)
    for i in range(N):
        system.u = np.roll(x, i)
        assert np.allclose(np.roll(f, i), system.f_interactions)


def test_line2d_laplace_interactions(FrictionQPotSpringBlock):
    """test_Line2d.py:21-57"""
    rows, cols = 5, 4
    k_interactions = float(np.random.random(1)[0])
    system = FrictionQPotSpringBlock.Line2d.System_Cuspy_Laplace(
        m=1, eta=1, mu=1, k_interactions=k_interactions, k_frame=0.1, dt=1, shape=[rows, cols],
        seed=0, distribution="delta", parameters=[1.0], offset=-49.5, nchunk=100)
    assert system.residual < 1e-5
    assert list(system.shape) == [rows, cols]
    assert system.size == rows * cols
    c = -4
    f0 = np.array([[c, 1, 0, 1], [1, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [1, 0, 0, 0]]) \
        * k_interactions
    u0 = np.array([[1, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]])
    for i in range(rows):
        for j in range(cols):
            x = np.roll(np.roll(u0, i, axis=0), j, axis=1)
            system.u = x
            f = np.roll(np.roll(f0, i, axis=0), j, axis=1)
            assert np.allclose(system.f_interactions, f)
            assert np.allclose(system.u, x)


def test_line2d_quarticgradient_basic(FrictionQPotSpringBlock):
    """test_Line2d.py:60-95"""
    rows, cols = 5, 4
    k2, k4 = 0.12, 0.0
    system = FrictionQPotSpringBlock.Line2d.System_Cuspy_QuarticGradient(
        m=1, eta=1, mu=1, k2=k2, k4=k4, k_frame=0.1, dt=1, shape=[rows, cols], seed=0,
        distribution="delta", parameters=[1.0], offset=-49.5, nchunk=100)
    assert system.residual < 1e-5
    c = -4
    f0 = np.array([[c, 1, 0, 1], [1, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [1, 0, 0, 0]]) * k2
    u0 = np.array([[1, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]])
    for i in range(rows):
        for j in range(cols):
            x = np.roll(np.roll(u0, i, axis=0), j, axis=1)
            system.u = x
            f = np.roll(np.roll(f0, i, axis=0), j, axis=1)
            assert np.allclose(system.f_interactions, f)
            assert np.allclose(system.u, x)


def test_line2d_quarticgradient(FrictionQPotSpringBlock):
    """test_Line2d.py:97-173"""
    rows = cols = 5
    k2, k4 = 0.12, 0.34
    system = FrictionQPotSpringBlock.Line2d.System_Cuspy_QuarticGradient(
        m=1, eta=1, mu=1, k2=k2, k4=k4, k_frame=0.1, dt=1, shape=[rows, cols], seed=0,
        distribution="delta", parameters=[1.0], offset=-49.5, nchunk=100)
    assert system.residual < 1e-5
    u0 = np.zeros((5, 5))
    u0[2, 2] = 1
    d2udx2 = np.zeros((5, 5))
    d2udx2[2, 1:4] = [1, -2, 1]
    d2udy2 = d2udx2.T
    dudx = np.zeros((5, 5))
    dudx[2, 1] = 0.5
    dudx[2, 3] = -0.5
    dudy = dudx.T
    d2udxdy = np.zeros((5, 5))
    f0 = (d2udx2 + d2udy2) * (k2 + k4 / 3) + 2 / 3 * k4 * (
        dudx**2 * d2udx2 + dudy**2 * d2udy2 + 2 * dudx * dudy * d2udxdy)
    for i in range(rows):
        for j in range(cols):
            x = np.roll(np.roll(u0, i, axis=0), j, axis=1)
            system.u = x
            f = np.roll(np.roll(f0, i, axis=0), j, axis=1)
            assert np.allclose(system.f_interactions, f)
