"""GPU parity of the thermal systems (External = RandomNormalForcing, SURVEY.md section 8f row
N4; /root/reference/include/FrictionQPotSpringBlock/detail.h:881-1000, Line1d.h:261-330,486-556).

The schedule (`next`), the generator state and the well indices are integers: bit-exact against
the oracle. The random forces go through erf_inv -- the oracle restates boost's long-double
evaluation, the device uses its own double-precision erf_inv_dev (a few 1e-16) -- so forces and the
trajectory are compared at 1e-12 of their scale (BASELINE.json north_star), the reference's own
golden with the reference's own np.allclose.
"""

import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests import protocol
from tests.helpers import product

pytestmark = pytest.mark.gpu

PHYS = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, dt=0.1, seed=3, distribution="random",
            parameters=[2.0], offset=-50)


def close(x, y, rtol=1e-12):
    x, y = np.asarray(x), np.asarray(y)
    scale = max(np.abs(x).max(), 1e-300)
    return np.abs(x - y).max() <= rtol * scale


def forcing(N, seed=5, period=7):
    rng = np.random.default_rng(seed)
    return dict(mean=0.01, stddev=0.05, seed_forcing=11,
                dinc_init=rng.integers(0, period, N).astype(np.int64),
                dinc=rng.integers(1, period, N).astype(np.int64))


def assert_same_thermal_state(o, p, rtol=1e-12):
    assert np.array_equal(o.external.next, p.external.next)
    assert o.external.state == p.external.state
    assert o.inc == p.inc
    assert np.array_equal(o.chunk.index_at_align, p.chunk.index_at_align)
    assert close(o.external.f_thermal, p.external.f_thermal, 1e-14)
    for name in ("u", "v", "a", "f", "f_potential", "f_frame", "f_interactions", "f_damping"):
        assert close(getattr(o, name), getattr(p, name), rtol), name


def test_reference_test_interactions():
    """tests/test_Line1d.py:566-600 on the device."""
    F = product()
    N = 10
    system = F.Line1d.System_Cuspy_Laplace_RandomForcing(
        m=1, eta=1, mu=1, k_interactions=1, k_frame=0.1, dt=1, mean=0, stddev=1, seed_forcing=0,
        dinc_init=np.ones(N, dtype=int), dinc=np.ones(N, dtype=int), shape=[N], seed=0,
        distribution="delta", parameters=[1.0], offset=-49.5, nchunk=100)
    assert system.residual < 1e-5
    draws = orc.pcg32_normal(0, 2 * N, 0, 1)  # gen = prrng.pcg32(0); gen.normal([N], 0, 1)
    system.inc += 1
    system.refresh()
    assert system.residual > 1e-5
    assert np.allclose(system.external.f_thermal, draws[:N])
    system.inc += 1
    system.refresh()
    assert system.residual > 1e-5
    assert np.allclose(system.external.f_thermal, draws[N:])
    assert close(system.external.f_thermal, draws[N:], 1e-14)
    with pytest.raises(RuntimeError, match="Minimisation not implementated"):  # detail.h:1692
        system.minimise()
    assert repr(system.external) == "<FrictionQPotSpringBlock.detail.RandomNormalForcing_1>"


@pytest.mark.parametrize("cls,extra", [
    ("Line1d.System_Cuspy_Laplace_RandomForcing", dict(k_interactions=1.0)),
    ("Line1d.System_Cuspy_Quartic_RandomForcing", dict(a1=1.0, a2=0.7)),
    ("Particles.System_Cuspy_RandomForcing", dict()),
])
@pytest.mark.parametrize("N", [7, 300, 1500, 2500])
@pytest.mark.parametrize("kernel", [0, 2], ids=["resident", "stream"])
def test_thermal_steps_match_oracle(cls, extra, N, kernel):
    """timeSteps / flowSteps with blocks redrawn on ragged schedules (several draws per step,
    blocks due at construction, N beyond one scan chunk), on the resident kernel (K2t, every
    blocks-per-thread configuration) and on the streaming path."""
    F = product()
    module, name = cls.split(".")
    kw = dict(shape=[N], k_frame=1.0 / N, **extra, **PHYS, **forcing(N))
    o = getattr(getattr(orc, module), name)(**kw)
    p = getattr(getattr(F, module), name)(kernel=kernel, **kw)
    assert_same_thermal_state(o, p)  # the draws of initSystem's refresh() (dinc_init == 0)
    for s in (o, p):
        s.u_frame = 0.7
        s.timeStep()
    assert_same_thermal_state(o, p)
    for s in (o, p):
        s.timeSteps(40)
        s.flowSteps(60, 0.3)
    assert_same_thermal_state(o, p)
    assert np.isclose(o.temperature, p.temperature, rtol=1e-12, atol=0)
    assert np.isclose(o.residual, p.residual, rtol=1e-12, atol=0)
    assert p.last_kernel == ("stream" if kernel == 2 else "resident_thermal")


@pytest.mark.parametrize("kernel", [0, 2], ids=["resident", "stream"])
def test_thermal_extreme_schedules(kernel):
    """schedule entries beyond the on-chip 31-bit window: never due, always due, huge period"""
    F = product()
    N = 96
    f = forcing(N)
    f["dinc_init"][:8] = 2**40      # never due
    f["dinc_init"][8:16] = -(2**35)  # overdue for the whole run: redrawn every step
    f["dinc"][16:24] = 2**36         # drawn once, then never again
    f["dinc_init"][16:24] = 3
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, **PHYS, **f)
    o = orc.Line1d.System_Cuspy_Laplace_RandomForcing(**kw)
    p = F.Line1d.System_Cuspy_Laplace_RandomForcing(kernel=kernel, **kw)
    for n in (2, 30):
        for s in (o, p):
            s.timeSteps(n)
        assert_same_thermal_state(o, p)


def test_thermal_resident_equals_streaming_bitwise():
    """both device paths draw through the same erf_inv_dev: identical bits, sparse schedule
    (dinc = 100 as in the reference's example) over several launches"""
    F = product()
    N = 1000
    rng = np.random.default_rng(1)
    f = dict(mean=0.0, stddev=0.05, seed_forcing=0, dinc_init=rng.integers(0, 100, N),
             dinc=100 * np.ones(N, dtype=np.int64))
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, **PHYS, **f)
    a = F.Line1d.System_Cuspy_Laplace_RandomForcing(kernel=0, **kw)
    b = F.Line1d.System_Cuspy_Laplace_RandomForcing(kernel=2, **kw)
    for n in (1, 250, 333):
        for s in (a, b):
            s.flowSteps(n, 5e-2)
        for name in ("u", "v", "a", "f"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), name
        assert np.array_equal(a.external.f_thermal, b.external.f_thermal)
        assert np.array_equal(a.external.next, b.external.next)
        assert a.external.state == b.external.state and a.inc == b.inc
    assert (a.last_kernel, b.last_kernel) == ("resident_thermal", "stream")
    # timeStepsUntilEvent is a stop mode: streaming path on both
    assert a.timeStepsUntilEvent() == b.timeStepsUntilEvent()
    assert np.array_equal(a.u, b.u)


def test_external_setters_and_set_inc():
    """external.state / next / f_thermal round trips; set_inc redraws (detail.h:1246)."""
    F = product()
    N = 64
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, **PHYS, **forcing(N))
    o = orc.Line1d.System_Cuspy_Laplace_RandomForcing(**kw)
    p = F.Line1d.System_Cuspy_Laplace_RandomForcing(**kw)
    for s in (o, p):
        s.timeSteps(10)
    state, nxt = o.external.state, o.external.next
    assert p.external.state == state and np.array_equal(p.external.next, nxt)
    for s in (o, p):
        s.timeSteps(10)
        s.external.state = state  # replay the same draws on a shifted schedule
        s.external.next = nxt + 3
        s.external.f_thermal = np.linspace(-1, 1, N)
        s.inc = s.inc + 4  # updated_inc(): every block with next <= inc is redrawn once
    assert_same_thermal_state(o, p)
    for s in (o, p):
        s.timeSteps(25)
    assert_same_thermal_state(o, p)
    with pytest.raises(RuntimeError, match="has_shape"):
        p.external.f_thermal = np.zeros(N + 1)
    with pytest.raises(AttributeError):
        F.Line1d.System_Cuspy_Laplace(shape=[N], k_frame=1.0 / N, k_interactions=1.0,
                                      **PHYS).external


def test_device_normal_draws_accuracy():
    """2.6e5 draws of the device's erf_inv (Chebyshev fits, tools/erfinv_fit.py) against the
    oracle's long-double solve: a few 1e-16 relative, tails included"""
    F = product()
    N, R = 4096, 64
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, **PHYS)
    ens = F.Line1d.Ensemble_Cuspy_Laplace_RandomForcing(
        nrealisations=R, mean=0.0, stddev=1.0, seed_forcing=123, dinc_init=np.ones(N, dtype=int),
        dinc=np.ones(N, dtype=int), **kw)
    ens.timeStep()
    got = ens.external.f_thermal
    worst = 0.0
    for r in range(R):
        ref = orc.pcg32_normal(123 + r, N, 0.0, 1.0)
        worst = max(worst, np.max(np.abs(got[r] - ref) / np.abs(ref)))
    assert worst < 2e-15, worst
    assert np.abs(got).max() > 4.0  # the sample reaches the tails


def test_thermal_ensemble_equals_independent_systems():
    """realisation r = the single system with seed + r*N and seed_forcing + r"""
    F = product()
    N, R = 200, 5
    f = forcing(N)
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, **PHYS)
    ens = F.Line1d.Ensemble_Cuspy_Laplace_RandomForcing(nrealisations=R, **kw, **f)
    ens.u_frame = np.full(R, 0.4)
    ens.flowSteps(150, 0.2)
    for r in range(R):
        one = dict(kw, seed=PHYS["seed"] + r * N)
        fr = dict(f, seed_forcing=f["seed_forcing"] + r)
        s = F.Line1d.System_Cuspy_Laplace_RandomForcing(**one, **fr)
        s.u_frame = 0.4
        s.flowSteps(150, 0.2)
        assert np.array_equal(ens.u[r], s.u) and np.array_equal(ens.v[r], s.v)
        assert np.array_equal(ens.external.f_thermal[r], s.external.f_thermal)
        assert ens.external.state[r] == s.external.state
        assert ens.inc[r] == s.inc


def test_thermal_golden_on_gpu(golden_dir):
    """examples/Line1d_System_Cuspy_Laplace_RandomForcing.py against its committed .h5
    (x_frame, f_frame, t_insta np.allclose); FQSB_FULL_GOLDEN=1 runs all 500 outputs."""
    F = product()
    golden = np.load(golden_dir / "Line1d_System_Cuspy_Laplace_RandomForcing.npz")
    full = os.environ.get("FQSB_FULL_GOLDEN", "0") == "1"
    system = protocol.make_thermal(F.Line1d, orc.pcg32_randint)
    protocol.check_thermal(golden, *protocol.run_thermal(system, 500 if full else 25))
