"""Opt-in contracted arithmetic (``contracted=True`` = fqsb_params.kernel bit 7, FQSB_KERNEL_FMA):
the resident kernels compiled with FMA contraction. Not bit-identical by construction; what must
hold: the yield landscape is exactly the reference's, fixed-step trajectories agree with the oracle
to rounding, and the reference's goldens (avalanche sizes S) still reproduce exactly."""

import numpy as np
import pytest

from tests import protocol

pytestmark = pytest.mark.gpu


def product():
    import frictionqpotspringblock_b200 as F

    return F


CASES = [
    ("System_Cuspy_Laplace", dict(k_interactions=1.0)),
    ("System_Cuspy_Quartic", dict(a1=1.0, a2=0.7)),
    ("System_Cuspy_QuarticGradient", dict(k2=1.0, k4=0.3)),
]


@pytest.mark.parametrize("cls,extra", CASES)
@pytest.mark.parametrize("N", [1000, 4096])
def test_contracted_trajectory_agrees_with_the_oracle_to_rounding(cls, extra, N):
    from oracle import oracle as orc

    F = product()
    kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_frame=1.0 / N, dt=0.1, shape=[N],
              seed=3, distribution="random", parameters=[2.0], offset=-50, **extra)
    gpu = getattr(F.Line1d, cls)(kernel=1, contracted=True, **kw)
    exact = getattr(F.Line1d, cls)(kernel=1, **kw)
    cpu = getattr(orc.Line1d, cls)(**kw)
    for s in (gpu, exact, cpu):
        s.u_frame = 3.0
        s.timeSteps(400)
    assert gpu.last_kernel == "resident"
    # the bit-exact default is untouched by the extra build
    assert np.array_equal(exact.u, cpu.u) and np.array_equal(exact.v, cpu.v)
    # contracted: same wells, same landscape, slips / velocities equal to rounding
    assert np.array_equal(gpu.chunk.index_at_align, cpu.chunk.index_at_align)
    assert np.array_equal(gpu.chunk.left_of_align, cpu.chunk.left_of_align)
    assert np.array_equal(gpu.chunk.right_of_align, cpu.chunk.right_of_align)
    scale = np.max(np.abs(cpu.u)) + 1.0
    assert np.max(np.abs(gpu.u - cpu.u)) < 1e-12 * scale
    assert np.max(np.abs(gpu.v - cpu.v)) < 1e-12 * (np.max(np.abs(cpu.v)) + 1.0)
    assert not np.array_equal(gpu.v, cpu.v)  # (it really is a different arithmetic)


@pytest.mark.parametrize("cls,extra", CASES)
def test_contracted_blocked_kernel_agrees_with_the_oracle_to_rounding(cls, extra):
    """the temporally blocked kernel (lines beyond one CTA, config #3) has the same opt-in build"""
    from oracle import oracle as orc

    F = product()
    N = 6000
    kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_frame=1.0 / N, dt=0.1, shape=[N],
              seed=5, distribution="random", parameters=[2.0], offset=-50, **extra)
    gpu = getattr(F.Line1d, cls)(contracted=True, **kw)
    cpu = getattr(orc.Line1d, cls)(**kw)
    for s in (gpu, cpu):
        s.u_frame = 3.0
        s.timeSteps(300)
    assert gpu.last_kernel == "blocked_1d"
    assert np.array_equal(gpu.chunk.index_at_align, cpu.chunk.index_at_align)
    assert np.array_equal(gpu.chunk.right_of_align, cpu.chunk.right_of_align)
    scale = np.max(np.abs(cpu.u)) + 1.0
    assert np.max(np.abs(gpu.u - cpu.u)) < 1e-12 * scale
    assert np.max(np.abs(gpu.v - cpu.v)) < 1e-12 * (np.max(np.abs(cpu.v)) + 1.0)
    assert not np.array_equal(gpu.v, cpu.v)
    # the stop modes take the same decisions
    niter_g = gpu.minimise()
    niter_c = cpu.minimise()
    assert niter_g == niter_c
    assert np.array_equal(gpu.chunk.index_at_align, cpu.chunk.index_at_align)


@pytest.mark.parametrize("name", ["Line1d_Cuspy_Laplace", "Line1d_Cuspy_Quartic"])
def test_goldens_reproduce_with_contracted_arithmetic(name, golden_dir):
    """examples/<name>.py against its committed .h5, full length, with the FMA-contracted kernels:
    S exact at every step, frame position and force allclose."""
    F = product()
    golden = np.load(golden_dir / f"{name}.npz")
    nstep = len(golden["S"])

    class L1:  # Line1d with contracted=True injected
        def __getattr__(self, cls):
            return lambda *a, **k: getattr(F.Line1d, cls)(*a, contracted=True, **k)

    system = protocol.make(L1(), F.Line2d, name, F.Particles)
    protocol.check(golden, *protocol.run(system, nstep))
