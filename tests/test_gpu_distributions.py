"""The eight prrng distributions (detail.h:31-66) on the device against the oracle: the yield
landscape (``chunk.data``), well indices along a driven run, and -- for the advisor's point --
BACKWARD well changes: ``random`` / ``delta`` landscapes are exact in both directions (multiples
of 2^-31), the others re-associate the cumulative sum when walking left exactly like prrng's own
backward redraw does, so they are compared with allclose there (reference: tests/test_Line1d.py:
325-326 only requires allclose)."""

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

CASES = [
    ("random", [2.0, 0.5], True),
    ("delta", [1.5, 0.25], True),
    ("exponential", [2.0, 0.1], False),
    ("power", [3.0, 0.2], False),
    ("pareto", [2.0, 1.5, 0.1], False),
    ("weibull", [2.0, 1.1, 1e-3], False),
    ("normal", [2.0, 0.2, 0.0], False),
    ("gamma", [2.5, 1.3, 0.05], False),
    ("gamma", [0.6, 1.0, 0.5], False),
]


def pair(dist, par, N=257, **extra):
    import frictionqpotspringblock_b200 as F

    kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_interactions=1.0, k_frame=1.0 / N,
              dt=0.1, shape=[N], seed=13, distribution=dist, parameters=par, offset=-50.0,
              nchunk=300)
    return orc.Line1d.System_Cuspy_Laplace(**kw), F.Line1d.System_Cuspy_Laplace(**kw, **extra)


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("dist,par,exact", CASES)
def test_landscape_and_driven_run(dist, par, exact, kernel):
    o, p = pair(dist, par, kernel=kernel)
    assert np.array_equal(o.chunk.start, p.chunk.start)
    yo, yp = o.chunk.data, p.chunk.data
    if exact:
        assert np.array_equal(yo, yp)
    else:  # device log / pow / erf_inv / gamma_p_inv against glibc / the 80-bit restatements
        # (1e-13 of the scale of the landscape: the sums of 300 spacings differ by a few ulp each)
        assert np.abs(yo - yp).max() <= 1e-13 * np.abs(yo).max(), np.abs(yo - yp).max()
    i0 = o.chunk.index_at_align
    assert np.array_equal(i0, p.chunk.index_at_align)
    for s in (o, p):
        s.u_frame = 1500.0  # k_frame = 1/257: far beyond the pinning force, everything slides
        s.timeSteps(400)
    assert np.array_equal(o.chunk.index_at_align, p.chunk.index_at_align)
    assert np.sum(o.chunk.index_at_align - i0) > 3 * 257  # the blocks changed wells many times
    for name in ("u", "v", "f_potential"):
        x, y = getattr(o, name), getattr(p, name)
        if exact:
            assert np.array_equal(x, y), name
        else:
            assert np.allclose(x, y, rtol=1e-11, atol=1e-11), (name, np.abs(x - y).max())


@pytest.mark.parametrize("dist,par,exact", CASES)
def test_backward_well_changes(dist, par, exact):
    """Walk 30 wells forward, then 25 back and 10 forward again: indices exact; yield positions
    bit-identical where the landscape is exact in floating point, allclose otherwise."""
    o, p = pair(dist, par, N=64)
    for target in (30, 5, 15):
        # a position inside well `target` of every block, from the ORACLE's landscape
        st = o.chunk.start
        yo = o.chunk.data
        col = target - st
        assert np.all(col >= 0) and np.all(col + 1 < yo.shape[1])
        rows = np.arange(yo.shape[0])
        u = 0.5 * (yo[rows, col] + yo[rows, col + 1])
        for s in (o, p):
            s.u = u
        assert np.array_equal(o.chunk.index_at_align, np.full(64, target))
        assert np.array_equal(p.chunk.index_at_align, np.full(64, target))
        for name in ("left_of_align", "right_of_align"):
            x, y = getattr(o.chunk, name), getattr(p.chunk, name)
            if exact:
                assert np.array_equal(x, y), (target, name)
            else:
                assert np.allclose(x, y, rtol=1e-13, atol=1e-12), (target, name)


def test_unknown_distribution_and_too_many_parameters():
    import frictionqpotspringblock_b200 as F

    kw = dict(m=1.0, eta=0.1, mu=1.0, k_interactions=1.0, k_frame=0.1, dt=0.1, shape=[8], seed=0,
              offset=-50.0)
    with pytest.raises(RuntimeError, match="Unknown distribution: lognormal"):  # detail.h:65
        F.Line1d.System_Cuspy_Laplace(distribution="lognormal", parameters=[1.0], **kw)
    with pytest.raises(RuntimeError, match="at most 4"):
        F.Line1d.System_Cuspy_Laplace(distribution="random", parameters=[1, 0, 0, 0, 0], **kw)
