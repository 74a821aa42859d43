"""BASELINE.json's configs at their FULL sizes: comparisons with the CPU oracle over a fixed number
of steps / on sampled realisations (bit-exact state and well indices; 1e-12 of the force scale for
the LongRange GEMM), plus size-independent properties of the domain."""

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import product

pytestmark = pytest.mark.gpu

PHYS = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, dt=0.1, distribution="random",
            parameters=[2.0], offset=-50)


def wells_consistent(s):
    """y_left < u <= y_right for every block (prrng lower_bound convention)."""
    u, yl, yr = s.u, s.chunk.left_of_align, s.chunk.right_of_align
    return bool(np.all(yl < u) and np.all(u <= yr))


def test_config2_ensemble_16384_x_4096():
    F = product()
    N, R = 4096, 16384
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, **PHYS)
    ens = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, seed=0, **kw)
    ens.u_frame = np.full(R, 0.5)
    assert np.all(ens.minimise() == 0)
    i_n = ens.chunk.index_at_align.copy()
    inc_n = ens.inc.copy()
    ens.eventDrivenStep(1e-3, False)
    du = ens.eventDrivenStep(1e-3, True)
    assert np.all(du > 0)
    assert np.all(ens.minimise() == 0)
    assert wells_consistent(ens)
    assert np.all(ens.residual < 1e-5)
    # device-side avalanche statistics == host arithmetic on the indices
    S, A = ens.avalanche(i_n)
    idx = ens.chunk.index_at_align
    assert np.array_equal(S, np.sum(idx - i_n, axis=1))
    assert np.array_equal(A, np.sum(idx != i_n, axis=1))
    assert np.all(S >= 0) and S.max() > 0
    # idempotence: minimising an equilibrium again needs at least niter_tol and only a few steps
    # (the quenched state restarts from rest) and moves no block to another well
    inc1 = ens.inc.copy()
    assert np.all(ens.minimise() == 0)
    again = ens.inc - inc1
    assert again.min() >= 10 and again.max() < 500
    assert np.array_equal(ens.chunk.index_at_align, idx)
    # sampled realisations equal independent single systems with seed + r*N (bit for bit)
    for r in (0, 1, 8191, 16383):
        s = F.Line1d.System_Cuspy_Laplace(seed=r * N, **kw)
        s.u_frame = 0.5
        assert s.minimise() == 0
        s.eventDrivenStep(1e-3, False)
        assert s.eventDrivenStep(1e-3, True) == du[r]
        assert s.minimise() == 0
        assert s.minimise() == 0
        assert np.array_equal(s.u, ens.u[r])
        assert np.array_equal(s.chunk.index_at_align, idx[r])
        assert s.inc == ens.inc[r]
    assert len(np.unique(inc1 - inc_n)) > 100  # realisations stop at their own steps


@pytest.mark.parametrize("cls,extra", [("System_Cuspy_Quartic", dict(a1=1.0, a2=1.0)),
                                       ("System_SemiSmooth_Laplace",
                                        dict(k_interactions=1.0, kappa=1.0))])
def test_config3_line_2pow20_blocked_equals_streaming(cls, extra):
    """config #3 at full size: the temporally blocked kernel (default for lines beyond one CTA)
    and the one-step-per-launch streaming kernel give bit-identical slips and well indices; the
    number of minimisation steps may differ by the H1 effect only (different association of the
    residual sums), which would show up in `inc`."""
    F = product()
    N = 1 << 20
    kw = dict(shape=[N], k_frame=1.0 / N, seed=0, **extra, **PHYS)
    runs = []
    for kernel, name in ((0, "blocked_1d"), (2, "stream_1d"), (0, "blocked_1d")):
        s = getattr(F.Line1d, cls)(kernel=kernel, **kw)
        s.u_frame = 0.5
        ret = s.minimise()
        assert ret == 0 and s.last_kernel == name
        s.eventDrivenStep(1e-3, False)
        s.eventDrivenStep(1e-3, True)
        s.timeSteps(100)
        assert wells_consistent(s)
        runs.append((s.inc, s.u.copy(), s.chunk.index_at_align.copy(), s.v.copy()))
    for other in runs[1:]:
        assert runs[0][0] == other[0]
        assert np.array_equal(runs[0][1], other[1])
        assert np.array_equal(runs[0][2], other[2])
        assert np.array_equal(runs[0][3], other[3])


def test_config4_longrange_8192_x_1024_force_properties():
    """the DMMA Toeplitz GEMM at full size: zero net interaction force (Newton's third law),
    invariance under a uniform shift (the H5 cancellation), agreement with the exact resident sum
    on a sampled realisation."""
    F = product()
    N, R = 8192, 1024
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, alpha=1.5, **PHYS)
    ens = F.Line1d.Ensemble_Cuspy_LongRange(nrealisations=R, seed=0, **kw)
    rng = np.random.default_rng(1)
    u = rng.standard_normal((R, N)) * 2.0
    ens.u = u
    f = ens.f_interactions
    scale = np.abs(f).max()
    assert np.abs(f.sum(axis=1)).max() <= 1e-9 * scale
    ens.u_frame = np.full(R, 12345.0)
    ens.u = u + 12345.0
    f2 = ens.f_interactions
    assert np.abs(f2 - f).max() <= 1e-9 * scale
    ens.timeSteps(3)
    assert ens.last_kernel == "stream_longrange_dmma"
    assert wells_consistent(ens)


def test_config5_interface_4096_x_4096_nopassing_monotone():
    """Middleton's no-passing rule: relaxing from below, no block ever moves backwards; the fixed
    point satisfies the full force balance; Verlet minimisation of the same interface reaches an
    equilibrium too."""
    F = product()
    shape = [4096, 4096]
    n = shape[0] * shape[1]
    s = F.Line2d.System_Cuspy_Laplace_Nopassing(
        mu=1.0, k_interactions=1.0, k_frame=1.0 / n, shape=shape, seed=0,
        distribution="random", parameters=[2.0], offset=-50)
    s.u_frame = 1.0
    assert s.minimise() == 0  # (the flat initial state is not "from below": no monotonicity yet)
    assert s.residual < 1e-5
    i_n = s.chunk.index_at_align.copy()
    # after a forward event-driven step every force is >= 0: the relaxation must be monotone
    s.eventDrivenStep(1e-3, False)
    s.eventDrivenStep(1e-3, True)
    # (the starting point is an equilibrium only to the tolerance 1e-5 of minimise(), so blocks
    #  may still settle backwards by a correspondingly tiny amount; none may re-enter a well)
    slack = 1e-5
    u_prev = s.u.copy()
    for sweeps in (1, 3, 10, 30):
        s.minimise(max_iter=sweeps, max_iter_is_error=False)
        u = s.u
        assert np.all(u >= u_prev - slack)
        assert np.all(s.chunk.index_at_align >= i_n)
        u_prev = u.copy()
    assert s.minimise() == 0
    assert np.all(s.u >= u_prev - slack)
    assert np.all(s.chunk.index_at_align >= i_n)
    S, A = s.avalanche(i_n)
    assert S == np.sum(s.chunk.index_at_align - i_n) and S >= 1
    assert wells_consistent(s)
    assert s.residual < 1e-5
    assert s.last_kernel == "stream_nopassing"


# ---- oracle comparisons at the BASELINE sizes ---------------------------------------------------
def _same_state(o, p):
    assert np.array_equal(o.chunk.index_at_align, p.chunk.index_at_align)
    for name in ("u", "v", "a"):
        assert np.array_equal(getattr(o, name), getattr(p, name)), name
    assert np.array_equal(o.chunk.left_of_align, p.chunk.left_of_align)
    assert np.array_equal(o.chunk.right_of_align, p.chunk.right_of_align)
    assert o.inc == p.inc


@pytest.mark.parametrize("kernel", [0, 2])
@pytest.mark.parametrize("cls,extra", [("System_Cuspy_Quartic", dict(a1=1.0, a2=1.0)),
                                       ("System_SemiSmooth_Laplace",
                                        dict(k_interactions=1.0, kappa=1.0))])
def test_config3_line_2pow20_50_steps_equal_the_oracle(cls, extra, kernel):
    """config #3, N = 2^20: 50 velocity-Verlet steps of a driven line, temporally blocked and
    streaming kernels, bit for bit against the oracle (state, wells, indices, forces)."""
    F = product()
    N = 1 << 20
    kw = dict(shape=[N], k_frame=1.0 / N, seed=0, **extra, **PHYS)
    o = getattr(orc.Line1d, cls)(**kw)
    p = getattr(F.Line1d, cls)(kernel=kernel, **kw)
    rng = np.random.default_rng(5)
    u0 = rng.uniform(0.0, 3.0, N)  # blocks spread over several wells: plenty of hops to come
    for s in (o, p):
        s.u_frame = 40.0
        s.u = u0
        s.timeSteps(50)
    assert p.last_kernel == ("blocked_1d" if kernel == 0 else "stream_1d")
    _same_state(o, p)
    assert np.array_equal(o.f, p.f) and np.array_equal(o.f_interactions, p.f_interactions)
    assert np.array_equal(o.f_potential, p.f_potential)
    hops = int(np.sum(p.chunk.index_at_align != orc_index_at(u0, cls, kw)))
    assert hops > 1000  # the 50 steps did change wells
    assert np.isclose(o.residual, p.residual, rtol=1e-10)


def orc_index_at(u0, cls, kw):
    o = getattr(orc.Line1d, cls)(**kw)
    o.u = u0
    return o.chunk.index_at_align


def test_config5_interface_4096_steps_and_sweeps_equal_the_oracle():
    """config #5, 4096 x 4096: 20 Verlet steps (TMA-staged row-marching kernel) and 20 no-passing
    sweeps of the FULL interface on the GPU, bit for bit against the oracle. The oracle integrates
    bands of 296 rows (the full interface costs it 20 GB and minutes): a band with the generators
    of global rows r0 .. r0 + 296 (seed = r0 * cols) evolves exactly like those rows of the full
    interface except for the error entering through its periodic wrap, which moves one row per
    step / sweep -- after 20 of them rows 20 .. 276 of the band are exact."""
    F = product()
    shape = [4096, 4096]
    n = shape[0] * shape[1]
    base = dict(mu=1.0, k_interactions=1.0, k_frame=1.0 / n, distribution="random",
                parameters=[2.0], offset=-50)
    rng = np.random.default_rng(9)
    u0 = rng.uniform(0.0, 2.0, shape)
    dyn = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, dt=0.1)
    nband, k = 296, 20
    bands = (0, 2037, 4096 - nband)

    def check_bands(p, cls, steps, **extra):
        idx, u = p.chunk.index_at_align, p.u
        yl = p.chunk.left_of_align
        v = p.v if extra else None
        moved = 0
        for r0 in bands:
            o = cls(shape=[nband, shape[1]], seed=r0 * shape[1], **base, **extra)
            i0 = o.chunk.index_at_align
            o.u_frame = 30.0
            o.u = u0[r0:r0 + nband]
            steps(o)
            sl = slice(k, nband - k)
            gl = slice(r0 + k, r0 + nband - k)
            assert np.array_equal(o.chunk.index_at_align[sl], idx[gl])
            assert np.array_equal(o.u[sl], u[gl])
            assert np.array_equal(o.chunk.left_of_align[sl], yl[gl])
            if v is not None:
                assert np.array_equal(o.v[sl], v[gl])
            moved += int(np.sum(o.chunk.index_at_align[sl] != i0[sl]))
        assert moved > 1000  # the steps did change wells

    p = F.Line2d.System_Cuspy_Laplace(shape=shape, seed=0, **base, **dyn)
    p.u_frame = 30.0
    p.u = u0
    p.timeSteps(k)
    assert p.last_kernel == "stream_2d"
    check_bands(p, orc.Line2d.System_Cuspy_Laplace, lambda o: o.timeSteps(k), **dyn)
    del p

    def sweeps(s):
        assert s.minimise(tol=1e-300, max_iter=k, max_iter_is_error=False) == k + 1  # quirk Q4

    p = F.Line2d.System_Cuspy_Laplace_Nopassing(shape=shape, seed=0, **base)
    p.u_frame = 30.0
    p.u = u0
    sweeps(p)
    assert p.last_kernel == "stream_nopassing"
    check_bands(p, orc.Line2d.System_Cuspy_Laplace_Nopassing, sweeps)


def test_config2_sampled_realisations_equal_the_oracle():
    """config #2 at full size (16384 x 4096, resident kernel): three realisations through
    minimise -> eventDrivenStep -> kick -> minimise against the ORACLE: well indices, S, A, the
    step counts and the whole state bit for bit."""
    F = product()
    N, R = 4096, 16384
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, **PHYS)
    ens = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, seed=0, **kw)
    assert np.all(ens.minimise() == 0)
    ens.mark_indices()
    ens.eventDrivenStep(1e-3, False)
    du = ens.eventDrivenStep(1e-3, True)
    assert np.all(ens.minimise() == 0)
    assert ens.last_kernel == "resident"
    S, A = ens.avalanche_since_mark()
    sample = (0, 4097, 16383)
    u, v, idx, inc, uf = ens.u[sample, :], ens.v[sample, :], ens.chunk.index_at_align[sample, :], \
        ens.inc, ens.u_frame
    for k, r in enumerate(sample):
        o = orc.Line1d.System_Cuspy_Laplace(seed=r * N, **kw)
        assert o.minimise() == 0
        i_n = o.chunk.index_at_align
        o.eventDrivenStep(1e-3, False)
        assert o.eventDrivenStep(1e-3, True) == du[r]
        assert o.minimise() == 0
        assert np.array_equal(o.chunk.index_at_align, idx[k])
        assert int(np.sum(o.chunk.index_at_align - i_n)) == int(S[r])
        assert int(np.sum(o.chunk.index_at_align != i_n)) == int(A[r])
        assert o.inc == inc[r] and o.u_frame == uf[r]
        assert np.array_equal(o.u, u[k]) and np.array_equal(o.v, v[k])


def test_config4_longrange_8192_25_steps_equal_the_oracle():
    """config #4 geometry (N = 8192, alpha = 1.5) on the DMMA Toeplitz GEMM: 25 steps of a driven
    ensemble, two realisations against the oracle's exact sequential sum -- well indices exactly,
    positions / velocities / forces to 1e-12 of their scale (BASELINE north_star)."""
    F = product()
    N, R = 8192, 64
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, alpha=1.5, **PHYS)
    ens = F.Line1d.Ensemble_Cuspy_LongRange(nrealisations=R, seed=0, **kw)
    rng = np.random.default_rng(2)
    u0 = rng.uniform(0.0, 3.0, (R, N))
    ens.u_frame = np.full(R, 40.0)
    ens.u = u0
    ens.timeSteps(25)
    assert ens.last_kernel == "stream_longrange_dmma"
    sample = (0, R - 1)
    for r in sample:
        o = orc.Line1d.System_Cuspy_LongRange(seed=r * N, **kw)
        o.u_frame = 40.0
        o.u = u0[r]
        i0 = o.chunk.index_at_align
        o.timeSteps(25)
        assert np.array_equal(o.chunk.index_at_align, ens.chunk.index_at_align[r])
        assert np.sum(o.chunk.index_at_align != i0) > 100
        for name in ("u", "v", "a", "f", "f_interactions"):
            x, y = getattr(o, name), getattr(ens, name)[r]
            scale = np.abs(x).max()
            assert np.abs(x - y).max() <= 1e-12 * scale, (name, np.abs(x - y).max(), scale)
