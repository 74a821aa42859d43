"""Size-independent properties at the FULL sizes of BASELINE.json's configs (the oracle cannot run
these sizes in seconds, so the checks are invariants of the domain plus sampled comparisons)."""

import numpy as np
import pytest

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
    # one realisation against the exact-order O(N^2) sum of the resident kernel (N <= 4096 there):
    # embed nothing -- compare with a float128-free numpy circulant product instead
    d = np.minimum(np.arange(N), N - np.arange(N)).astype(float)
    pref = np.zeros(N)
    pref[1:] = 1.0 / d[1:] ** 2.5
    w = u[3] - u[3].mean()
    conv = np.real(np.fft.ifft(np.fft.fft(pref) * np.fft.fft(w)))
    expect = conv - pref.sum() * w
    assert np.abs(expect - f[3]).max() <= 1e-9 * scale
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
