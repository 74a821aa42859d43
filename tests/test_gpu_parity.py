"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeds.

Bars: well indices, S, A bit-exact; u, v, a, forces over a fixed number
of steps within 1e-12 relative -- and in fact bit-identical for every polynomial force law,
because both sides evaluate the reference's expressions in the same order without FMA.
"""

import numpy as np
import pytest

from tests import protocol
from tests.helpers import assert_same_state, pair, product

pytestmark = pytest.mark.gpu

PHYS = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, dt=0.1, seed=3, distribution="random",
            parameters=[2.0], offset=-50)

# kernel selector of the temporally blocked path (K2b): 3 | steps per launch << 8 | owned blocks
# per tile << 16 -- small tiles and short batches so that tile seams, ragged last tiles, partial
# batches and redone batches are all exercised at test sizes
BLOCKED = 3 | (6 << 8) | (101 << 16)
KERNELS = dict(argvalues=[1, 2, BLOCKED], ids=["resident", "stream", "blocked"])
KERNEL_NAME = {1: "resident", 2: "stream", BLOCKED: "blocked"}

SYSTEMS_1D = [
    ("System_Cuspy_Laplace", dict(k_interactions=1.0), True),
    ("System_Cuspy_Quartic", dict(a1=1.0, a2=0.7), True),
    ("System_Cuspy_QuarticGradient", dict(k2=1.0, k4=0.3), True),
    ("System_Cuspy_LongRange", dict(k_interactions=1.0, alpha=1.5), True),
    ("System_SemiSmooth_Laplace", dict(k_interactions=1.0, kappa=0.9), True),
    ("System_Smooth_Laplace", dict(k_interactions=1.0), False),  # sin(): 1e-12, not bit-exact
]


def kicked(o, p):
    """minimise, then kick: an avalanche is under way in both systems.

    The frame is displaced first: with u_frame = 0 the residual |f| / |f_frame| of the very first
    minimisation is a ratio of two rounding-noise-sized norms, and the step at which the StopList
    criterion fires then depends on the summation order of the norms -- the one
    thing a parallel reduction cannot reproduce from a sequential one."""
    for s in (o, p):
        s.u_frame = 0.5
        assert s.minimise() == 0
    assert o.inc == p.inc, "stop step differs: residual reduction order (SURVEY.md H1)"
    for s in (o, p):
        s.eventDrivenStep(1e-3, False)
        s.eventDrivenStep(1e-3, True)


@pytest.mark.parametrize("kernel", **KERNELS)
@pytest.mark.parametrize("cls,extra,exact", SYSTEMS_1D, ids=[s[0] for s in SYSTEMS_1D])
@pytest.mark.parametrize("N", [7, 300, 1000])
def test_fixed_steps_line1d(cls, extra, exact, N, kernel):
    if kernel == BLOCKED and cls == "System_Cuspy_LongRange":
        pytest.skip("all-to-all interaction: no temporal blocking")
    o, p = pair("Line1d", cls, shape=[N], k_frame=1.0 / N, kernel=kernel, **extra, **PHYS)
    rtol = 1e-12
    if cls == "System_Cuspy_LongRange" and kernel == 2 and N >= 256:
        exact, rtol = False, 1e-10  # tensor-core GEMM re-associates the O(N^2) sum (K7)
    assert_same_state(o, p, exact, rtol)
    if cls == "System_Smooth_Laplace":
        for s in (o, p):  # no event-driven protocol for the smooth potential (detail.h:420)
            s.u_frame = 3.0
    else:
        kicked(o, p)
    for n in (1, 2, 37, 160):
        o.timeSteps(n)
        p.timeSteps(n)
        assert_same_state(o, p, exact, rtol)
    assert p.last_kernel.startswith(KERNEL_NAME[kernel])


@pytest.mark.parametrize("kernel", [1, 2], ids=["resident", "stream"])
@pytest.mark.parametrize("cls,extra", [("System_Cuspy_Laplace", dict(k_interactions=1.0)),
                                       ("System_Cuspy_QuarticGradient", dict(k2=1.0, k4=0.3))])
@pytest.mark.parametrize("shape", [[5, 4], [50, 50], [37, 61], [70, 1030]])
def test_fixed_steps_line2d(cls, extra, shape, kernel):
    if kernel == 1 and shape[0] * shape[1] > 4096:
        pytest.skip("beyond the resident kernel")
    n = shape[0] * shape[1]
    o, p = pair("Line2d", cls, shape=shape, k_frame=1.0 / n, kernel=kernel, **extra, **PHYS)
    kicked(o, p)
    for nstep in (1, 50, 111):
        o.timeSteps(nstep)
        p.timeSteps(nstep)
        assert_same_state(o, p)


def test_large_resident_configurations():
    """every blocks-per-thread configuration of the resident kernel (N up to 4096)."""
    for N in (256, 257, 1024, 1025, 2048, 2049, 3000, 4096):
        o, p = pair("Line1d", "System_Cuspy_Laplace", shape=[N], k_frame=1.0 / N,
                    k_interactions=1.0, kernel=1, **PHYS)
        for s in (o, p):
            s.u_frame = 2.5
            s.timeSteps(64)
        assert_same_state(o, p)


@pytest.mark.parametrize("kernel", **KERNELS)
def test_minimise_and_event_driven_match_oracle(kernel):
    N = 400
    o, p = pair("Line1d", "System_Cuspy_Laplace", shape=[N], k_frame=1.0 / N, k_interactions=1.0,
                kernel=kernel, **PHYS)
    for s in (o, p):
        s.u_frame = 0.5  # well-conditioned residual for the first minimisation (see kicked())
    for step in range(30):
        i_n = o.chunk.index_at_align
        if step > 0:
            du_o = o.eventDrivenStep(1e-3, step % 2 == 0)
            du_p = p.eventDrivenStep(1e-3, step % 2 == 0)
            assert du_o == du_p
        if step % 2 == 0:
            assert o.minimise() == 0
            assert p.minimise() == 0
        # identical stopping step => identical state; a +-1 step difference of the stop test
        # (different summation order of the residual) would show up in `inc`
        assert o.inc == p.inc
        assert np.array_equal(o.chunk.index_at_align, p.chunk.index_at_align)
        assert np.array_equal(o.u, p.u)
        assert o.u_frame == p.u_frame
        S, A = p.avalanche(i_n)
        assert S == np.sum(o.chunk.index_at_align - i_n)
        assert A == np.sum(o.chunk.index_at_align != i_n)
        assert np.isclose(o.residual, p.residual, rtol=1e-9, atol=1e-300)


@pytest.mark.parametrize("kernel", **KERNELS)
def test_time_steps_until_event(kernel):
    N = 300
    o, p = pair("Line1d", "System_Cuspy_Laplace", shape=[N], k_frame=1.0 / N, k_interactions=1.0,
                kernel=kernel, **PHYS)
    kicked(o, p)
    seen = set()
    for _ in range(40):
        ro = o.timeStepsUntilEvent()
        rp = p.timeStepsUntilEvent()
        assert ro == rp
        seen.add(ro > 0)
        assert_same_state(o, p)
        if ro == 0:
            for s in (o, p):
                s.eventDrivenStep(1e-3, False)
                s.eventDrivenStep(1e-3, True)
    assert seen == {True, False}
    # max_iter reached without event: returns max_iter + 1 (quirk Q4, detail.h:1621)
    for s in (o, p):
        s.eventDrivenStep(1e-3, False)
    assert o.timeStepsUntilEvent(max_iter=3) == p.timeStepsUntilEvent(max_iter=3)


@pytest.mark.parametrize("kernel", **KERNELS)
def test_minimise_truncate_and_activity(kernel):
    N = 500
    o, p = pair("Line1d", "System_Cuspy_Laplace", shape=[N], k_frame=1.0 / N, k_interactions=1.0,
                kernel=kernel, **PHYS)
    for s in (o, p):
        s.u_frame = 0.5
        assert s.minimise() == 0
    for A_t, S_t in [(5, 0), (0, 12), (40, 100), (0, 0)]:
        for s in (o, p):
            s.eventDrivenStep(1e-3, False)
            s.eventDrivenStep(1e-3, True)
        i_n = o.chunk.index_at_align
        ro = o.minimise_truncate(i_n=i_n, A_truncate=A_t, S_truncate=S_t)
        rp = p.minimise_truncate(i_n=i_n, A_truncate=A_t, S_truncate=S_t)
        assert ro == rp
        assert o.quasistaticActivityFirst == p.quasistaticActivityFirst
        assert o.quasistaticActivityLast == p.quasistaticActivityLast
        assert_same_state(o, p)
        ro = o.minimise(time_activity=True)
        rp = p.minimise(time_activity=True)
        assert ro == rp == 0
        assert o.quasistaticActivityFirst == p.quasistaticActivityFirst
        assert o.quasistaticActivityLast == p.quasistaticActivityLast
        assert_same_state(o, p)


@pytest.mark.parametrize("N", [2048, 5000, 8192, 10001])
def test_streaming_multi_tile(N):
    """tile boundaries, ragged last tile and odd N (generic kernel) of the streaming path."""
    o, p = pair("Line1d", "System_Cuspy_Laplace", shape=[N], k_frame=1.0 / N,
                k_interactions=1.0, kernel=2, **PHYS)
    for s in (o, p):
        s.u_frame = 2.5
        s.timeSteps(33)
    assert_same_state(o, p)
    for s in (o, p):
        s.timeSteps(8)
    assert_same_state(o, p)
    ro = o.timeStepsUntilEvent()
    rp = p.timeStepsUntilEvent()
    assert ro == rp
    assert_same_state(o, p)
    assert p.last_kernel == ("stream_1d" if N % 2 == 0 else "stream")


@pytest.mark.parametrize("cls,extra", [
    ("System_Cuspy_Laplace", dict(k_interactions=1.0)),
    ("System_Cuspy_Quartic", dict(a1=1.0, a2=1.0)),
    ("System_SemiSmooth_Laplace", dict(k_interactions=1.0, kappa=1.0)),
])
@pytest.mark.parametrize("N", [4097, 9000, 20000])
def test_blocked_default_geometry(cls, extra, N):
    """lines beyond one CTA take the temporally blocked kernel by default (planner's tiles, 32
    steps per launch): fixed steps, a partial batch, the stop modes and a quasistatic cycle."""
    o, p = pair("Line1d", cls, shape=[N], k_frame=1.0 / N, **extra, **PHYS)
    for s in (o, p):
        s.u_frame = 2.5
        s.timeSteps(70)
    assert p.last_kernel == "blocked_1d"
    assert_same_state(o, p)
    ro = o.timeStepsUntilEvent()
    rp = p.timeStepsUntilEvent()
    assert ro == rp
    assert_same_state(o, p)
    assert o.minimise() == p.minimise() == 0
    assert_same_state(o, p)
    i_n = o.chunk.index_at_align
    for s in (o, p):
        s.eventDrivenStep(1e-3, False)
        s.eventDrivenStep(1e-3, True)
    assert o.minimise(time_activity=True) == p.minimise(time_activity=True) == 0
    assert o.quasistaticActivityFirst == p.quasistaticActivityFirst
    assert o.quasistaticActivityLast == p.quasistaticActivityLast
    assert_same_state(o, p)
    S, A = p.avalanche(i_n)
    assert S == np.sum(o.chunk.index_at_align - i_n)
    assert A == np.sum(o.chunk.index_at_align != i_n)
    assert p.last_kernel == "blocked_1d"


def test_blocked_ensemble_equals_streaming():
    """an ensemble of long lines: every realisation takes its own stop decision per batch."""
    F = product()
    N, R = 6000, 5
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, nrealisations=R, **PHYS)
    a = F.Line1d.Ensemble_Cuspy_Laplace(kernel=2, **kw)
    b = F.Line1d.Ensemble_Cuspy_Laplace(**kw)
    for s in (a, b):
        s.u_frame = np.full(R, 0.5)
        assert s.minimise().tolist() == [0] * R
        s.eventDrivenStep(1e-3, False)
        s.eventDrivenStep(1e-3, True)
        s.timeSteps(101)
        assert s.minimise().tolist() == [0] * R
    assert a.last_kernel == "stream_1d" and b.last_kernel == "blocked_1d"
    assert np.array_equal(a.u, b.u)
    assert np.array_equal(a.inc, b.inc)
    assert np.array_equal(a.chunk.index_at_align, b.chunk.index_at_align)


def test_streaming_ensemble_matches_resident():
    F = product()
    N, R = 2048, 6
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, nrealisations=R, **PHYS)
    a = F.Line1d.Ensemble_Cuspy_Laplace(kernel=1, **kw)
    b = F.Line1d.Ensemble_Cuspy_Laplace(kernel=2, **kw)
    for s in (a, b):
        s.u_frame = np.full(R, 0.5)
        assert s.minimise().tolist() == [0] * R
        s.eventDrivenStep(1e-3, False)
        s.eventDrivenStep(1e-3, True)
        s.timeSteps(101)
        assert s.minimise().tolist() == [0] * R
    assert np.array_equal(a.u, b.u)
    assert np.array_equal(a.inc, b.inc)
    assert np.array_equal(a.chunk.index_at_align, b.chunk.index_at_align)


@pytest.mark.parametrize("kernel", **KERNELS)
def test_flow_steps_and_temperature(kernel):
    N = 256
    o, p = pair("Line1d", "System_Cuspy_Laplace", shape=[N], k_frame=1.0, k_interactions=1.0,
                kernel=kernel, **{**PHYS, "eta": 0.1})
    for v_frame in (0.1, 1.3):
        o.flowSteps(200, v_frame)
        p.flowSteps(200, v_frame)
        assert_same_state(o, p)
        assert np.isclose(o.temperature, p.temperature, rtol=1e-12)


def test_no_convergence_and_nan_errors():
    F = product()
    N = 64
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, **PHYS)
    p = F.Line1d.System_Cuspy_Laplace(**kw)
    p.eventDrivenStep(1e-3, True)
    with pytest.raises(RuntimeError, match="No convergence found"):  # detail.h:1788
        p.minimise(max_iter=3)
    assert p.minimise(max_iter=3, max_iter_is_error=False) == 4  # quirk Q4
    with pytest.raises(RuntimeError, match="tol < 1.0"):
        p.minimise(tol=2.0)
    u = p.u
    u[3] = np.nan
    p.u = u
    with pytest.raises(RuntimeError, match="NaN entries found"):  # detail.h:1568
        p.timeStep()
    s = F.Line1d.System_Smooth_Laplace(**kw)
    with pytest.raises(RuntimeError, match="Operation not possible."):  # detail.h:420
        s.eventDrivenStep(1e-3, False)
    with pytest.raises(RuntimeError, match="has_shape"):
        p.u = np.zeros(N + 1)
    with pytest.raises(RuntimeError, match="Deprecated, use 'u'"):
        p.x


@pytest.mark.parametrize("name", [
    "Line1d_Cuspy_Laplace",
    "Line1d_Cuspy_Laplace_Nopassing",
    "Line1d_Cuspy_Quartic",
    "Line1d_SemiSmooth_Laplace",
    "Line1d_Cuspy_Laplace_LongRange",
    "Line2d_Cuspy_Laplace",
    "Particles_Cuspy",
])
def test_golden_protocol_on_gpu(name, golden_dir):
    """The reference's own regression, FULL LENGTH: examples/<name>.py against its committed .h5
    (1000 / 200 / 2000 protocol steps; S exact at every step, frame position and force allclose as
    examples/Line1d_Cuspy_Laplace.py:64-67 asserts). FQSB_GOLDEN_PREFIX=n shortens the runs."""
    import os

    F = product()
    golden = np.load(golden_dir / f"{name}.npz")
    nstep = len(golden["S"])
    prefix = int(os.environ.get("FQSB_GOLDEN_PREFIX", "0"))
    if prefix > 0:
        nstep = min(nstep, prefix)
    system = protocol.make(F.Line1d, F.Line2d, name, F.Particles)
    protocol.check(golden, *protocol.run(system, nstep))


def test_nopassing_matches_oracle_1d_and_2d():
    base = dict(mu=1.0, k_interactions=1.0, seed=11, distribution="random", parameters=[2.0],
                offset=-50)
    for module, shape, kernel in [("Line1d", [333], 1), ("Line1d", [333], 2),
                                  ("Line2d", [24, 31], 1), ("Line2d", [24, 31], 2)]:
        n = int(np.prod(shape))
        o, p = pair(module, "System_Cuspy_Laplace_Nopassing", shape=shape, k_frame=1.0 / n,
                    kernel=kernel, **base)
        for s in (o, p):
            s.u_frame = 0.5
        for step in range(60):
            if step > 0:
                assert o.eventDrivenStep(1e-3, step % 2 == 0) == \
                    p.eventDrivenStep(1e-3, step % 2 == 0)
            if step % 2 == 0:
                assert o.minimise() == p.minimise() == 0
            assert np.array_equal(o.chunk.index_at_align, p.chunk.index_at_align)
            assert np.array_equal(o.u, p.u)
        assert np.allclose(o.f, p.f, rtol=1e-12, atol=1e-15)
        # fixed point of the full force balance
        assert p.residual < 1e-5


def test_ensemble_equals_independent_systems():
    F = product()
    N, R = 128, 5
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, **PHYS)
    ens = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, **kw)
    singles = [F.Line1d.System_Cuspy_Laplace(**{**kw, "seed": PHYS["seed"] + r * N})
               for r in range(R)]
    assert ens.minimise().tolist() == [0] * R
    for s in singles:
        assert s.minimise() == 0
    ens.eventDrivenStep(1e-3, False)
    du = ens.eventDrivenStep(1e-3, True)
    for r, s in enumerate(singles):
        s.eventDrivenStep(1e-3, False)
        assert s.eventDrivenStep(1e-3, True) == du[r]
    ens.timeSteps(77)
    ret = ens.minimise()
    for r, s in enumerate(singles):
        s.timeSteps(77)
        assert s.minimise() == ret[r] == 0
        assert np.array_equal(ens.u[r], s.u)
        assert np.array_equal(ens.chunk.index_at_align[r], s.chunk.index_at_align)
        assert ens.inc[r] == s.inc
        assert ens.u_frame[r] == s.u_frame
    assert len(set(ens.inc.tolist())) > 1  # realisations stop at their own step


@pytest.mark.parametrize("N,alpha", [(512, 1.0), (1000, 1.5), (4400, 0.5)])
def test_longrange_tensor_core_gemm_matches_oracle(N, alpha):
    """K7: LongRange through the DMMA Toeplitz GEMM (streaming path) against the oracle's exact
    O(N^2) sum: forces within 1e-12 of the force scale, well indices exact over a fixed number of
    steps (the GEMM re-associates the sum, so positions agree to rounding, not bit for bit)."""
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, alpha=alpha, **PHYS)
    o, p = pair("Line1d", "System_Cuspy_LongRange", kernel=2, **kw)
    rng = np.random.default_rng(5)
    u = 3.0 * rng.standard_normal(N) + 20000.0  # large offset: exercises the cancellation (H5)
    for s in (o, p):
        s.u_frame = 20001.0
        s.u = u
    scale = np.abs(o.f_interactions).max()
    assert np.abs(o.f_interactions - p.f_interactions).max() <= 1e-12 * scale
    assert np.isclose(o.residual, p.residual, rtol=1e-10)
    for s in (o, p):
        s.timeSteps(25)
    assert p.last_kernel == "stream_longrange_dmma"
    assert np.array_equal(o.chunk.index_at_align, p.chunk.index_at_align)
    assert np.abs(o.u - p.u).max() <= 1e-10
    assert np.abs(o.v - p.v).max() <= 1e-10
    ro = o.minimise()
    rp = p.minimise()
    assert ro == rp == 0
    assert abs(o.inc - p.inc) <= 1
    assert np.array_equal(o.chunk.index_at_align, p.chunk.index_at_align)


def test_longrange_gemm_ensemble_matches_resident():
    F = product()
    N, R = 1024, 70  # ragged realisation tile (64 + 6)
    kw = dict(shape=[N], k_frame=1.0 / N, k_interactions=1.0, alpha=1.5, nrealisations=R, **PHYS)
    a = F.Line1d.Ensemble_Cuspy_LongRange(kernel=1, **kw)  # exact O(N^2) sum, resident
    b = F.Line1d.Ensemble_Cuspy_LongRange(**kw)  # auto: tensor-core GEMM for an ensemble
    for s in (a, b):
        s.u_frame = np.full(R, 1.5)
        s.timeSteps(40)
    assert a.last_kernel == "resident" and b.last_kernel == "stream_longrange_dmma"
    assert np.array_equal(a.chunk.index_at_align, b.chunk.index_at_align)
    assert np.abs(a.u - b.u).max() <= 1e-11
    scale = np.abs(a.f_interactions).max()
    assert np.abs(a.f_interactions - b.f_interactions).max() <= 1e-11 * scale


@pytest.mark.parametrize("kernel", [1, 2], ids=["resident", "stream"])
@pytest.mark.parametrize("N", [2, 3, 4, 32, 33])
def test_tiny_lines(N, kernel):
    """smallest periodic lines (N = 2: both neighbours are the same block) and warp-edge sizes."""
    o, p = pair("Line1d", "System_Cuspy_Laplace", shape=[N], k_frame=0.1, k_interactions=1.0,
                kernel=kernel, **PHYS)
    for s in (o, p):
        s.u_frame = 3.0
        s.timeSteps(50)
    assert_same_state(o, p)
    for s in (o, p):
        assert s.minimise() == 0
    assert o.inc == p.inc
    assert np.array_equal(o.chunk.index_at_align, p.chunk.index_at_align)


@pytest.mark.parametrize("kernel", [1, 2], ids=["resident", "stream"])
def test_backward_driving_and_negative_direction(kernel):
    """direction = -1 (detail.h:1949-1959): wells are regenerated backwards (inverse LCG)."""
    N = 128
    o, p = pair("Line1d", "System_Cuspy_Laplace", shape=[N], k_frame=1.0 / N, k_interactions=1.0,
                kernel=kernel, **PHYS)
    for s in (o, p):
        s.u_frame = 0.5
        assert s.minimise() == 0
    for step in range(12):
        i_n = o.chunk.index_at_align
        for s in (o, p):
            s.eventDrivenStep(1e-3, False, direction=-1)
            s.eventDrivenStep(1e-3, True, direction=-1)
            assert s.minimise() == 0
        assert o.inc == p.inc
        assert np.array_equal(o.chunk.index_at_align, p.chunk.index_at_align)
        assert np.array_equal(o.u, p.u)
        assert o.u_frame == p.u_frame
    assert np.sum(o.chunk.index_at_align - i_n) <= 0


def test_landscape_underflow_is_reported():
    """u below the first yield position: the reference's chunk cannot go there either."""
    F = product()
    kw = dict(shape=[16], k_frame=0.1, k_interactions=1.0, **{**PHYS, "offset": 5.0})
    with pytest.raises(RuntimeError, match="lower the offset"):
        F.Line1d.System_Cuspy_Laplace(**kw)
    p = F.Line1d.System_Cuspy_Laplace(shape=[16], k_frame=0.1, k_interactions=1.0, **PHYS)
    with pytest.raises(RuntimeError, match="yield landscape exhausted"):
        p.u = np.full(16, -1000.0)


def test_shape_and_argument_validation():
    F = product()
    kw = dict(k_frame=0.1, k_interactions=1.0, **PHYS)
    p = F.Line1d.System_Cuspy_Laplace(shape=[8], **kw)
    with pytest.raises(RuntimeError, match="has_shape"):
        p.v = np.zeros(7)
    with pytest.raises(RuntimeError, match="has_shape"):
        p.minimise_truncate(i_n=np.zeros(9, dtype=int))
    with pytest.raises(RuntimeError, match="direction == 1"):
        p.eventDrivenStep(1e-3, True, direction=2)
    with pytest.raises(RuntimeError, match="niter_tol"):
        p.minimise(niter_tol=0)
    ens = F.Line1d.Ensemble_Cuspy_Laplace(shape=[8], nrealisations=3, **kw)
    with pytest.raises(RuntimeError, match="niter_tol > 32 is available for single systems"):
        ens.minimise(niter_tol=33)
    with pytest.raises(TypeError):
        F.Line1d.System_Cuspy_Laplace(shape=[8], seed=0)
    with pytest.raises(AttributeError):
        F.Line1d.System_Cuspy_Laplace_Nopassing(
            mu=1.0, k_interactions=1.0, k_frame=0.1, shape=[8], seed=0, distribution="random",
            parameters=[2.0], offset=-50).timeStep()
    p.t = 12.3
    assert p.inc == 123 and np.isclose(p.t, 12.3)
    p.inc = 7
    assert p.quasistaticActivityFirst == 7 and p.quasistaticActivityLast == 7


@pytest.mark.parametrize("module,cls,extra", [
    ("Line1d", "System_Cuspy_Laplace", dict(k_interactions=1.0)),
    ("Line1d", "System_Cuspy_Quartic", dict(a1=1.0, a2=0.5)),
    ("Line1d", "System_Cuspy_Laplace_Nopassing", dict(k_interactions=1.0)),
    ("Line2d", "System_Cuspy_Laplace", dict(k_interactions=1.0)),
])
@pytest.mark.parametrize("niter_tol", [33, 100])
def test_stoplists_longer_than_the_device_ring(module, cls, extra, niter_tol):
    """The reference's StopList takes any length (detail.h:1676-1689); beyond the 32 entries the
    device keeps in the lanes of a warp the decisions are replayed on the host over logged
    batches. Step counts, well indices, S/A and the activity timestamps equal the oracle's."""
    shape = [24, 20] if module == "Line2d" else [300]
    n = int(np.prod(shape))
    kw = dict(mu=1.0, k_frame=1.0 / n, shape=shape, seed=5, distribution="random",
              parameters=[2.0], offset=-50, **extra)
    nopassing = "Nopassing" in cls
    if not nopassing:
        kw.update(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, dt=0.1)
    o, p = pair(module, cls, **kw)
    for s in (o, p):
        s.u_frame = 0.7
        assert s.minimise(niter_tol=niter_tol) == 0
    assert o.inc == p.inc
    assert np.array_equal(o.chunk.index_at_align, p.chunk.index_at_align)
    for _ in range(3):
        for s in (o, p):
            s.eventDrivenStep(1e-3, False)
            s.eventDrivenStep(1e-3, True)
            assert s.minimise(niter_tol=niter_tol, time_activity=not nopassing) == 0
        assert o.inc == p.inc
        assert np.array_equal(o.chunk.index_at_align, p.chunk.index_at_align)
        assert o.quasistaticActivityFirst == p.quasistaticActivityFirst
        assert o.quasistaticActivityLast == p.quasistaticActivityLast
        assert np.allclose(o.u, p.u, rtol=0, atol=1e-9)
    assert np.all(p.v == 0.0)  # quench() on convergence
    if nopassing:
        return
    # not converging is reported the same way (quirk Q4) and timeStepsUntilEvent agrees
    for s in (o, p):
        s.eventDrivenStep(1e-3, False)
        s.eventDrivenStep(1e-3, True)
    assert o.minimise(niter_tol=niter_tol, max_iter=70, max_iter_is_error=False) == \
        p.minimise(niter_tol=niter_tol, max_iter=70, max_iter_is_error=False) == 71
    assert o.timeStepsUntilEvent(niter_tol=niter_tol) == p.timeStepsUntilEvent(niter_tol=niter_tol)
    assert o.inc == p.inc
    assert np.array_equal(o.u, p.u)
