"""Host-facing paths around the stepping kernels: the fused, internally pipelined
``fqsb_run_from_host`` (state in -> timeSteps -> state out as one call) must equal the separate
public calls bit for bit, and the device-side avalanche bookkeeping (SURVEY.md section 8f row N1:
``mark_indices`` / ``avalanche_since_mark`` / ``event_record``) must equal the host arithmetic of
examples/Line1d_Cuspy_Laplace.py:59-61 and the oracle."""

import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def physics(N):
    return dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_interactions=1.0, k_frame=1.0 / N,
                dt=0.1, shape=[N], distribution="random", parameters=[2.0], offset=-50)


@pytest.mark.parametrize("R,N,chunk", [(40, 512, "7"), (333, 256, ""), (9, 1000, "4")])
def test_run_from_host_equals_the_separate_calls(R, N, chunk, monkeypatch):
    import frictionqpotspringblock_b200 as F

    if chunk:
        monkeypatch.setenv("FQSB_PIPE_CHUNK", chunk)  # several ragged chunks per stream
    kw = physics(N)
    a = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, seed=3, **kw)
    b = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, seed=3, **kw)
    rng = np.random.default_rng(0)
    u = rng.uniform(-1, 30, size=(R, N))
    v = rng.normal(size=(R, N))
    acc = rng.normal(size=(R, N))
    uf = rng.uniform(0, 30, size=R)
    a.u_frame = uf
    b.u_frame = uf
    # separate calls
    a.u, a.v, a.a = u, v, acc
    a.timeSteps(57)
    # fused call
    out_u, out_v, out_a = (np.empty((R, N)) for _ in range(3))
    mean = np.empty(R)
    b.run_from_host(57, u=u, v=v, a=acc, out_u=out_u, out_v=out_v, out_a=out_a, mean_f_frame=mean)
    assert np.array_equal(out_u, a.u) and np.array_equal(out_v, a.v) and np.array_equal(out_a, a.a)
    assert np.array_equal(b.u, a.u) and np.array_equal(b.chunk.index_at_align, a.chunk.index_at_align)
    assert np.array_equal(b.inc, a.inc)
    assert np.allclose(mean, a.mean_f_frame, rtol=1e-13, atol=0)
    assert np.allclose(mean, np.mean(a.f_frame, axis=1), rtol=1e-12, atol=0)
    # keep-the-state form: a second call without inputs continues from the device state
    a.timeSteps(5)
    b.run_from_host(5, out_u=out_u)
    assert np.array_equal(out_u, a.u)
    # and against the oracle for one realisation
    o = orc.Line1d.System_Cuspy_Laplace(seed=3 + 2 * N, **kw)
    o.u_frame = a.u_frame[2]
    o.u, o.v, o.a = u[2], v[2], acc[2]
    o.timeSteps(62)
    assert np.array_equal(o.u, out_u[2])


def test_run_from_host_checks_its_arguments():
    import frictionqpotspringblock_b200 as F

    s = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=4, seed=0, **physics(64))
    with pytest.raises(RuntimeError, match="has_shape"):
        s.run_from_host(3, u=np.zeros((4, 63)))
    with pytest.raises(RuntimeError, match="has_shape"):
        s.run_from_host(3, out_u=np.zeros((3, 64)))
    single = F.Line1d.System_Cuspy_Laplace(seed=0, **physics(64))  # falls back to the plain calls
    o = orc.Line1d.System_Cuspy_Laplace(seed=0, **physics(64))
    out = np.empty(64)
    mean = np.empty(())
    single.run_from_host(20, u=np.full(64, 0.3), out_u=out, mean_f_frame=mean)
    o.u = np.full(64, 0.3)
    o.timeSteps(20)
    assert np.array_equal(out, o.u)
    assert np.isclose(float(mean), np.mean(o.f_frame), rtol=1e-12)


def test_device_side_avalanche_bookkeeping():
    import frictionqpotspringblock_b200 as F

    R, N = 24, 300
    kw = physics(N)
    ens = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, seed=11, **kw)
    with pytest.raises(RuntimeError, match="no marked indices"):
        ens.avalanche_since_mark()
    assert np.all(ens.minimise() == 0)
    for _ in range(3):
        i_n = ens.chunk.index_at_align
        ens.mark_indices()
        ens.eventDrivenStep(1e-3, False)
        ens.eventDrivenStep(1e-3, True)
        i_k = ens.chunk.index_at_align  # after the kick: what minimise(time_activity) starts from
        # time_activity=True overwrites the handle's scratch copy of the indices: the mark survives
        assert np.all(ens.minimise(time_activity=True) == 0)
        S, A = ens.avalanche_since_mark()
        i = ens.chunk.index_at_align
        assert np.array_equal(S, np.sum(i - i_n, axis=1))
        assert np.array_equal(A, np.sum(i != i_n, axis=1))
        S2, A2 = ens.avalanche(i_n)
        assert np.array_equal(S, S2) and np.array_equal(A, A2)
        # the record the stepping kernel keeps itself (detail.h:1768-1778): relative to the start
        # of the minimisation
        S_abs, A_rec, first, last = ens.event_record()
        assert np.array_equal(S_abs, np.sum(np.abs(i - i_k), axis=1))
        assert np.array_equal(A_rec, np.sum(i != i_k, axis=1))
        assert np.array_equal(first, ens.quasistaticActivityFirst)
        assert np.array_equal(last, ens.quasistaticActivityLast)
    # the same events on the oracle, realisation 5
    o = orc.Line1d.System_Cuspy_Laplace(seed=11 + 5 * N, **kw)
    assert o.minimise() == 0
    for _ in range(3):
        j_n = o.chunk.index_at_align
        o.eventDrivenStep(1e-3, False)
        o.eventDrivenStep(1e-3, True)
        assert o.minimise(time_activity=True) == 0
    assert np.array_equal(o.chunk.index_at_align, i[5])
    assert int(np.sum(o.chunk.index_at_align - j_n)) == int(S[5])
    # wrong-sized reference indices are refused instead of read out of bounds
    with pytest.raises(RuntimeError, match="has_shape"):
        ens.avalanche(np.zeros((R, N - 1), dtype=np.int64))
    from frictionqpotspringblock_b200._capi import lib

    S = np.empty(R, dtype=np.int64)
    assert lib.fqsb_avalanche(ens._h, i_n.ctypes.data, i_n.size - 1, S.ctypes.data, None) == 3
