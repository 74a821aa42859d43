"""The fused single-pass CPU flavour of timeSteps (oracle/fqsb_oracle.c: orc_time_steps_fused --
BASELINE.md section 4's "best-case CPU", timed by bench.py next to the faithful multi-pass port)
must reproduce the faithful restatement bit for bit."""

import numpy as np

from oracle import oracle as orc


def test_fused_flavour_is_bit_identical_to_the_faithful_port():
    N = 300
    kw = dict(m=1.3, eta=0.4, mu=0.9, k_interactions=1.1, k_frame=1.0 / N, dt=0.1, shape=[N],
              seed=5, distribution="random", parameters=[2.0], offset=-50)
    a = orc.Line1d.System_Cuspy_Laplace(**kw)
    b = orc.Line1d.System_Cuspy_Laplace(**kw)
    for s in (a, b):
        assert s.minimise() == 0
        s.eventDrivenStep(1e-3, False)
        s.eventDrivenStep(1e-3, True)
    for n in (1, 2, 37, 400):
        a.timeSteps(n)
        b.timeSteps_fused(n)
        assert a.inc == b.inc
        assert np.array_equal(a.chunk.index_at_align, b.chunk.index_at_align)
        for name in ("u", "v", "a", "f", "f_potential", "f_frame", "f_interactions", "f_damping"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), (n, name)
    # strongly driven: many well changes per step
    for s in (a, b):
        s.u_frame = s.u_frame + 25.0
    a.timeSteps(300)
    b.timeSteps_fused(300)
    assert np.array_equal(a.u, b.u) and np.array_equal(a.v, b.v)
    assert np.array_equal(a.chunk.index_at_align, b.chunk.index_at_align)
    assert np.sum(a.chunk.index_at_align) > 0
    # other systems fall back to the faithful path
    q = orc.Line1d.System_Cuspy_Quartic(m=1.0, eta=0.3, mu=1.0, a1=1.0, a2=0.5, k_frame=1.0 / N,
                                        dt=0.1, shape=[N], seed=0, distribution="random",
                                        parameters=[2.0], offset=-50)
    r = orc.Line1d.System_Cuspy_Quartic(m=1.0, eta=0.3, mu=1.0, a1=1.0, a2=0.5, k_frame=1.0 / N,
                                        dt=0.1, shape=[N], seed=0, distribution="random",
                                        parameters=[2.0], offset=-50)
    for s in (q, r):
        s.u_frame = 3.0
    q.timeSteps(20)
    r.timeSteps_fused(20)
    assert np.array_equal(q.u, r.u)


def test_cpu_ensemble_flavours_agree():
    kw = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_frame=1.0 / 256, dt=0.1)
    par = orc.make_params("Cuspy", "Laplace1d", 0, [256], kw["m"], kw["eta"], kw["mu"], 0.0, 1.0,
                          0.0, kw["k_frame"], kw["dt"], 0, "random", [2.0], -50, 5000)
    e1 = orc.CpuEnsemble(par, 6, 3)
    e2 = orc.CpuEnsemble(par, 6, 3)
    e2.configure(fused=True)
    _, c1 = e1.time_steps(200)
    _, c2 = e2.time_steps(200)
    assert c1 == c2  # checksum of all slips
