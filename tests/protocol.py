"""The quasistatic event-driven protocol of the reference's examples
(/root/reference/examples/Line1d_Cuspy_Laplace.py:32-61), shared by oracle and product tests."""

import numpy as np

N = 1000
BASE = dict(
    m=1.0,
    eta=2.0 * np.sqrt(3.0) / 10.0,
    mu=1.0,
    k_frame=1.0 / N,
    dt=0.1,
    shape=[N],
    seed=0,
    distribution="random",
    parameters=[2.0],
    offset=-50,
)


def make(ns1d, ns2d, name, nsp=None):
    """Construct the system of example `name` from namespaces exposing the reference classes."""
    b = dict(BASE)
    if name == "Particles_Cuspy":
        return nsp.System_Cuspy(**b)
    if name == "Line1d_Cuspy_Laplace":
        return ns1d.System_Cuspy_Laplace(k_interactions=1.0, **b)
    if name == "Line1d_Cuspy_Laplace_Nopassing":
        for k in ("m", "eta", "dt"):
            b.pop(k)
        return ns1d.System_Cuspy_Laplace_Nopassing(k_interactions=1.0, **b)
    if name == "Line1d_Cuspy_Quartic":
        return ns1d.System_Cuspy_Quartic(a1=1.0, a2=1.0, **b)
    if name == "Line1d_SemiSmooth_Laplace":
        return ns1d.System_SemiSmooth_Laplace(k_interactions=1.0, kappa=1, **b)
    if name == "Line1d_Cuspy_Laplace_LongRange":
        return ns1d.System_Cuspy_LongRange(k_interactions=1.0, alpha=1, **b)
    if name == "Line2d_Cuspy_Laplace":
        b["shape"] = [50, 50]
        b["k_frame"] = 1.0 / 2500
        return ns2d.System_Cuspy_Laplace(k_interactions=1.0, **b)
    raise KeyError(name)


def run(system, nstep, xdelta=1e-3):
    ret_u_frame = np.empty([nstep], dtype=float)
    ret_f_frame = np.empty([nstep], dtype=float)
    ret_S = np.empty([nstep], dtype=np.int64)
    for step in range(nstep):
        i_n = np.copy(system.chunk.index_at_align)
        if step == 0:
            system.u_frame = 0.0
        else:
            system.eventDrivenStep(xdelta, step % 2 == 0)
        if step % 2 == 0:
            ret = system.minimise()
            assert ret == 0
        ret_u_frame[step] = system.u_frame
        ret_f_frame[step] = np.mean(system.f_frame)
        ret_S[step] = np.sum(system.chunk.index_at_align - i_n)
    return ret_u_frame, ret_f_frame, ret_S


def check(golden, u_frame, f_frame, S):
    n = len(S)
    assert np.all(S == golden["S"][:n])
    assert np.allclose(u_frame, golden["x_frame"][:n])
    assert np.allclose(f_frame, golden["f_frame"][:n])


# ---- thermal example (/root/reference/examples/Line1d_System_Cuspy_Laplace_RandomForcing.py)
def make_thermal(ns1d, randint, **extra):
    """The athermal system minimised first, then the thermal system started from its slips.
    `randint(initstate, n, high)` = prrng.pcg32(initstate).randint([n], high)."""
    athermal = ns1d.System_Cuspy_Laplace(k_interactions=1.0, **BASE, **extra)
    athermal.minimise()
    system = ns1d.System_Cuspy_Laplace_RandomForcing(
        k_interactions=1.0, mean=0.0, stddev=0.05, seed_forcing=0,
        dinc_init=randint(0, N, 100), dinc=100 * np.ones(N, dtype=np.int64), **BASE, **extra)
    system.u = np.copy(athermal.u)
    return system


def run_thermal(system, nout, dinc=1000, delta_gamma=5e-2):
    ret_u_frame = np.empty([nout], dtype=float)
    ret_f_frame = np.empty([nout], dtype=float)
    ret_t_insta = np.empty([nout], dtype=float)
    for iout in range(nout):
        system.flowSteps(dinc, delta_gamma)
        ret_u_frame[iout] = system.u_frame
        ret_f_frame[iout] = np.mean(system.f_frame)
        ret_t_insta[iout] = system.temperature
    return ret_u_frame, ret_f_frame, ret_t_insta


def check_thermal(golden, u_frame, f_frame, t_insta):
    n = len(u_frame)
    assert np.allclose(u_frame, golden["x_frame"][:n])
    assert np.allclose(f_frame, golden["f_frame"][:n])
    assert np.allclose(t_insta, golden["t_insta"][:n])
