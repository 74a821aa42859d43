"""The C++ host surface (include/fqsb.hpp: reference class names over the C ABI) compiles with a
plain g++, links against libfqsb.so, and on a GPU reproduces the reference's golden."""

import pathlib
import subprocess

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
PKG = ROOT / "frictionqpotspringblock_b200"
EXE = ROOT / "tests" / "cpp" / "example_line1d"


def build():
    subprocess.run(
        ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-I", str(ROOT / "include"),
         str(ROOT / "tests" / "cpp" / "example_line1d.cpp"), "-o", str(EXE),
         "-L", str(PKG), "-lfqsb", f"-Wl,-rpath,{PKG}"],
        check=True)


def test_cpp_host_header_compiles_and_links():
    build()
    assert EXE.exists()


@pytest.mark.gpu
def test_cpp_host_reproduces_golden(golden_dir):
    build()
    out = subprocess.run([str(EXE), "60"], check=True, capture_output=True, text=True).stdout
    rows = [line.split() for line in out.splitlines() if line and line[0].isdigit()]
    golden = np.load(golden_dir / "Line1d_Cuspy_Laplace.npz")
    n = len(rows)
    assert n == 60
    assert np.all(np.array([int(r[3]) for r in rows]) == golden["S"][:n])
    assert np.allclose([float(r[1]) for r in rows], golden["x_frame"][:n])
    assert np.allclose([float(r[2]) for r in rows], golden["f_frame"][:n])
    assert "error fqsb: assertion failed (tol < 1.0)" in out
    # system.chunk(), Ensemble<S> and Slab<S, 1> of include/fqsb.hpp
    assert "chunk ok" in out and "ensemble ok" in out and "slab ok" in out


def _pybind():
    import importlib
    import sys

    sys.path.insert(0, str(PKG))
    try:
        return importlib.import_module("_FrictionQPotSpringBlock")
    finally:
        sys.path.remove(str(PKG))


def test_pybind_module_builds_and_refuses_without_gpu():
    """The reference's binder templates (python/main.cpp) re-bound on include/fqsb.hpp."""
    P = _pybind()
    assert P.version() == "0.1.0"
    for name in ("System_Cuspy_Laplace", "System_Cuspy_Laplace_Nopassing",
                 "System_SemiSmooth_Laplace", "System_Smooth_Laplace", "System_Cuspy_Quartic",
                 "System_Cuspy_QuarticGradient", "System_Cuspy_LongRange",
                 "System_Cuspy_Laplace_RandomForcing", "System_Cuspy_Quartic_RandomForcing"):
        assert hasattr(P.Line1d, name)
    assert hasattr(P.detail, "RandomNormalForcing_1")
    for name in ("System_Cuspy", "System_Cuspy_RandomForcing", "System_SemiSmooth",
                 "System_Smooth"):
        assert hasattr(P.Particles, name)
    assert hasattr(P.Line2d, "System_Cuspy_Laplace")
    import frictionqpotspringblock_b200 as F

    if F.device_count() == 0:
        with pytest.raises(RuntimeError, match="no CUDA device"):
            P.Line1d.System_Cuspy_Laplace(m=1, eta=1, mu=1, k_interactions=1, k_frame=0.1,
                                          dt=0.1, shape=[8], seed=0, distribution="random",
                                          parameters=[2.0])


@pytest.mark.gpu
def test_pybind_module_reproduces_golden(golden_dir):
    from tests import protocol

    P = _pybind()

    system = P.Line1d.System_Cuspy_Laplace(k_interactions=1.0, **protocol.BASE)
    golden = np.load(golden_dir / "Line1d_Cuspy_Laplace.npz")
    # the examples read system.chunk.index_at_align (python/main.cpp:65-70)
    protocol.check(golden, *protocol.run(system, 60))
    ch = system.chunk
    i, st, data = ch.index_at_align, ch.start, ch.data
    rows = np.arange(i.size)
    assert data.shape == (i.size, ch.chunk_size)
    assert np.array_equal(data[rows, i - st], ch.left_of_align)
    assert np.array_equal(data[rows, i - st + 1], ch.right_of_align)
    assert np.array_equal(ch.chunk_index_at_align, i - st)
    state = ch.state_at(st)
    ch.restore(state, data[:, 0].copy(), st)
    assert np.array_equal(ch.index_at_align, i)
    ch.align(system.u + 1.0)
    assert np.all(ch.index_at_align >= i)
    system.refresh()
    assert np.array_equal(ch.index_at_align, i)


@pytest.mark.gpu
def test_pybind_thermal_system_matches_ctypes_surface():
    """System_Cuspy_Laplace_RandomForcing through include/fqsb.hpp + pybind == ctypes package."""
    import frictionqpotspringblock_b200 as F

    P = _pybind()
    N = 50
    kw = dict(m=1.0, eta=0.3, mu=1.0, k_interactions=1.0, k_frame=1.0 / N, dt=0.1, mean=0.0,
              stddev=0.1, seed_forcing=4, dinc_init=np.arange(N) % 5, dinc=3 * np.ones(N, dtype=int),
              shape=[N], seed=1, distribution="random", parameters=[2.0], offset=-50)
    a = P.Line1d.System_Cuspy_Laplace_RandomForcing(**kw)
    b = F.Line1d.System_Cuspy_Laplace_RandomForcing(**kw)
    for s in (a, b):
        s.flowSteps(40, 0.1)
    assert np.array_equal(a.u, b.u) and np.array_equal(a.v, b.v)
    assert np.array_equal(a.external.f_thermal, b.external.f_thermal)
    assert np.array_equal(a.external.next, b.external.next)
    assert a.external.state == b.external.state
    assert repr(a.external) == "<FrictionQPotSpringBlock.detail.RandomNormalForcing_1>"
    with pytest.raises(RuntimeError, match="Minimisation not implementated"):
        a.minimise()
