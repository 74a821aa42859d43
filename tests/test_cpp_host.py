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
