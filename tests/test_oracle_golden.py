"""Pins the CPU oracle on the reference's committed golden vectors.

The assertions are the reference's own (examples/Line1d_Cuspy_Laplace.py:64-67): S exact,
x_frame and f_frame np.allclose. By default the expensive examples are run over a prefix of
the protocol so the CPU suite stays within minutes; FQSB_FULL_GOLDEN=1 runs all of them to the
end (all six pass in full: see DESIGN.md "Oracle").
"""

import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests import protocol

FULL = os.environ.get("FQSB_FULL_GOLDEN", "0") == "1"

CASES = {
    # name: (steps by default, steps in the golden)
    "Line1d_Cuspy_Laplace": (120, 1000),
    "Line1d_Cuspy_Laplace_Nopassing": (1000, 1000),
    "Line1d_Cuspy_Quartic": (120, 1000),
    "Line1d_SemiSmooth_Laplace": (300, 1000),
    "Line1d_Cuspy_Laplace_LongRange": (12, 200),
    "Line2d_Cuspy_Laplace": (40, 1000),
    "Particles_Cuspy": (300, 2000),
}


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_golden(name, golden_dir):
    nstep = CASES[name][1] if FULL else CASES[name][0]
    golden = np.load(golden_dir / f"{name}.npz")
    system = protocol.make(orc.Line1d, orc.Line2d, name, orc.Particles)
    protocol.check(golden, *protocol.run(system, nstep))


def test_oracle_reproduces_thermal_golden(golden_dir):
    """examples/Line1d_System_Cuspy_Laplace_RandomForcing.py:76-80 (x_frame, f_frame, t_insta
    np.allclose): pins pcg32.normal (erf_inv), pcg32.randint and RandomNormalForcing."""
    golden = np.load(golden_dir / "Line1d_System_Cuspy_Laplace_RandomForcing.npz")
    system = protocol.make_thermal(orc.Line1d, orc.pcg32_randint)
    protocol.check_thermal(golden, *protocol.run_thermal(system, 500 if FULL else 60))


def test_oracle_random_forcing_reference_test():
    """tests/test_Line1d.py:566-600 (Test_System_Cuspy_Laplace_RandomForcing.test_interactions)
    with prrng.pcg32(0).normal restated by the oracle."""
    N = 10
    system = orc.Line1d.System_Cuspy_Laplace_RandomForcing(
        m=1, eta=1, mu=1, k_interactions=1, k_frame=0.1, dt=1, mean=0, stddev=1, seed_forcing=0,
        dinc_init=np.ones(N, dtype=int), dinc=np.ones(N, dtype=int), shape=[N], seed=0,
        distribution="delta", parameters=[1.0], offset=-49.5, nchunk=100)
    assert system.residual < 1e-5
    draws = orc.pcg32_normal(0, 2 * N, 0, 1)
    system.inc += 1
    system.refresh()
    assert system.residual > 1e-5
    assert np.allclose(system.external.f_thermal, draws[:N])
    system.inc += 1
    system.refresh()
    assert system.residual > 1e-5
    assert np.allclose(system.external.f_thermal, draws[N:])
    with pytest.raises(RuntimeError, match="Minimisation not implementated"):
        system.minimise()


def test_erf_inv_against_erf():
    """the restated boost::math::erf_inv inverts erf to the last bits over the 32-bit grid of
    prrng's doubles (including both tails)"""
    import math

    r = np.concatenate([orc.pcg32_draws(7, 2000), [2.0**-32, 1 - 2.0**-32, 0.5, 2.0**-20]])
    for z in 2.0 * r - 1.0:
        x = orc.lib().orc_erf_inv(float(z))
        # erf is flat in the tails: compare through erfc there
        if abs(z) < 0.5:
            assert abs(math.erf(x) - z) <= 4e-16 * max(abs(z), 1e-300)
        else:
            q = 1.0 - abs(z)
            assert abs(math.erfc(abs(x)) - q) <= 1e-14 * q
    assert orc.lib().orc_erf_inv(0.0) == 0.0


def test_pcg32_known_answer():
    """pcg32 reference vector: the PCG paper's demo seeding (42, 54) gives 0xa15c02b7 first."""
    import ctypes as C

    # next_double = (uint32 << 20 | 0x3ff...) - 1  ==  uint32 / 2^32 exactly
    d = orc.pcg32_draws(42, 6, initseq=54)
    u32 = (d * 2.0**32).astype(np.uint64)
    assert [hex(int(i)) for i in u32] == [
        "0xa15c02b7", "0x7b47f409", "0xba1d3330", "0x83d2f293", "0xbfa4784b", "0xcbed606e"]
    assert C.sizeof(orc.Params) == 160
