"""The eight prrng distributions (detail.h:31-66) in the oracle: yield spacings against closed
forms / scipy for the ones whose special functions the reference takes from boost (``normal``:
erf_inv, ``gamma``: gamma_p_inv -- both restated here, "parity unpinned")."""

import numpy as np
import pytest

from oracle import oracle as orc

scipy_special = pytest.importorskip("scipy.special")


def spacings(dist, parameters, n=4000, N=3):
    kw = dict(m=1.0, eta=0.3, mu=1.0, k_interactions=1.0, k_frame=0.1, dt=0.1, shape=[N], seed=9,
              distribution=dist, parameters=parameters, offset=-50.0, nchunk=n)
    s = orc.Line1d.System_Cuspy_Laplace(**kw)
    y = s.chunk.data
    return np.diff(y, axis=1), y[:, 0]


def uniform_draws(n, N=3):
    d, _ = spacings("random", [1.0], n, N)  # spacing = r itself
    return d


@pytest.mark.parametrize("dist,par,fn", [
    ("random", [2.0, 0.5], lambda r: r * 2.0 + 0.5),
    ("delta", [1.5, 0.25], lambda r: np.full_like(r, 1.75)),
    ("exponential", [2.0, 0.1], lambda r: -np.log(1 - r) * 2.0 + 0.1),
    ("power", [3.0, 0.2], lambda r: (1 - r) ** (1 / 4.0) + 0.2),
    ("pareto", [2.0, 1.5, 0.1], lambda r: 1.5 * (1 - r) ** (-0.5) + 0.1),
    ("weibull", [2.0, 1.1, 1e-3], lambda r: 1.1 * (-np.log(1 - r)) ** 0.5 + 1e-3),
    ("normal", [5.0, 0.7, 0.0],
     lambda r: 5.0 + 0.7 * np.sqrt(2.0) * scipy_special.erfinv(2 * r - 1)),
    ("gamma", [2.5, 1.3, 0.05], lambda r: 1.3 * scipy_special.gammaincinv(2.5, r) + 0.05),
    ("gamma", [0.6, 1.0, 0.5], lambda r: scipy_special.gammaincinv(0.6, r) + 0.5),
])
def test_spacings_follow_the_published_maps(dist, par, fn):
    d, _ = spacings(dist, par)
    r = uniform_draws(d.shape[1] + 1)[:, : d.shape[1]] if dist != "delta" else np.zeros_like(d)
    want = fn(r)
    # the cumulative sum rounds: compare spacings at the accuracy of a difference of sums ~ 1e4
    assert np.allclose(d, want, rtol=1e-9, atol=1e-9), np.abs(d - want).max()


def test_special_function_inverses():
    L = orc.lib()
    for z in (-0.999999, -0.5, -1e-9, 0.0, 1e-9, 0.3, 0.9, 0.999999999):
        assert np.isclose(L.orc_erf_inv(z), scipy_special.erfinv(z), rtol=1e-14, atol=1e-300)
    for a in (0.3, 1.0, 2.0, 7.5, 50.0):
        for p in (1e-9, 0.01, 0.3, 0.5, 0.9, 0.999999):
            assert np.isclose(L.orc_gamma_p_inv(a, p), scipy_special.gammaincinv(a, p),
                              rtol=1e-13, atol=0)
