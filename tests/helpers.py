"""Shared helpers of the GPU parity tests: build the same system in the CPU oracle and in the
CUDA product from one set of reference-style keyword arguments."""

import numpy as np

from oracle import oracle as orc


def product():
    import frictionqpotspringblock_b200 as F

    return F


def pair(module, cls, **kw):
    """(oracle system, product system) of class `cls` in `module` ("Line1d" / "Line2d")."""
    F = product()
    extra = {k: kw.pop(k) for k in ("kernel", "device") if k in kw}
    o = getattr(getattr(orc, module), cls)(**kw)
    p = getattr(getattr(F, module), cls)(**kw, **extra)
    return o, p


def assert_same_state(o, p, exact=True, rtol=1e-12):
    """positions, velocities, accelerations, forces, well indices of oracle `o` and product `p`."""
    assert np.array_equal(o.chunk.index_at_align, p.chunk.index_at_align)
    for name in ("u", "v", "a", "f", "f_potential", "f_frame", "f_interactions", "f_damping"):
        x, y = getattr(o, name), getattr(p, name)
        if exact:
            assert np.array_equal(x, y), name
        else:
            scale = max(np.abs(x).max(), 1e-300)
            assert np.abs(x - y).max() <= rtol * scale, (name, np.abs(x - y).max(), scale)
    assert np.array_equal(o.chunk.left_of_align, p.chunk.left_of_align)
    assert np.array_equal(o.chunk.right_of_align, p.chunk.right_of_align)
    assert o.inc == p.inc
    if exact:
        assert o.u_frame == p.u_frame
    else:
        assert np.isclose(o.u_frame, p.u_frame, rtol=1e-12, atol=0)
