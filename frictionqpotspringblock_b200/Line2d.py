"""``FrictionQPotSpringBlock.Line2d``: the 2-D systems of the reference
(/root/reference/include/FrictionQPotSpringBlock/Line2d.h:77-162) plus the overdamped
no-passing system on the 2-D lattice (new: BASELINE config #5 has no reference implementation,
SURVEY.md F7)."""

from . import Line1d as _l1
from ._system import Ensemble, System  # noqa: F401

__all__ = []

version_dependencies = _l1.version_dependencies
version_compiler = _l1.version_compiler


def _define(name, potential, interactions, lead, minimisation=0, doc=""):
    # reuse the Line1d factory, then move the classes here
    _l1._define("__2d_" + name, potential, interactions, lead, minimisation, doc)
    for prefix in ("System_", "Ensemble_"):
        cls = _l1.__dict__.pop(prefix + "__2d_" + name)
        _l1.__all__.remove(prefix + "__2d_" + name)
        cls.__name__ = cls.__qualname__ = prefix + name
        cls.__module__ = __name__
        globals()[prefix + name] = cls
        __all__.append(prefix + name)


_define("Cuspy_Laplace", "Cuspy", "Laplace2d",
        ("m", "eta", "mu", "k_interactions", "k_frame", "dt"), doc="Line2d.h:77-117.")
_define("Cuspy_QuarticGradient", "Cuspy", "QuarticGradient2d",
        ("m", "eta", "mu", "k2", "k4", "k_frame", "dt"), doc="Line2d.h:123-162.")
_define("Cuspy_Laplace_Nopassing", "Cuspy", "Laplace2d", ("mu", "k_interactions", "k_frame"),
        minimisation=1, doc="2-D generalisation of Line1d.h:173-238 (new).")
