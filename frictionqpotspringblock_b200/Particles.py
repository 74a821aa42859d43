"""``FrictionQPotSpringBlock.Particles``: independent particles (no interactions),
/root/reference/include/FrictionQPotSpringBlock/Particles.h:93-135. Only the cuspy system is
and its thermal variant are provided; the (semi-)smooth
variants are out of scope."""

from . import Line1d as _l1

__all__ = []


def _define(name, potential, interactions, lead, doc="", forcing=False):
    _l1._define("__p_" + name, potential, interactions, lead, 0, doc, forcing)
    for prefix in ("System_", "Ensemble_"):
        cls = _l1.__dict__.pop(prefix + "__p_" + name)
        _l1.__all__.remove(prefix + "__p_" + name)
        cls.__name__ = cls.__qualname__ = prefix + name
        cls.__module__ = __name__
        globals()[prefix + name] = cls
        __all__.append(prefix + name)


_define("Cuspy", "Cuspy", "None", ("m", "eta", "mu", "k_frame", "dt"), doc="Particles.h:93-135.")
_define("Cuspy_RandomForcing", "Cuspy", "None", ("m", "eta", "mu", "k_frame", "dt"), forcing=True,
        doc="Particles.h:168-230: System_Cuspy plus External = RandomNormalForcing.")
