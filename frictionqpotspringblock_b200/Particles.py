"""``FrictionQPotSpringBlock.Particles``: independent particles (no interactions),
/root/reference/include/FrictionQPotSpringBlock/Particles.h:93-311 (SURVEY.md section 8f, rows
N3/N4). The cuspy systems run the interaction-free kernels; the semi-smooth and smooth particle
systems reuse the ``*_Laplace`` line kernels with ``k_interactions = 0`` (the interaction term is
then an exact zero: same u, v, a and forces)."""

from . import Line1d as _l1

__all__ = []


def _define(name, potential, interactions, lead, doc="", forcing=False):
    # (`lead` without k_interactions / a1: Line1d._define then passes k1 = 0)
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
_define("SemiSmooth", "SemiSmooth", "Laplace1d", ("m", "eta", "mu", "kappa", "k_frame", "dt"),
        doc="Particles.h:233-270: semi-smooth potential, no interactions.")
_define("Smooth", "Smooth", "Laplace1d", ("m", "eta", "mu", "k_frame", "dt"),
        doc="Particles.h:276-311: smooth potential, no interactions.")
