"""``FrictionQPotSpringBlock.Line1d``: the 1-D systems of the reference
(/root/reference/include/FrictionQPotSpringBlock/Line1d.h:112-677, bound in
/root/reference/python/main.cpp:470-760) with identical constructor signatures.

Each ``System_*`` class also exists as ``Ensemble_*`` (same arguments plus ``nrealisations``):
the batch of independent disorder realisations the B200 build shards across GPUs.
Keyword-only extras on every class: ``device`` (CUDA ordinal, -1 = current), ``kernel``
(0 auto, 1 resident, 2 streaming, 3 temporally blocked) and ``contracted`` (default False: every
expression in the reference's evaluation order without FMA contraction, results bit-identical to
the reference arithmetic; True opts in to the resident kernels built with FMA contraction --
about 1.4x fewer FP64 instructions per step, same yield landscape, trajectories equal to rounding).
"""

from ._system import Ensemble, System

__all__ = []


def version_dependencies():
    """Line1d.h:34-37. The third-party headers of the reference are not used by this build."""
    from . import version

    return [f"frictionqpotspringblock_b200={version()}", "cuda=sm_100a", "prrng=restated",
            "goosefem=restated"]


def version_compiler():
    """Line1d.h:43-46."""
    return ["nvcc=12.9", "std=c++17", "arch=sm_100a"]


_FORCING = ("mean", "stddev", "seed_forcing", "dinc_init", "dinc")


def _define(name, potential, interactions, lead, minimisation=0, doc="", forcing=False):
    """Create System_<name> / Ensemble_<name>; `lead` = leading constructor arguments.
    `forcing`: External = RandomNormalForcing (Line1d.h:261-330: its five arguments follow
    `lead`; minimisation = None)."""
    if forcing:
        lead = tuple(lead) + _FORCING
        minimisation = 2

    def make(base, prefix):
        def __init__(self, *args, **kw):
            names = list(lead) + ["shape", "seed", "distribution", "parameters", "offset",
                                  "nchunk"]
            if minimisation == 1:
                names += ["eta", "dt"]  # Line1d.h:199-211
            if len(args) > len(names):
                raise TypeError(f"{prefix}{name}: too many positional arguments")
            for key, val in zip(names, args):
                if key in kw:
                    raise TypeError(f"{prefix}{name}: multiple values for argument '{key}'")
                kw[key] = val
            required = list(lead) + ["shape", "seed", "distribution", "parameters"]
            missing = [k for k in required if k not in kw]
            if missing:
                raise TypeError(f"{prefix}{name}: missing arguments {missing}")
            k1 = k2 = kappa = 0.0
            for key in ("k_interactions", "a1"):
                if key in lead:
                    k1 = kw.pop(key)
            if "k2" in lead:
                k1 = kw.pop("k2")
            for key in ("a2", "k4", "alpha"):
                if key in lead:
                    k2 = kw.pop(key)
            if "kappa" in lead:
                kappa = kw.pop("kappa")
            if forcing:
                kw["forcing"] = tuple(kw.pop(key) for key in _FORCING)
            base.__init__(
                self, potential, interactions, kw.pop("shape"),
                m=kw.pop("m", 1.0), eta=kw.pop("eta", 0.0), mu=kw.pop("mu"), kappa=kappa,
                k1=k1, k2=k2, k_frame=kw.pop("k_frame"), dt=kw.pop("dt", 0.0),
                seed=kw.pop("seed"), distribution=kw.pop("distribution"),
                parameters=kw.pop("parameters"), offset=kw.pop("offset", -100.0),
                nchunk=kw.pop("nchunk", 5000), minimisation=minimisation, **kw)

        cls = type(prefix + name, (base,), {"__init__": __init__, "__doc__": doc})
        cls.__module__ = __name__
        return cls

    g = globals()
    g["System_" + name] = make(System, "System_")
    g["Ensemble_" + name] = make(Ensemble, "Ensemble_")
    __all__.extend(["System_" + name, "Ensemble_" + name])


_STD = ("m", "eta", "mu", "k_interactions", "k_frame", "dt")

_define("Cuspy_Laplace", "Cuspy", "Laplace1d", _STD,
        doc="Line1d.h:112-162: cuspy potential, Laplace interactions, velocity Verlet.")
_define("Cuspy_Laplace_Nopassing", "Cuspy", "Laplace1d", ("mu", "k_interactions", "k_frame"),
        minimisation=1,
        doc="Line1d.h:173-238: overdamped no-passing minimisation only (no dynamics).")
_define("SemiSmooth_Laplace", "SemiSmooth", "Laplace1d",
        ("m", "eta", "mu", "kappa", "k_interactions", "k_frame", "dt"),
        doc="Line1d.h:336-377.")
_define("Smooth_Laplace", "Smooth", "Laplace1d", _STD, doc="Line1d.h:383-422.")
_define("Cuspy_Quartic", "Cuspy", "Quartic1d", ("m", "eta", "mu", "a1", "a2", "k_frame", "dt"),
        doc="Line1d.h:428-480.")
_define("Cuspy_QuarticGradient", "Cuspy", "QuarticGradient1d",
        ("m", "eta", "mu", "k2", "k4", "k_frame", "dt"), doc="Line1d.h:562-614.")
_define("Cuspy_Laplace_RandomForcing", "Cuspy", "Laplace1d", _STD, forcing=True,
        doc="Line1d.h:261-330: System_Cuspy_Laplace plus a random force per block drawn from a "
            "normal distribution (External = RandomNormalForcing); no minimisation.")
_define("Cuspy_Quartic_RandomForcing", "Cuspy", "Quartic1d",
        ("m", "eta", "mu", "a1", "a2", "k_frame", "dt"), forcing=True, doc="Line1d.h:486-556.")
_define("Cuspy_LongRange", "Cuspy", "LongRange1d",
        ("m", "eta", "mu", "k_interactions", "alpha", "k_frame", "dt"), doc="Line1d.h:620-672.")
