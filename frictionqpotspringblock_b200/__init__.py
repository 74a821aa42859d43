"""B200-native integrator with the class surface of tdegeus/FrictionQPotSpringBlock.

    import frictionqpotspringblock_b200 as FrictionQPotSpringBlock
    system = FrictionQPotSpringBlock.Line1d.System_Cuspy_Laplace(m=1, eta=..., ...)

All numerical work runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/fqsb.h`` (``libfqsb.so``); importing this package without that library fails.
"""

from . import _capi  # noqa: F401  (raises ImportError when libfqsb.so is missing)
from ._system import Ensemble, System  # noqa: F401
from . import Line1d, Line2d, Particles  # noqa: F401


def version() -> str:
    """config.h:209-212."""
    return _capi.lib.fqsb_version().decode()


def device_count() -> int:
    return int(_capi.lib.fqsb_device_count())
