"""ctypes binding of ``libfqsb.so`` (C ABI: ``include/fqsb.h``).

There is no CPU fallback: if the CUDA library is missing this module raises at import time, and
constructing a system without a CUDA device raises ``RuntimeError`` from ``fqsb_create``.
"""

from __future__ import annotations

import ctypes as C
import pathlib

_HERE = pathlib.Path(__file__).resolve().parent
LIBRARY = _HERE / "libfqsb.so"

POT = {"Cuspy": 0, "SemiSmooth": 1, "Smooth": 2}
INT = {
    "None": 0,
    "Laplace1d": 1,
    "Quartic1d": 2,
    "QuarticGradient1d": 3,
    "LongRange1d": 4,
    "Laplace2d": 5,
    "QuarticGradient2d": 6,
}
DIST = {  # detail.h:31-66
    "random": 0,
    "delta": 1,
    "exponential": 2,
    "power": 3,
    "gamma": 4,
    "pareto": 5,
    "weibull": 6,
    "normal": 7,
}
ARRAY = {"u": 0, "v": 1, "a": 2, "f": 3, "f_potential": 4, "f_frame": 5, "f_interactions": 6,
         "f_damping": 7}


class Params(C.Structure):
    """``fqsb_params`` of include/fqsb.h."""

    _fields_ = [
        ("potential", C.c_int32),
        ("interactions", C.c_int32),
        ("minimisation", C.c_int32),
        ("rank", C.c_int32),
        ("shape", C.c_int64 * 2),
        ("m", C.c_double),
        ("eta", C.c_double),
        ("mu", C.c_double),
        ("kappa", C.c_double),
        ("k1", C.c_double),
        ("k2", C.c_double),
        ("k_frame", C.c_double),
        ("dt", C.c_double),
        ("seed", C.c_uint64),
        ("distribution", C.c_int32),
        ("nparameters", C.c_int32),
        ("parameters", C.c_double * 4),
        ("offset", C.c_double),
        ("nchunk", C.c_int64),
        ("nrealisations", C.c_int64),
        ("seed_stride", C.c_int64),
        ("device", C.c_int32),
        ("kernel", C.c_int32),
        ("seed_first", C.c_int64),
        ("seed_period", C.c_int64),
    ]


# name: (restype, argtypes) -- every symbol include/fqsb.h declares
_P = C.c_void_p
SIGNATURES = {
    "fqsb_last_error": (C.c_char_p, []),
    "fqsb_abi_version": (C.c_int, []),
    "fqsb_version": (C.c_char_p, []),
    "fqsb_device_count": (C.c_int, []),
    "fqsb_create": (C.c_int, [_P, _P]),
    "fqsb_destroy": (None, [_P]),
    "fqsb_get_params": (C.c_int, [_P, _P]),
    "fqsb_size": (C.c_int64, [_P]),
    "fqsb_nrealisations": (C.c_int64, [_P]),
    "fqsb_set_stream": (C.c_int, [_P, _P]),
    "fqsb_get_stream": (_P, [_P]),
    "fqsb_set_u": (C.c_int, [_P, _P, C.c_int64]),
    "fqsb_set_v": (C.c_int, [_P, _P, C.c_int64]),
    "fqsb_set_a": (C.c_int, [_P, _P, C.c_int64]),
    "fqsb_set_u_frame": (C.c_int, [_P, _P]),
    "fqsb_set_inc": (C.c_int, [_P, _P]),
    "fqsb_set_t": (C.c_int, [_P, _P]),
    "fqsb_refresh": (C.c_int, [_P]),
    "fqsb_quench": (C.c_int, [_P]),
    "fqsb_get": (C.c_int, [_P, C.c_int, _P, C.c_int64]),
    "fqsb_get_device": (C.c_int, [_P, C.c_int, _P]),
    "fqsb_get_u_frame": (C.c_int, [_P, _P]),
    "fqsb_get_inc": (C.c_int, [_P, _P]),
    "fqsb_get_t": (C.c_int, [_P, _P]),
    "fqsb_residual": (C.c_int, [_P, _P]),
    "fqsb_temperature": (C.c_int, [_P, _P]),
    "fqsb_mean_f_frame": (C.c_int, [_P, _P]),
    "fqsb_qs_activity": (C.c_int, [_P, _P, _P]),
    "fqsb_time_steps": (C.c_int, [_P, C.c_int64]),
    "fqsb_flow_steps": (C.c_int, [_P, C.c_int64, C.c_double]),
    "fqsb_run_from_host": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P, _P, _P]),
    "fqsb_time_steps_until_event": (C.c_int, [_P, C.c_double, C.c_int64, C.c_int64, _P]),
    "fqsb_minimise": (C.c_int, [_P, C.c_double, C.c_int64, C.c_int64, C.c_int, C.c_int, _P]),
    "fqsb_minimise_truncate": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_double, C.c_int64,
                                         C.c_int64, C.c_int, C.c_int, _P]),
    "fqsb_max_uniform_displacement": (C.c_int, [_P, C.c_int, _P]),
    "fqsb_event_driven_step": (C.c_int, [_P, C.c_double, C.c_int, C.c_int, _P]),
    "fqsb_trigger": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_double, C.c_int]),
    "fqsb_advance_to_fixed_force": (C.c_int, [_P, _P, C.c_int]),
    "fqsb_chunk_index_at_align": (C.c_int, [_P, _P, C.c_int64]),
    "fqsb_chunk_left_of_align": (C.c_int, [_P, _P, C.c_int64]),
    "fqsb_chunk_right_of_align": (C.c_int, [_P, _P, C.c_int64]),
    "fqsb_chunk_align": (C.c_int, [_P, _P, C.c_int64]),
    "fqsb_chunk_data": (C.c_int, [_P, _P, C.c_int64, _P]),
    "fqsb_chunk_state_at": (C.c_int, [_P, _P, _P, C.c_int64]),
    "fqsb_chunk_restore": (C.c_int, [_P, _P, _P, _P, C.c_int64]),
    "fqsb_avalanche": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "fqsb_mark_indices": (C.c_int, [_P]),
    "fqsb_avalanche_since_mark": (C.c_int, [_P, _P, _P]),
    "fqsb_event_record": (C.c_int, [_P, _P, _P, _P, _P]),
    "fqsb_set_owned_range": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "fqsb_logged_steps": (C.c_int, [_P, C.c_int64, _P]),
    "fqsb_snapshot": (C.c_int, [_P]),
    "fqsb_rollback": (C.c_int, [_P]),
    "fqsb_export_cells": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int]),
    "fqsb_import_cells": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int]),
    "fqsb_advance_uniformly": (C.c_int, [_P, _P, _P]),
    "fqsb_reduce_sums": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int64, _P]),
    "fqsb_slab_init": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, C.c_int]),
    "fqsb_slab_ipc_handle": (C.c_int, [_P, _P]),
    "fqsb_slab_connect": (C.c_int, [_P, _P, _P]),
    "fqsb_slab_info": (C.c_int, [_P, _P]),
    "fqsb_plan_blocked": (C.c_int, [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, _P]),
    "fqsb_slab_exchange": (C.c_int, [_P, C.c_int]),
    "fqsb_slab_time_steps": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_double]),
    "fqsb_slab_minimise": (C.c_int, [_P, C.c_int, C.c_double, C.c_int64, C.c_int64, C.c_int64,
                                     C.c_int, _P, _P]),
    "fqsb_slab_sums": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "fqsb_slab_mark_indices": (C.c_int, [_P, C.c_int]),
    "fqsb_slab_event_driven_step": (C.c_int, [_P, C.c_int, C.c_double, C.c_int, C.c_int, _P]),
    "fqsb_slab_first_stop": (C.c_int64, [_P, C.c_int64, C.c_double, C.c_int64, _P, _P]),
    "fqsb_enable_random_forcing": (C.c_int, [_P, C.c_double, C.c_double, C.c_uint64, C.c_int64,
                                             _P, _P, C.c_int64]),
    "fqsb_external_get_f_thermal": (C.c_int, [_P, _P, C.c_int64]),
    "fqsb_external_set_f_thermal": (C.c_int, [_P, _P, C.c_int64]),
    "fqsb_external_get_next": (C.c_int, [_P, _P, C.c_int64]),
    "fqsb_external_set_next": (C.c_int, [_P, _P, C.c_int64]),
    "fqsb_external_get_state": (C.c_int, [_P, _P]),
    "fqsb_external_set_state": (C.c_int, [_P, _P]),
    "fqsb_host_alloc": (_P, [C.c_size_t]),
    "fqsb_host_free": (None, [_P]),
    "fqsb_launch_count": (C.c_int64, [_P]),
    "fqsb_step_count": (C.c_int64, [_P]),
    "fqsb_last_kernel": (C.c_char_p, [_P]),
    "fqsb_last_kernel_seconds": (C.c_double, [_P]),
    "fqsb_last_kernel_launches": (C.c_int64, [_P]),
}


def _load():
    if not LIBRARY.exists():
        raise ImportError(
            f"{LIBRARY} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C "
            "frictionqpotspringblock_b200/csrc). There is no CPU fallback."
        )
    lib = C.CDLL(str(LIBRARY))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int):
    """Map an fqsb_status to the reference's exception type (std::runtime_error -> RuntimeError)."""
    if rc != 0:
        raise RuntimeError(lib.fqsb_last_error().decode())
