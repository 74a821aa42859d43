"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE).

Mirrors the reference's Python surface (/root/reference/python/main.cpp:46-226) on top of
``oracle/libfqsb_oracle.so`` so that parity tests read like the reference's own tests.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline / --impl reference)
may import this module; the product package never does.
"""

from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None

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
DIST = {
    "random": 0,
    "delta": 1,
    "exponential": 2,
    "power": 3,
    "gamma": 4,
    "pareto": 5,
    "weibull": 6,
    "normal": 7,
}


class Params(C.Structure):
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
    ]


def build(force: bool = False) -> pathlib.Path:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    so = _HERE / "libfqsb_oracle.so"
    src = _HERE / "fqsb_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = _HERE / "libfqsb_oracle.so"
        if not so.exists():
            build()
        L = C.CDLL(str(so))
        L.orc_last_error.restype = C.c_char_p
        L.orc_size.restype = C.c_int64
        L.orc_u_frame.restype = C.c_double
        L.orc_inc.restype = C.c_int64
        L.orc_residual.restype = C.c_double
        L.orc_temperature.restype = C.c_double
        L.orc_qs_first.restype = C.c_int64
        L.orc_qs_last.restype = C.c_int64
        L.orc_draw_to_spacing.restype = C.c_double
        L.orc_draw_to_spacing.argtypes = [C.c_double, C.c_int32, C.c_void_p]
        L.orc_ensemble_create.restype = C.c_void_p
        L.orc_ensemble_create.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        L.orc_ensemble_time_steps.restype = C.c_double
        L.orc_ensemble_time_steps.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_ensemble_destroy.argtypes = [C.c_void_p]
        L.orc_ensemble_kick.argtypes = [C.c_void_p]
        L.orc_set_u_frame.argtypes = [C.c_void_p, C.c_double]
        L.orc_set_t.argtypes = [C.c_void_p, C.c_double]
        L.orc_set_inc.argtypes = [C.c_void_p, C.c_int64]
        L.orc_time_steps.argtypes = [C.c_void_p, C.c_int64]
        L.orc_time_steps_fused.argtypes = [C.c_void_p, C.c_int64]
        L.orc_ensemble_configure.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_flow_steps.argtypes = [C.c_void_p, C.c_int64, C.c_double]
        L.orc_time_steps_until_event.argtypes = [
            C.c_void_p, C.c_double, C.c_int64, C.c_int64, C.c_void_p]
        L.orc_minimise.argtypes = [
            C.c_void_p, C.c_double, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p]
        L.orc_minimise_truncate.argtypes = [
            C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_int64, C.c_int64,
            C.c_int, C.c_int, C.c_void_p]
        L.orc_max_uniform_displacement.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_event_driven_step.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_void_p]
        L.orc_trigger.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_int]
        L.orc_advance_to_fixed_force.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.orc_chunk_yield.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.orc_pcg32_draws.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_void_p]
        L.orc_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        for name in ("orc_set_u", "orc_set_v", "orc_set_a", "orc_chunk_index_at_align",
                     "orc_chunk_left_of_align", "orc_chunk_right_of_align"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
        L.orc_chunk_state_at.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_chunk_restore.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        for name in ("orc_refresh", "orc_quench", "orc_destroy"):
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("orc_size", "orc_u_frame", "orc_inc", "orc_residual", "orc_temperature",
                     "orc_qs_first", "orc_qs_last"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_create.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_create_thermal.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_uint64,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_thermal_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_thermal_set.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_pcg32_normal.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_double, C.c_double,
                                       C.c_void_p]
        L.orc_pcg32_randint.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_uint32, C.c_void_p]
        L.orc_erf_inv.argtypes = [C.c_double]
        L.orc_erf_inv.restype = C.c_double
        L.orc_gamma_p_inv.argtypes = [C.c_double, C.c_double]
        L.orc_gamma_p_inv.restype = C.c_double
        _LIB = L
    return _LIB


def make_params(potential, interactions, minimisation, shape, m, eta, mu, kappa, k1, k2, k_frame,
                dt, seed, distribution, parameters, offset, nchunk) -> Params:
    if distribution not in DIST:
        raise RuntimeError("Unknown distribution: " + str(distribution))  # detail.h:65
    p = Params()
    p.potential = POT[potential]
    p.interactions = INT[interactions]
    p.minimisation = int(minimisation)
    shape = [int(i) for i in shape]
    p.rank = len(shape)
    p.shape[0] = shape[0]
    p.shape[1] = shape[1] if len(shape) == 2 else 1
    p.m, p.eta, p.mu, p.kappa = float(m), float(eta), float(mu), float(kappa)
    p.k1, p.k2, p.k_frame, p.dt = float(k1), float(k2), float(k_frame), float(dt)
    p.seed = int(seed)
    p.distribution = DIST[distribution]
    parameters = [float(i) for i in parameters]
    p.nparameters = len(parameters)
    for k, val in enumerate(parameters[:4]):
        p.parameters[k] = val
    p.offset = float(offset)
    p.nchunk = int(nchunk)
    return p


def pcg32_draws(initstate: int, n: int, initseq: int = 0) -> np.ndarray:
    out = np.empty(n, dtype=np.float64)
    lib().orc_pcg32_draws(int(initstate), int(initseq), n, out.ctypes.data)
    return out


PCG32_INITSEQ = 0xDA3E39CB94B95BDB  # prrng's default stream (App. A.1)


def pcg32_normal(initstate: int, n: int, mean=0.0, stddev=1.0, initseq: int = PCG32_INITSEQ):
    """prrng.pcg32(initstate).normal([n], mean, stddev)"""
    out = np.empty(n, dtype=np.float64)
    lib().orc_pcg32_normal(int(initstate), int(initseq), n, mean, stddev, out.ctypes.data)
    return out


def pcg32_randint(initstate: int, n: int, high: int, initseq: int = PCG32_INITSEQ):
    """prrng.pcg32(initstate).randint([n], high)"""
    out = np.empty(n, dtype=np.int64)
    lib().orc_pcg32_randint(int(initstate), int(initseq), n, int(high), out.ctypes.data)
    return out


def lower_bound(y: np.ndarray, u: np.ndarray) -> np.ndarray:
    """prrng.lower_bound: per row, i with y[i] < u <= y[i+1]."""
    return np.array([np.searchsorted(y[r], u[r], side="left") - 1 for r in range(y.shape[0])])


class _Chunk:
    """The python-prrng ``pcg32_tensor_cumsum`` surface exercised by the reference's tests
    (/root/reference/tests/test_Line1d.py:83-86,309-326)."""

    _MARGIN = 30
    _BUFFER = 2

    def __init__(self, system):
        self._s = system
        self._start = np.zeros(system.shape, dtype=np.int64)

    def _arr(self, fn, dtype):
        out = np.empty(self._s.shape, dtype=dtype)
        self._s._check(fn(self._s._h, out.ctypes.data))
        return out

    @property
    def index_at_align(self):
        return self._arr(lib().orc_chunk_index_at_align, np.int64)

    @property
    def left_of_align(self):
        return self._arr(lib().orc_chunk_left_of_align, np.float64)

    @property
    def right_of_align(self):
        return self._arr(lib().orc_chunk_right_of_align, np.float64)

    @property
    def start(self):
        i = self.index_at_align
        n = self._s._nchunk
        loc = i - self._start
        move = (loc < self._BUFFER) | (loc >= n - 1 - self._BUFFER)
        self._start = np.where(move, np.maximum(i - self._MARGIN, 0), self._start)
        return self._start.copy()

    @property
    def chunk_index_at_align(self):
        return self.index_at_align - self.start

    @property
    def data(self):
        start = self.start
        n = self._s._nchunk
        N = self._s.size
        out = np.empty((N, n), dtype=np.float64)
        flat = start.ravel()
        if np.all(flat == flat[0]):
            self._s._check(lib().orc_chunk_yield(self._s._h, int(flat[0]), n, out.ctypes.data))
        else:
            lo, hi = int(flat.min()), int(flat.max())
            tmp = np.empty((N, hi - lo + n), dtype=np.float64)
            self._s._check(lib().orc_chunk_yield(self._s._h, lo, hi - lo + n, tmp.ctypes.data))
            for p in range(N):
                out[p] = tmp[p, flat[p] - lo: flat[p] - lo + n]
        return out.reshape(tuple(self._s.shape) + (n,))

    def state_at(self, index):
        index = np.ascontiguousarray(np.broadcast_to(index, self._s.shape), dtype=np.int64)
        out = np.empty(self._s.shape, dtype=np.uint64)
        self._s._check(lib().orc_chunk_state_at(self._s._h, index.ctypes.data, out.ctypes.data))
        return out

    def restore(self, state, value, index):
        state = np.ascontiguousarray(state, dtype=np.uint64)
        value = np.ascontiguousarray(value, dtype=np.float64)
        index = np.ascontiguousarray(index, dtype=np.int64)
        self._s._check(lib().orc_chunk_restore(
            self._s._h, state.ctypes.data, value.ctypes.data, index.ctypes.data))
        self._start = index.reshape(self._s.shape).copy()


class _External:
    """detail::RandomNormalForcing as bound by python/main.cpp:250-275."""

    def __init__(self, system):
        self._s = system

    @property
    def f_thermal(self):
        out = np.empty(self._s._shape, dtype=np.float64)
        self._s._check(lib().orc_thermal_get(self._s._h, out.ctypes.data, None, None))
        return out

    @f_thermal.setter
    def f_thermal(self, arg):
        arg = np.ascontiguousarray(arg, dtype=np.float64)
        if arg.shape != self._s._shape:
            raise RuntimeError("assertion failed (xt::has_shape(f_thermal, m_f_thermal.shape()))")
        self._s._check(lib().orc_thermal_set(self._s._h, arg.ctypes.data, None, None))

    @property
    def next(self):
        out = np.empty(self._s._shape, dtype=np.int64)
        self._s._check(lib().orc_thermal_get(self._s._h, None, out.ctypes.data, None))
        return out

    @next.setter
    def next(self, arg):
        arg = np.ascontiguousarray(arg, dtype=np.int64)
        if arg.shape != self._s._shape:
            raise RuntimeError("assertion failed (xt::has_shape(next, m_next.shape()))")
        self._s._check(lib().orc_thermal_set(self._s._h, None, arg.ctypes.data, None))

    @property
    def state(self):
        out = C.c_uint64()
        self._s._check(lib().orc_thermal_get(self._s._h, None, None, C.byref(out)))
        return out.value

    @state.setter
    def state(self, arg):
        val = C.c_uint64(int(arg))
        self._s._check(lib().orc_thermal_set(self._s._h, None, None, C.byref(val)))


class System:
    """Generic oracle system (detail::System, detail.h:1046-2051)."""

    def __init__(self, potential, interactions, shape, *, m=1.0, eta=0.0, mu=1.0, kappa=0.0,
                 k1=0.0, k2=0.0, k_frame=1.0, dt=0.0, seed=0, distribution="random",
                 parameters=(), offset=-100.0, nchunk=5000, minimisation=0, forcing=None):
        self._par = make_params(potential, interactions, minimisation, shape, m, eta, mu, kappa,
                                k1, k2, k_frame, dt, seed, distribution, parameters, offset, nchunk)
        self._shape = tuple(int(i) for i in shape)
        self._nchunk = int(nchunk)
        self._h = C.c_void_p()
        if forcing is None:
            self._check(lib().orc_create(C.byref(self._par), C.byref(self._h)))
        else:  # Line1d.h:293-319: External = RandomNormalForcing
            mean, stddev, seed_forcing, dinc_init, dinc = forcing
            dinc_init = np.ascontiguousarray(dinc_init, dtype=np.int64)
            dinc = np.ascontiguousarray(dinc, dtype=np.int64)
            assert dinc_init.shape == self._shape and dinc.shape == self._shape
            self._check(lib().orc_create_thermal(
                C.byref(self._par), float(mean), float(stddev), int(seed_forcing),
                dinc_init.ctypes.data, dinc.ctypes.data, C.byref(self._h)))
            self.external = _External(self)
        self._chunk = _Chunk(self)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    @staticmethod
    def _check(rc):
        if rc != 0:
            raise RuntimeError(lib().orc_last_error().decode())

    # ---- parameters
    chunk = property(lambda self: self._chunk)
    size = property(lambda self: int(np.prod(self._shape)))
    shape = property(lambda self: list(self._shape))
    dt = property(lambda self: self._par.dt)
    mu = property(lambda self: self._par.mu)
    eta = property(lambda self: self._par.eta)
    m = property(lambda self: self._par.m)
    k_frame = property(lambda self: self._par.k_frame)

    def _get(self, which):
        out = np.empty(self._shape, dtype=np.float64)
        self._check(lib().orc_get(self._h, which, out.ctypes.data))
        return out

    def _set(self, fn, arg):
        arg = np.ascontiguousarray(arg, dtype=np.float64)
        if arg.shape != self._shape:
            raise RuntimeError("assertion failed (xt::has_shape(arg, m_u.shape()))")
        self._check(fn(self._h, arg.ctypes.data))

    u = property(lambda self: self._get(0), lambda self, x: self._set(lib().orc_set_u, x))
    v = property(lambda self: self._get(1), lambda self, x: self._set(lib().orc_set_v, x))
    a = property(lambda self: self._get(2), lambda self, x: self._set(lib().orc_set_a, x))
    f = property(lambda self: self._get(3))
    f_potential = property(lambda self: self._get(4))
    f_frame = property(lambda self: self._get(5))
    f_interactions = property(lambda self: self._get(6))
    f_damping = property(lambda self: self._get(7))
    inc = property(lambda self: int(lib().orc_inc(self._h)),
                   lambda self, x: self._check(lib().orc_set_inc(self._h, int(x))))
    t = property(lambda self: lib().orc_inc(self._h) * self._par.dt,
                 lambda self, x: self._check(lib().orc_set_t(self._h, float(x))))
    u_frame = property(lambda self: lib().orc_u_frame(self._h),
                       lambda self, x: self._check(lib().orc_set_u_frame(self._h, float(x))))
    temperature = property(lambda self: lib().orc_temperature(self._h))
    residual = property(lambda self: lib().orc_residual(self._h))
    quasistaticActivityFirst = property(lambda self: int(lib().orc_qs_first(self._h)))
    quasistaticActivityLast = property(lambda self: int(lib().orc_qs_last(self._h)))

    def refresh(self):
        self._check(lib().orc_refresh(self._h))

    def quench(self):
        self._check(lib().orc_quench(self._h))

    def maxUniformDisplacement(self, direction=1):
        out = C.c_double()
        self._check(lib().orc_max_uniform_displacement(self._h, int(direction), C.byref(out)))
        return out.value

    def trigger(self, p, eps, direction=1):
        self._check(lib().orc_trigger(self._h, int(p), float(eps), int(direction)))

    def advanceToFixedForce(self, f_frame, allow_plastic=False):
        self._check(lib().orc_advance_to_fixed_force(self._h, float(f_frame), int(allow_plastic)))

    def minimise(self, tol=1e-5, niter_tol=10, max_iter=int(1e9), time_activity=False,
                 max_iter_is_error=True):
        ret = C.c_int64()
        self._check(lib().orc_minimise(self._h, tol, int(niter_tol), int(max_iter),
                                       int(time_activity), int(max_iter_is_error), C.byref(ret)))
        return ret.value

    def minimise_truncate(self, i_n, A_truncate=0, S_truncate=0, tol=1e-5, niter_tol=10,
                          max_iter=int(1e9), time_activity=True, max_iter_is_error=True):
        i_n = np.ascontiguousarray(i_n, dtype=np.int64)
        ret = C.c_int64()
        self._check(lib().orc_minimise_truncate(
            self._h, i_n.ctypes.data, int(A_truncate), int(S_truncate), tol, int(niter_tol),
            int(max_iter), int(time_activity), int(max_iter_is_error), C.byref(ret)))
        return ret.value

    def eventDrivenStep(self, eps, kick, direction=1):
        out = C.c_double()
        self._check(lib().orc_event_driven_step(self._h, float(eps), int(bool(kick)),
                                                int(direction), C.byref(out)))
        return out.value

    def timeStep(self):
        self._check(lib().orc_time_steps(self._h, 1))

    def timeSteps(self, n):
        self._check(lib().orc_time_steps(self._h, int(n)))

    def timeSteps_fused(self, n):
        """The same steps through the fused single-pass CPU flavour (fqsb_oracle.c:
        orc_time_steps_fused): bit-identical results, best-case CPU cost."""
        self._check(lib().orc_time_steps_fused(self._h, int(n)))

    def timeStepsUntilEvent(self, tol=1e-5, niter_tol=10, max_iter=int(1e9)):
        ret = C.c_int64()
        self._check(lib().orc_time_steps_until_event(self._h, tol, int(niter_tol), int(max_iter),
                                                     C.byref(ret)))
        return ret.value

    def flowSteps(self, n, v_frame):
        self._check(lib().orc_flow_steps(self._h, int(n), float(v_frame)))


class _Namespace:
    pass


def _common(kw):
    return dict(seed=kw.pop("seed"), distribution=kw.pop("distribution"),
                parameters=kw.pop("parameters"), offset=kw.pop("offset", -100.0),
                nchunk=kw.pop("nchunk", 5000))


Line1d = _Namespace()
Line2d = _Namespace()
Particles = _Namespace()


def _mk(potential, interactions, k1name=None, k2name=None, kappa=False, minimisation=0,
        forcing=False):
    def ctor(**kw):
        kw = dict(kw)
        com = _common(kw)
        if forcing:
            com["forcing"] = (kw.pop("mean"), kw.pop("stddev"), kw.pop("seed_forcing"),
                              kw.pop("dinc_init"), kw.pop("dinc"))
        k1 = kw.pop(k1name) if k1name else 0.0
        k2 = kw.pop(k2name) if k2name else 0.0
        kap = kw.pop("kappa") if kappa else 0.0
        shape = kw.pop("shape")
        return System(potential, interactions, shape, m=kw.pop("m", 1.0), eta=kw.pop("eta", 0.0),
                      mu=kw.pop("mu"), kappa=kap, k1=k1, k2=k2, k_frame=kw.pop("k_frame"),
                      dt=kw.pop("dt", 0.0), minimisation=minimisation, **com, **kw)
    return ctor


# Line1d.h:112-677
Line1d.System_Cuspy_Laplace = _mk("Cuspy", "Laplace1d", "k_interactions")
Line1d.System_Cuspy_Laplace_Nopassing = _mk("Cuspy", "Laplace1d", "k_interactions", minimisation=1)
Line1d.System_SemiSmooth_Laplace = _mk("SemiSmooth", "Laplace1d", "k_interactions", kappa=True)
Line1d.System_Smooth_Laplace = _mk("Smooth", "Laplace1d", "k_interactions")
Line1d.System_Cuspy_Quartic = _mk("Cuspy", "Quartic1d", "a1", "a2")
Line1d.System_Cuspy_QuarticGradient = _mk("Cuspy", "QuarticGradient1d", "k2", "k4")
Line1d.System_Cuspy_LongRange = _mk("Cuspy", "LongRange1d", "k_interactions", "alpha")
# Line1d.h:261-330, 486-556 (thermal: External = RandomNormalForcing, Minimisation = None)
Line1d.System_Cuspy_Laplace_RandomForcing = _mk("Cuspy", "Laplace1d", "k_interactions",
                                                minimisation=2, forcing=True)
Line1d.System_Cuspy_Quartic_RandomForcing = _mk("Cuspy", "Quartic1d", "a1", "a2",
                                                minimisation=2, forcing=True)
# Particles.h:93-135 (no interactions)
Particles.System_Cuspy = _mk("Cuspy", "None")
Particles.System_Cuspy_RandomForcing = _mk("Cuspy", "None", minimisation=2, forcing=True)
Particles.System_SemiSmooth = _mk("SemiSmooth", "None", kappa=True)  # Particles.h:233-270
Particles.System_Smooth = _mk("Smooth", "None")  # Particles.h:276-311
# Line2d.h:77-162 (+ the new 2-D no-passing system)
Line2d.System_Cuspy_Laplace = _mk("Cuspy", "Laplace2d", "k_interactions")
Line2d.System_Cuspy_QuarticGradient = _mk("Cuspy", "QuarticGradient2d", "k2", "k4")
Line2d.System_Cuspy_Laplace_Nopassing = _mk("Cuspy", "Laplace2d", "k_interactions", minimisation=1)


class CpuEnsemble:
    """`nsys` independent oracle lines advanced by a thread pool (bench.py's CPU arm)."""

    def __init__(self, par: Params, nsys: int, nthreads: int):
        self.nsys, self.nthreads = int(nsys), int(nthreads)
        self._e = lib().orc_ensemble_create(C.byref(par), self.nsys, self.nthreads)

    def time_steps(self, nsteps: int):
        cs = C.c_double()
        sec = lib().orc_ensemble_time_steps(self._e, int(nsteps), C.byref(cs))
        return sec, cs.value

    def configure(self, fused=False, ftz=False):
        """fused: timeSteps as the single-pass flavour; ftz: flush denormals (timing only)."""
        lib().orc_ensemble_configure(self._e, int(fused), int(ftz))

    def kick(self):
        """eventDrivenStep(1e-3, False) + eventDrivenStep(1e-3, True) on every line."""
        lib().orc_ensemble_kick(self._e)

    def __del__(self):
        if getattr(self, "_e", None):
            lib().orc_ensemble_destroy(self._e)
            self._e = None
