"""Host-side mirror of the reference's system classes over the C ABI.

``System`` reproduces the Python surface bound in /root/reference/python/main.cpp:46-226
(``mySystemNd`` + ``mySystemNdAthermal`` + ``mySystemNdDynamics``) for one realisation;
``Ensemble`` is the same surface over ``nrealisations`` independent realisations in one handle
(new: the reference has no batch class). No arithmetic happens here.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import lib, check


def _params(potential, interactions, minimisation, shape, m, eta, mu, kappa, k1, k2, k_frame, dt,
            seed, distribution, parameters, offset, nchunk, nrealisations, seed_stride, device,
            kernel, seed_first=0, seed_period=0):
    if distribution not in _capi.DIST:
        raise RuntimeError("Unknown distribution: " + str(distribution))  # detail.h:65
    p = _capi.Params()
    p.potential = _capi.POT[potential]
    p.interactions = _capi.INT[interactions]
    p.minimisation = int(minimisation)
    shape = [int(i) for i in shape]
    p.rank = len(shape)
    p.shape[0] = shape[0]
    p.shape[1] = shape[1] if len(shape) == 2 else 1
    p.m, p.eta, p.mu, p.kappa = float(m), float(eta), float(mu), float(kappa)
    p.k1, p.k2, p.k_frame, p.dt = float(k1), float(k2), float(k_frame), float(dt)
    p.seed = int(seed)
    p.distribution = _capi.DIST[distribution]
    parameters = [float(i) for i in parameters]
    if len(parameters) > 4:
        raise RuntimeError("at most 4 distribution parameters")
    p.nparameters = len(parameters)
    for k, val in enumerate(parameters[:4]):
        p.parameters[k] = val
    p.offset = float(offset)
    p.nchunk = int(nchunk)
    p.nrealisations = int(nrealisations)
    p.seed_stride = int(seed_stride)
    p.device = int(device)
    p.kernel = int(kernel)
    p.seed_first = int(seed_first)
    p.seed_period = int(seed_period)
    return p


class _Chunk:
    """``system.chunk``: the python-prrng ``pcg32_tensor_cumsum`` surface the reference exposes
    (main.cpp:65-70) as exercised by tests/test_Line1d.py:83-86,309-326. The device keeps no
    chunk (only the current well and the generator state), so ``data`` / ``start`` are a window
    of ``nchunk`` yield positions regenerated on demand."""

    _MARGIN = 30  # prrng::alignment(buffer=2, margin=30, ...), Line1d.h:154
    _BUFFER = 2

    def __init__(self, owner):
        self._o = owner
        self._start = np.zeros(owner._full_shape, dtype=np.int64)

    def _arr(self, fn, dtype):
        o = self._o
        out = np.empty(o._full_shape, dtype=dtype)
        check(fn(o._h, out.ctypes.data, out.size))
        return o._squeeze(out)

    @property
    def index_at_align(self):
        return self._arr(lib.fqsb_chunk_index_at_align, np.int64)

    @property
    def left_of_align(self):
        return self._arr(lib.fqsb_chunk_left_of_align, np.float64)

    @property
    def right_of_align(self):
        return self._arr(lib.fqsb_chunk_right_of_align, np.float64)

    def _update_start(self):
        o = self._o
        i = np.empty(o._full_shape, dtype=np.int64)
        check(lib.fqsb_chunk_index_at_align(o._h, i.ctypes.data, i.size))
        loc = i - self._start
        move = (loc < self._BUFFER) | (loc >= o._nchunk - 1 - self._BUFFER)
        self._start = np.where(move, np.maximum(i - self._MARGIN, 0), self._start)
        return i

    @property
    def start(self):
        self._update_start()
        return self._o._squeeze(self._start.copy())

    @property
    def chunk_index_at_align(self):
        i = self._update_start()
        return self._o._squeeze(i - self._start)

    @property
    def chunk_size(self):
        return self._o._nchunk

    @property
    def data(self):
        o = self._o
        self._update_start()
        first = np.ascontiguousarray(self._start, dtype=np.int64)
        out = np.empty(o._full_shape + (o._nchunk,), dtype=np.float64)
        check(lib.fqsb_chunk_data(o._h, first.ctypes.data, o._nchunk, out.ctypes.data))
        return o._squeeze(out)

    def align(self, u):
        """Align the landscape with positions ``u`` (python-prrng ``align``): afterwards
        ``index_at_align`` / ``left_of_align`` / ``right_of_align`` refer to ``u``. The system's
        own slips are untouched (its next ``updated_u`` re-aligns to them)."""
        o = self._o
        u = np.ascontiguousarray(u, dtype=np.float64)
        if u.shape != o._user_shape:
            raise RuntimeError("assertion failed (xt::has_shape(u, m_u.shape()))")
        check(lib.fqsb_chunk_align(o._h, u.ctypes.data, u.size))

    @property
    def generators(self):
        """The per-block pcg32 generators (python-prrng ``pcg32_array`` as far as the systems use
        it): ``initstate``, ``initseq``, and ``state()`` = the state after the last entry of the
        current window, i.e. where prrng's own generators stand."""
        return _Generators(self)

    def state_at(self, index):
        o = self._o
        index = np.ascontiguousarray(np.broadcast_to(index, o._user_shape).reshape(o._full_shape),
                                     dtype=np.int64)
        out = np.empty(o._full_shape, dtype=np.uint64)
        check(lib.fqsb_chunk_state_at(o._h, index.ctypes.data, out.ctypes.data, out.size))
        return o._squeeze(out)

    def restore(self, state, value, index):
        o = self._o
        state = np.ascontiguousarray(np.asarray(state, dtype=np.uint64).reshape(o._full_shape))
        value = np.ascontiguousarray(np.asarray(value, dtype=np.float64).reshape(o._full_shape))
        index = np.ascontiguousarray(np.asarray(index, dtype=np.int64).reshape(o._full_shape))
        check(lib.fqsb_chunk_restore(o._h, state.ctypes.data, value.ctypes.data,
                                     index.ctypes.data, state.size))
        self._start = index.copy()


class _Generators:
    """``system.chunk.generators``: see :attr:`_Chunk.generators`."""

    def __init__(self, chunk):
        self._c = chunk

    @property
    def shape(self):
        return list(self._c._o._user_shape)

    @property
    def size(self):
        return int(np.prod(self._c._o._user_shape))

    @property
    def initstate(self):
        """seed + flat block index (+ realisation * seed_stride), Line1d.h:148-151."""
        o = self._c._o
        n = int(np.prod(o._shape))
        stride = int(o._par.seed_stride) or n
        r = np.arange(o._R, dtype=np.uint64).reshape((o._R,) + (1,) * len(o._shape))
        p = np.arange(n, dtype=np.uint64).reshape(o._shape)
        if int(o._par.seed_period):
            p = (np.uint64(int(o._par.seed_first)) + p) % np.uint64(int(o._par.seed_period))
        return o._squeeze(np.uint64(int(o._par.seed)) + r * np.uint64(stride) + p)

    @property
    def initseq(self):
        return np.zeros(self._c._o._user_shape, dtype=np.uint64)  # Line1d.h:151

    def state(self):
        c = self._c
        start = c.start
        return c.state_at(start + c.chunk_size)


class _External:
    """``system.external``: the ``detail::RandomNormalForcing`` object of a thermal system
    (main.cpp:250-275: ``state``, ``f_thermal``, ``next``)."""

    def __init__(self, owner):
        self._o = owner

    def _get(self, fn, dtype):
        o = self._o
        out = np.empty(o._full_shape, dtype=dtype)
        check(fn(o._h, out.ctypes.data, out.size))
        return o._squeeze(out)

    def _set(self, fn, arg, dtype, text):
        o = self._o
        arg = np.ascontiguousarray(arg, dtype=dtype)
        if arg.shape != o._user_shape:  # detail.h:978,997
            raise RuntimeError(f"assertion failed (xt::has_shape({text}, m_{text}.shape()))")
        check(fn(o._h, arg.ctypes.data, arg.size))

    f_thermal = property(
        lambda self: self._get(lib.fqsb_external_get_f_thermal, np.float64),
        lambda self, x: self._set(lib.fqsb_external_set_f_thermal, x, np.float64, "f_thermal"),
        doc="Random force")
    next = property(
        lambda self: self._get(lib.fqsb_external_get_next, np.int64),
        lambda self, x: self._set(lib.fqsb_external_set_next, x, np.int64, "next"),
        doc="Next draw increment")

    @property
    def state(self):
        """State of RNG"""
        out = np.empty(self._o._R, dtype=np.uint64)
        check(lib.fqsb_external_get_state(self._o._h, out.ctypes.data))
        return int(out[0]) if self._o._squeeze_realisation else out

    @state.setter
    def state(self, arg):
        arg = np.ascontiguousarray(np.broadcast_to(np.asarray(arg, dtype=np.uint64),
                                                   (self._o._R,)))
        check(lib.fqsb_external_set_state(self._o._h, arg.ctypes.data))

    def __repr__(self):
        return "<FrictionQPotSpringBlock.detail.RandomNormalForcing_1>"


class Ensemble:
    """``nrealisations`` independent systems in one device-resident handle.

    Realisation ``r`` is the reference system constructed with ``seed + r * seed_stride``
    (``seed_stride`` defaults to the number of blocks, so the pcg32 initstates are disjoint).
    Array properties have shape ``[nrealisations, *shape]``, scalar properties ``[nrealisations]``.
    """

    _squeeze_realisation = False

    def __init__(self, potential, interactions, shape, *, m=1.0, eta=0.0, mu=1.0, kappa=0.0,
                 k1=0.0, k2=0.0, k_frame=1.0, dt=0.0, seed=0, distribution="random",
                 parameters=(), offset=-100.0, nchunk=5000, minimisation=0, nrealisations=1,
                 seed_stride=0, device=-1, kernel=0, seed_first=0, seed_period=0, forcing=None,
                 seed_forcing_stride=1, contracted=False):
        self._h = C.c_void_p()
        if contracted:  # FQSB_KERNEL_FMA: resident kernels built with FMA contraction (opt-in; the
            kernel = int(kernel) | 0x80  # dynamics then agree with the reference to rounding only)
        self._par = _params(potential, interactions, minimisation, shape, m, eta, mu, kappa, k1,
                            k2, k_frame, dt, seed, distribution, parameters, offset, nchunk,
                            nrealisations, seed_stride, device, kernel, seed_first, seed_period)
        self._shape = tuple(int(i) for i in shape)
        self._R = int(nrealisations)
        self._nchunk = int(nchunk)
        self._full_shape = (self._R,) + self._shape
        self._user_shape = self._shape if self._squeeze_realisation else self._full_shape
        self._kind = (potential, interactions, int(minimisation))
        check(lib.fqsb_create(C.byref(self._par), C.byref(self._h)))
        self._chunk = _Chunk(self)
        if forcing is not None:  # Line1d.h:316-318: External = RandomNormalForcing
            mean, stddev, seed_forcing, dinc_init, dinc = forcing
            arrs = []
            for arr in (dinc_init, dinc):
                arr = np.asarray(arr, dtype=np.int64)
                if arr.shape == self._shape:  # one schedule shared by every realisation
                    arr = np.broadcast_to(arr, self._full_shape)
                if arr.shape != self._full_shape:
                    raise RuntimeError("assertion failed (xt::has_shape(dinc, shape))")
                arrs.append(np.ascontiguousarray(arr))
            check(lib.fqsb_enable_random_forcing(
                self._h, float(mean), float(stddev), int(seed_forcing), int(seed_forcing_stride),
                arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[0].size))
            self._external = _External(self)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.fqsb_destroy(h)
            self._h = None

    # ---- helpers
    def _squeeze(self, arr):
        return arr[0] if self._squeeze_realisation else arr

    def _scalar(self, arr):
        return arr[0].item() if self._squeeze_realisation else arr

    def _get_array(self, name):
        out = np.empty(self._full_shape, dtype=np.float64)
        check(lib.fqsb_get(self._h, _capi.ARRAY[name], out.ctypes.data, out.size))
        return self._squeeze(out)

    def _set_array(self, fn, arg):
        arg = np.ascontiguousarray(arg, dtype=np.float64)
        if arg.shape != self._user_shape:
            # detail.h:1278 FRICTIONQPOTSPRINGBLOCK_ASSERT(xt::has_shape(arg, m_u.shape()))
            raise RuntimeError("assertion failed (xt::has_shape(arg, m_u.shape()))")
        check(fn(self._h, arg.ctypes.data, arg.size))

    def _get_scalars(self, fn, dtype=np.float64):
        out = np.empty(self._R, dtype=dtype)
        check(fn(self._h, out.ctypes.data))
        return self._scalar(out)

    def _set_scalars(self, fn, arg, dtype=np.float64):
        arg = np.ascontiguousarray(np.broadcast_to(np.asarray(arg, dtype=dtype), (self._R,)))
        check(fn(self._h, arg.ctypes.data))

    # ---- deprecated names (main.cpp:49-63)
    @property
    def x(self):
        raise RuntimeError("Deprecated, use 'u'")

    @property
    def x_frame(self):
        raise RuntimeError("Deprecated, use 'u_frame'")

    @property
    def f_neighbours(self):
        raise RuntimeError("Deprecated, use 'f_interactions'")

    @property
    def external(self):
        """Class adding external force (main.cpp:207); thermal systems only."""
        ext = getattr(self, "_external", None)
        if ext is None:
            raise AttributeError("external: not a RandomForcing system")
        return ext

    # ---- parameters (main.cpp:65-79)
    chunk = property(lambda self: self._chunk, doc="Chunk of random numbers")
    size = property(lambda self: int(np.prod(self._shape)), doc="Number of particles")
    shape = property(lambda self: list(self._shape), doc="Shape of the system")
    nrealisations = property(lambda self: self._R)
    dt = property(lambda self: self._par.dt, doc="Time step (parameter)")
    mu = property(lambda self: self._par.mu, doc="Curvature of each well (parameter)")
    eta = property(lambda self: self._par.eta, doc="Damping coefficient (parameter)")
    m = property(lambda self: self._par.m, doc="Mass of each particle (parameter)")
    k_frame = property(lambda self: self._par.k_frame, doc="Loading frame stiffness (parameter)")

    # ---- state (main.cpp:80-129)
    u = property(lambda self: self._get_array("u"),
                 lambda self, x: self._set_array(lib.fqsb_set_u, x), doc="Particle slip.")
    v = property(lambda self: self._get_array("v"),
                 lambda self, x: self._set_array(lib.fqsb_set_v, x), doc="Particle velocities.")
    a = property(lambda self: self._get_array("a"),
                 lambda self, x: self._set_array(lib.fqsb_set_a, x), doc="Particle accelerations.")
    inc = property(lambda self: self._get_scalars(lib.fqsb_get_inc, np.int64),
                   lambda self, x: self._set_scalars(lib.fqsb_set_inc, x, np.int64))
    t = property(lambda self: self._get_scalars(lib.fqsb_get_t),
                 lambda self, x: self._set_scalars(lib.fqsb_set_t, x))
    u_frame = property(lambda self: self._get_scalars(lib.fqsb_get_u_frame),
                       lambda self, x: self._set_scalars(lib.fqsb_set_u_frame, x))
    f = property(lambda self: self._get_array("f"), doc="Residual forces")
    f_potential = property(lambda self: self._get_array("f_potential"), doc="Elastic forces")
    f_frame = property(lambda self: self._get_array("f_frame"), doc="Frame forces")
    f_interactions = property(lambda self: self._get_array("f_interactions"))
    f_damping = property(lambda self: self._get_array("f_damping"))
    temperature = property(lambda self: self._get_scalars(lib.fqsb_temperature))
    residual = property(lambda self: self._get_scalars(lib.fqsb_residual))
    mean_f_frame = property(lambda self: self._get_scalars(lib.fqsb_mean_f_frame),
                            doc="np.mean(f_frame) per realisation, reduced on the device")

    @property
    def quasistaticActivityFirst(self):
        first = np.empty(self._R, dtype=np.int64)
        check(lib.fqsb_qs_activity(self._h, first.ctypes.data, None))
        return self._scalar(first)

    @property
    def quasistaticActivityLast(self):
        last = np.empty(self._R, dtype=np.int64)
        check(lib.fqsb_qs_activity(self._h, None, last.ctypes.data))
        return self._scalar(last)

    def refresh(self):
        check(lib.fqsb_refresh(self._h))

    def quench(self):
        check(lib.fqsb_quench(self._h))

    def maxUniformDisplacement(self, direction=1):
        out = np.empty(self._R, dtype=np.float64)
        check(lib.fqsb_max_uniform_displacement(self._h, int(direction), out.ctypes.data))
        return self._scalar(out)

    def trigger(self, p, eps, direction=1, realisation=0):
        check(lib.fqsb_trigger(self._h, int(realisation), int(p), float(eps), int(direction)))

    def advanceToFixedForce(self, f_frame, allow_plastic=False):
        f = np.ascontiguousarray(np.broadcast_to(np.asarray(f_frame, dtype=np.float64),
                                                 (self._R,)))
        check(lib.fqsb_advance_to_fixed_force(self._h, f.ctypes.data, int(allow_plastic)))

    # ---- athermal protocol (main.cpp:153-203)
    def minimise(self, tol=1e-5, niter_tol=10, max_iter=int(1e9), time_activity=False,
                 max_iter_is_error=True):
        ret = np.empty(self._R, dtype=np.int64)
        check(lib.fqsb_minimise(self._h, float(tol), int(niter_tol), int(max_iter),
                                int(time_activity), int(max_iter_is_error), ret.ctypes.data))
        return self._scalar(ret)

    def minimise_truncate(self, i_n, A_truncate=0, S_truncate=0, tol=1e-5, niter_tol=10,
                          max_iter=int(1e9), time_activity=True, max_iter_is_error=True):
        i_n = np.ascontiguousarray(i_n, dtype=np.int64)
        if i_n.shape != self._user_shape:
            raise RuntimeError("assertion failed (xt::has_shape(i_n, m_u.shape()))")
        ret = np.empty(self._R, dtype=np.int64)
        check(lib.fqsb_minimise_truncate(
            self._h, i_n.ctypes.data, int(A_truncate), int(S_truncate), float(tol),
            int(niter_tol), int(max_iter), int(time_activity), int(max_iter_is_error),
            ret.ctypes.data))
        return self._scalar(ret)

    def eventDrivenStep(self, eps, kick, direction=1):
        out = np.empty(self._R, dtype=np.float64)
        check(lib.fqsb_event_driven_step(self._h, float(eps), int(bool(kick)), int(direction),
                                         out.ctypes.data))
        return self._scalar(out)

    def avalanche(self, i_n):
        """(S, A) per realisation since ``i_n``: S = sum(i - i_n), A = #(i != i_n), reduced on
        the device (what examples/Line1d_Cuspy_Laplace.py:61 computes on the host)."""
        i_n = np.ascontiguousarray(i_n, dtype=np.int64)
        if i_n.shape != self._user_shape:
            raise RuntimeError("assertion failed (xt::has_shape(i_n, m_u.shape()))")
        S = np.empty(self._R, dtype=np.int64)
        A = np.empty(self._R, dtype=np.int64)
        check(lib.fqsb_avalanche(self._h, i_n.ctypes.data, i_n.size, S.ctypes.data,
                                 A.ctypes.data))
        return self._scalar(S), self._scalar(A)

    def mark_indices(self):
        """Keep ``i_n = system.chunk.index_at_align`` on the device:
        the start of an event, against which :meth:`avalanche_since_mark` reduces S and A."""
        check(lib.fqsb_mark_indices(self._h))

    def avalanche_since_mark(self):
        """(S, A) per realisation since :meth:`mark_indices`, reduced on the device: nothing of
        size R*N crosses PCIe (examples/Line1d_Cuspy_Laplace.py:59-61 without the host arrays)."""
        S = np.empty(self._R, dtype=np.int64)
        A = np.empty(self._R, dtype=np.int64)
        check(lib.fqsb_avalanche_since_mark(self._h, S.ctypes.data, A.ctypes.data))
        return self._scalar(S), self._scalar(A)

    def event_record(self):
        """(sum |i - i_n|, #(i != i_n), quasistaticActivityFirst, quasistaticActivityLast) of the
        last ``minimise(time_activity=True)`` / ``minimise_truncate`` (detail.h:1768-1778)."""
        out = [np.empty(self._R, dtype=np.int64) for _ in range(4)]
        check(lib.fqsb_event_record(self._h, *[o.ctypes.data for o in out]))
        return tuple(self._scalar(o) for o in out)

    # ---- dynamics (main.cpp:205-226)
    def _no_dynamics(self):
        if self._kind[2] == 1:  # Line1d.h:231-234
            raise AttributeError("the no-passing system has no dynamics")

    def timeStep(self):
        self._no_dynamics()
        check(lib.fqsb_time_steps(self._h, 1))

    def timeSteps(self, n):
        self._no_dynamics()
        check(lib.fqsb_time_steps(self._h, int(n)))

    def run_from_host(self, n, u=None, v=None, a=None, out_u=None, out_v=None, out_a=None,
                      mean_f_frame=None):
        """``self.u = u; self.v = v; self.a = a; self.timeSteps(n)`` and the read-back of
        ``u, v, a`` / ``np.mean(f_frame)`` per realisation as ONE call whose host<->device copies
        overlap the kernels (chunks of realisations on internal streams of the handle). Arrays
        are C-contiguous float64 of the ensemble's shape (pinned memory gives real overlap);
        ``None`` keeps the current array / skips the read-back."""
        self._no_dynamics()

        def ptr(x, shape, writable):
            if x is None:
                return None
            if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous
                    and x.shape == shape and (x.flags.writeable or not writable)):
                raise RuntimeError("assertion failed (xt::has_shape(arg, m_u.shape()))")
            return x.ctypes.data

        sh = self._user_shape
        check(lib.fqsb_run_from_host(
            self._h, ptr(u, sh, False), ptr(v, sh, False), ptr(a, sh, False),
            int(np.prod(self._full_shape)), int(n), ptr(out_u, sh, True), ptr(out_v, sh, True),
            ptr(out_a, sh, True),
            ptr(mean_f_frame, () if self._squeeze_realisation else (self._R,), True)))

    def timeStepsUntilEvent(self, tol=1e-5, niter_tol=10, max_iter=int(1e9)):
        self._no_dynamics()
        ret = np.empty(self._R, dtype=np.int64)
        check(lib.fqsb_time_steps_until_event(self._h, float(tol), int(niter_tol), int(max_iter),
                                              ret.ctypes.data))
        return self._scalar(ret)

    def flowSteps(self, n, v_frame):
        self._no_dynamics()
        check(lib.fqsb_flow_steps(self._h, int(n), float(v_frame)))

    # ---- instrumentation
    @property
    def launch_count(self):
        return int(lib.fqsb_launch_count(self._h))

    @property
    def step_count(self):
        return int(lib.fqsb_step_count(self._h))

    @property
    def last_kernel(self):
        return lib.fqsb_last_kernel(self._h).decode()

    @property
    def last_kernel_seconds(self):
        return float(lib.fqsb_last_kernel_seconds(self._h))

    @property
    def last_kernel_launches(self):
        return int(lib.fqsb_last_kernel_launches(self._h))

    def set_stream(self, cuda_stream: int):
        check(lib.fqsb_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def __repr__(self):
        pot, inter, mini = self._kind
        return (f"<frictionqpotspringblock_b200 {pot}/{inter}"
                f"{'/Nopassing' if mini == 1 else ('/RandomForcing' if mini == 2 else '')} "
                f"shape={list(self._shape)} "
                f"nrealisations={self._R}>")


class System(Ensemble):
    """One realisation with the reference's exact (unbatched) Python surface."""

    _squeeze_realisation = True
