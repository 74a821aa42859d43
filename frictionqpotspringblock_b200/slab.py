"""Slab domain decomposition of ONE very large line (Line1d) or interface (Line2d) over several
GPUs (BASELINE configs #3 and #5) -- a thin mirror of the ``fqsb_slab_*``
entry points of ``include/fqsb.h``. Everything that moves data or takes a decision lives in
``libfqsb.so``: halo rows travel as NVLink peer stores from one GPU's kernel into its neighbours'
memory, the per-step stop decision of the reference (detail.h:1754-1785) is replayed per batch on
rank-ordered global sums. No ``torch`` here.

Two deployments of the same kernels:

* one process, several GPUs::

      s = SlabSystem("Line2d", "System_Cuspy_Laplace", devices=[0, 1, 2, 3], halo=32, **kw)

* one process per GPU (``torchrun``): every rank builds its member and hands over a callable that
  all-gathers a ``bytes`` object over the ranks (e.g. ``distributed.allgather_bytes``); it is used
  once, for the 64-byte CUDA IPC handles of the mailboxes::

      s = SlabSystem("Line2d", "System_Cuspy_Laplace", rank=r, world=G, device=local_rank,
                     allgather=allgather_bytes, halo=32, **kw)

Scheme: member g owns a contiguous range of rows and integrates them extended by ``halo`` rows
per side that mirror its neighbours' rows. A batch of ``k <= halo`` steps needs no communication
(the local array is treated as periodic by the unchanged kernels; the garbage entering through
the outermost rows travels one row per step and has not reached the owned rows after k steps).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import Line1d, Line2d
from ._capi import check, lib
from .distributed import shard_realisations

NLOG = 5  # fqsb_device.cuh: FQSB_NLOG

# nearest-neighbour stencils with per-block disorder: the only systems a halo of rows is exact for.
# LongRange couples every pair of blocks; the *_RandomForcing classes consume one ordered pcg32
# stream per realisation.
SUPPORTED = {
    "Line1d": ("System_Cuspy_Laplace", "System_Cuspy_Laplace_Nopassing", "System_SemiSmooth_Laplace",
               "System_Smooth_Laplace", "System_Cuspy_Quartic", "System_Cuspy_QuarticGradient"),
    "Line2d": ("System_Cuspy_Laplace", "System_Cuspy_Laplace_Nopassing",
               "System_Cuspy_QuarticGradient"),
}


# ---- pure host logic (unit-tested on CPU) ---------------------------------------------------------
def slab_plan(shape, rank: int, world: int, halo: int) -> dict:
    """Rows and seeds of member ``rank``: it owns rows [lo, lo + cnt) of the global system and
    holds ``cnt + 2 * halo`` local rows starting at global row ``lo - halo`` (periodic)."""
    shape = [int(i) for i in shape]
    rows_total = shape[0]
    unit = int(np.prod(shape[1:])) if len(shape) > 1 else 1
    lo, cnt = shard_realisations(rows_total, rank, world)
    if halo < 1 or cnt < halo:
        raise ValueError("every member must own at least `halo` rows")
    first_row = (lo - halo) % rows_total
    return dict(lo=lo, cnt=cnt, unit=unit, local_shape=[cnt + 2 * halo] + shape[1:],
                seed_first=first_row * unit, seed_period=rows_total * unit,
                halo_cells=halo * unit, own=(halo * unit, (halo + cnt) * unit))


class StopList:
    """GooseFEM::Iterate::StopList in the (num, den) form of the library:
    residual_k^2 = num_k / den_k, residuals are only ever compared, by cross-multiplication."""

    def __init__(self, n: int):
        self.num = np.full(int(n), np.inf)
        self.den = np.ones(int(n))


def first_stop(log: np.ndarray, ring: StopList, tol: float) -> int:
    """Replay the per-step decisions (detail.h:1764-1784) over a batch log [k][NLOG] with the
    library's own host code; returns the 1-based step at which the criterion fires, 0 if none."""
    log = np.ascontiguousarray(log, dtype=np.float64)
    stop = int(lib.fqsb_slab_first_stop(log.ctypes.data, log.shape[0], float(tol), ring.num.size,
                                        ring.num.ctypes.data, ring.den.ctypes.data))
    if stop < 0:
        raise RuntimeError("NaN entries found")  # detail.h:1568
    return stop


def residual_from_sums(sf: float, sff: float) -> float:
    """detail.h:1512-1520."""
    r_fres, r_fext = np.sqrt(sf), np.sqrt(sff)
    return float(r_fres / r_fext if r_fext != 0.0 else r_fres)


# ---- the decomposed system ------------------------------------------------------------------------
class SlabSystem:
    """``Line1d.System_*`` / ``Line2d.System_*`` of global ``shape`` spread over several GPUs. The
    method names follow the reference; scalars of the whole system (return codes, ``u_frame``,
    ``residual``, S) are identical on every member / rank."""

    def __init__(self, module: str, cls: str, *, halo: int = 32, devices=None, rank: int = 0,
                 world: int = 1, device: int = -1, allgather=None, batch=None, kernel=None, **kw):
        if cls not in SUPPORTED.get(module, ()):
            raise RuntimeError(f"slab decomposition is not available for {module}.{cls}: it needs a "
                               "nearest-neighbour, athermal system")
        shape = [int(i) for i in kw.pop("shape")]
        self.shape = shape
        self.size = int(np.prod(shape))
        self.halo = int(halo)
        self._overdamped = "Nopassing" in cls
        # Steps per batch. Verlet: after k steps the garbage entering through the outermost halo
        # row has corrupted v, a of halo row k-1 but not yet its position, so the owned rows and
        # their forces are exact for k = halo. Jacobi sweeps: the residual of state k reads the
        # neighbouring halo row AT state k, which is exact only for k <= halo - 1.
        kmax = self.halo - 1 if self._overdamped else self.halo
        # kernel of the members: 1-D dynamic lines take the temporally blocked kernel (one launch =
        # one batch of <= 64 steps, stop decision on the device), everything else streams (one
        # launch per step, CUDA-graph batches). `kernel=2` forces the streaming kernels.
        if kernel is None:
            kernel = 0 if (module == "Line1d" and not self._overdamped) else 2
        self.kernel = int(kernel)
        if (self.kernel & 15) != 2:
            kmax = min(kmax, 64)  # FQSB_BK_MAXSTEPS
        self.batch = kmax if batch is None else int(batch)
        if not 1 <= self.batch <= kmax:
            raise ValueError("batch must be in [1, halo] (halo - 1 for the no-passing sweeps)")
        ns = Line2d if module == "Line2d" else Line1d
        if devices is not None:  # one process drives every member
            self.world = len(devices)
            ranks = list(range(self.world))
            devs = [int(d) for d in devices]
        else:
            self.world = int(world)
            ranks = [int(rank)]
            devs = [int(device)]
        self.ranks = ranks
        self.plans = [slab_plan(shape, r, self.world, self.halo) for r in ranks]
        self.members = []
        for r, d, plan in zip(ranks, devs, self.plans):
            m = getattr(ns, cls)(shape=plan["local_shape"], kernel=self.kernel, device=d,
                                 seed_first=plan["seed_first"], seed_period=plan["seed_period"],
                                 **kw)
            check(lib.fqsb_slab_init(m._h, r, self.world, plan["halo_cells"], self.halo))
            self.members.append(m)
        self._arr = (C.c_void_p * len(self.members))(*[m._h for m in self.members])
        self._n = len(self.members)
        # connect the mailboxes
        if devices is not None:
            locals_ = (C.c_void_p * self.world)(*[m._h for m in self.members])
            for m in self.members:
                check(lib.fqsb_slab_connect(m._h, locals_, None))
        else:
            handle = (C.c_ubyte * 64)()
            check(lib.fqsb_slab_ipc_handle(self.members[0]._h, handle))
            if self.world > 1:
                if allgather is None:
                    raise ValueError("world > 1 needs `allgather` (bytes -> list of bytes)")
                every = allgather(bytes(handle))
                if len(every) != self.world or any(len(b) != 64 for b in every):
                    raise RuntimeError("allgather must return one 64-byte handle per rank")
                blob = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(every))
            else:
                blob = None
            check(lib.fqsb_slab_connect(self.members[0]._h, None, blob))
        self._allgather = allgather
        self.sys = self.members[0]
        check(lib.fqsb_slab_exchange(self._arr, self._n))

    # ---- plumbing
    def _sums(self, what, direction=1):
        out = np.empty(4, dtype=np.float64)
        check(lib.fqsb_slab_sums(self._arr, self._n, int(what), int(direction), out.ctypes.data))
        return out

    def info(self, member=0):
        out = np.zeros(10, dtype=np.int64)
        check(lib.fqsb_slab_info(self.members[member]._h, out.ctypes.data))
        return dict(zip(("rank", "world", "halo_cells", "own_lo", "own_hi", "batches", "redone",
                         "graphs", "discarded", "blocked"), (int(i) for i in out)))

    # ---- reference surface
    @property
    def u_frame(self):
        return self.members[0].u_frame

    @u_frame.setter
    def u_frame(self, x):
        for m in self.members:
            m.u_frame = x

    @property
    def inc(self):
        return self.members[0].inc

    @property
    def residual(self):
        s = self._sums(1)
        return residual_from_sums(s[0], s[1])

    @property
    def mean_f_frame(self):
        return float(self._sums(2)[1] / self.size)

    @property
    def temperature(self):
        return float(0.5 * self.members[0].m * self._sums(2)[0] / self.size)  # detail.h:1502

    @property
    def step_count(self):
        return self.members[0].step_count

    def owned(self, name, member=None):
        """Owned part of a per-block array (``u``, ``v``, ``f`` ...; ``index_at_align`` etc. from
        ``chunk``) of one member, or -- one process driving all members -- of the whole system."""
        def part(m, plan):
            src = m.chunk if name in ("index_at_align", "left_of_align", "right_of_align") else m
            return getattr(src, name).reshape(-1)[plan["own"][0]:plan["own"][1]].copy()

        if member is not None:
            return part(self.members[member], self.plans[member])
        return np.concatenate([part(m, p) for m, p in zip(self.members, self.plans)])

    def index_at_align_owned(self):
        return self.owned("index_at_align")

    def gather(self, values: np.ndarray) -> np.ndarray:
        """Global flat array from the ranks' owned slices (one process per GPU: through the
        ``allgather`` callable; one process: ``owned`` already returns the whole system)."""
        if self._n == self.world:
            return values.copy()
        parts = self._allgather(np.ascontiguousarray(values).tobytes())
        return np.concatenate([np.frombuffer(b, dtype=values.dtype) for b in parts])

    def mark_indices(self):
        """Device-side ``i_n = chunk.index_at_align`` (start of an event)."""
        check(lib.fqsb_slab_mark_indices(self._arr, self._n))

    def avalanche_since_mark(self):
        """Global (S, A) since :meth:`mark_indices`: S = sum(i - i_n), A = #(i != i_n)."""
        s = self._sums(4)
        return int(round(s[0])), int(round(s[1]))

    def timeSteps(self, n: int):
        check(lib.fqsb_slab_time_steps(self._arr, self._n, int(n), self.batch, 0, 0.0))

    def flowSteps(self, n: int, v_frame: float):
        check(lib.fqsb_slab_time_steps(self._arr, self._n, int(n), self.batch, 1, float(v_frame)))

    def minimise(self, tol=1e-5, niter_tol=10, max_iter=int(1e9), max_iter_is_error=True):
        """detail.h:1676-1792 (dynamic or overdamped), decided per batch on identical global sums."""
        ret = C.c_int64(0)
        steps = C.c_int64(0)
        check(lib.fqsb_slab_minimise(self._arr, self._n, float(tol), int(niter_tol), int(max_iter),
                                     self.batch, int(max_iter_is_error), C.byref(ret),
                                     C.byref(steps)))
        self.last_minimise_steps = int(steps.value)
        return int(ret.value)

    def maxUniformDisplacement(self, direction=1):
        s = self._sums(3, direction)
        return 0.0 if s[1] > 0 else float(s[3])

    def eventDrivenStep(self, eps, kick, direction=1):
        """detail.h:1933-1960; the displacement is agreed globally, applied locally (halos move
        with their originals, no exchange needed)."""
        out = C.c_double(0.0)
        check(lib.fqsb_slab_event_driven_step(self._arr, self._n, float(eps), int(bool(kick)),
                                              int(direction), C.byref(out)))
        return float(out.value)
