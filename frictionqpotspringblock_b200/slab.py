"""Slab domain decomposition of ONE very large line (Line1d) or interface (Line2d) over the ranks of
a ``torch.distributed`` group (one process per GPU; SURVEY.md section 8e, BASELINE configs #3/#5).

Communication-avoiding scheme. Rank g owns a contiguous range of rows (cells for a line) and
integrates a local array extended by ``halo`` rows on each side that mirror its neighbours' rows.
A batch of ``k <= halo`` steps needs NO communication: the local array is simply treated as
periodic, the garbage that enters through its outermost rows travels one row per step and has
not reached the owned rows after k steps. Per batch:

1. snapshot the local state (device copy);
2. ``fqsb_logged_steps(k)``: k fused steps, the per-step sums over the OWNED rows
   (sum f^2, sum f_frame^2, well changes) are logged on the device instead of deciding;
3. one all-reduce (SUM) of the k x 5 log; every rank replays the reference's StopList criterion
   (detail.h:1764-1784) on the identical global sums and finds the same stopping step s*;
4. if s* fell inside the batch: roll back and redo exactly s* steps;
5. one halo exchange of the full block state (u, v, a, y_l, y_r, index, pcg32 state) of k rows
   per side (NCCL send/recv on device buffers; gloo stages through the host).

The per-step collective of the naive scheme (8 B per neighbour + 3 scalars, latency-bound) is
replaced by two collectives per k steps. ``torch.distributed`` is plumbing only.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import Line1d, Line2d
from ._capi import check, lib
from .distributed import shard_realisations

NLOG = 5  # fqsb_device.cuh: FQSB_NLOG


# ---- pure host logic (unit-tested on CPU) ---------------------------------------------------------
class StopList:
    """GooseFEM::Iterate::StopList + the criterion of detail.h:1615,1780, in
    the form the device uses (fqsb_device.cuh: ring_stop): entry k is the pair (num, den) with
    residual_k^2 = num / den and residuals are only ever compared, by cross-multiplication."""

    def __init__(self, n: int):
        self.num = np.full(int(n), np.inf)
        self.den = np.ones(int(n))

    def roll_insert(self, sf: float, sff: float):
        self.num[:-1] = self.num[1:]
        self.den[:-1] = self.den[1:]
        self.num[-1] = sf
        self.den[-1] = sff if sff != 0.0 else 1.0  # detail.h:1516-1519

    def stop(self, tol: float) -> bool:
        tol2 = tol * tol
        tol4 = tol2 * tol2
        with np.errstate(invalid="ignore"):
            descending = not bool(np.any(self.num[1:] * self.den[:-1] > self.num[:-1] * self.den[1:]))
            less1 = bool(np.all(self.num < tol2 * self.den))
            less2 = bool(np.all(self.num < tol4 * self.den))
        return (descending and less1) or less2

    def state(self):
        return self.num.copy(), self.den.copy()

    def restore(self, state):
        self.num, self.den = state[0].copy(), state[1].copy()


def residual_from_sums(sf: float, sff: float) -> float:
    """detail.h:1512-1520."""
    r_fres, r_fext = np.sqrt(sf), np.sqrt(sff)
    return r_fres / r_fext if r_fext != 0.0 else r_fres


def first_stop(log: np.ndarray, ring: StopList, tol: float) -> int:
    """Replay the per-step decisions over a batch log [k][NLOG]; returns the 1-based step at which
    the criterion fires (the ring then holds the state at that step), or 0 if it does not."""
    for j in range(log.shape[0]):
        if np.isnan(log[j, 0]):
            raise RuntimeError("NaN entries found")  # detail.h:1568
        ring.roll_insert(log[j, 0], log[j, 1])
        if ring.stop(tol):
            return j + 1
    return 0


def halo_plan(rank: int, world: int):
    """(send order, recv order) of one exchange. Each entry is (peer, which): ``which`` names the
    local rows involved -- sends: "top"/"bottom" owned rows; recvs: "bottom_halo"/"top_halo".
    A rank sends its top rows to the previous rank first, so every rank must receive from its
    NEXT rank first: with world == 2 both messages travel between the same pair and only the
    order tells them apart."""
    prev, nxt = (rank - 1) % world, (rank + 1) % world
    sends = [(prev, "top"), (nxt, "bottom")]
    recvs = [(nxt, "bottom_halo"), (prev, "top_halo")]
    return sends, recvs


def exchange_halos(export_cells, import_cells, layout, rank, world, group=None, device="cpu",
                   buffers=None):
    """Refresh the halo rows from the neighbours' owned rows.

    ``export_cells(first, count, tensor)`` / ``import_cells(first, count, tensor)`` move the packed
    state of ``count`` cells starting at local cell ``first`` to / from a torch int64 tensor of
    7*count words on ``device``. ``layout`` = dict(top=(first,count), bottom=..., top_halo=...,
    bottom_halo=...). ``buffers``: optional dict reused across calls (one tensor per entry)."""
    import torch
    import torch.distributed as dist

    def buf(which):
        if buffers is None:
            return torch.empty(7 * layout[which][1], dtype=torch.int64, device=device)
        if which not in buffers:
            buffers[which] = torch.empty(7 * layout[which][1], dtype=torch.int64, device=device)
        return buffers[which]

    if world == 1:  # periodic wrap onto oneself
        for src, dst in (("top", "bottom_halo"), ("bottom", "top_halo")):
            t = buf(src)
            export_cells(*layout[src], t)
            import_cells(*layout[dst], t)
        return
    sends, recvs = halo_plan(rank, world)
    ops, inbox = [], []
    for peer, which in sends:
        t = buf(which)
        export_cells(*layout[which], t)
        ops.append(dist.P2POp(dist.isend, t, peer, group))
    for peer, which in recvs:
        t = buf(which)
        inbox.append((which, t))
        ops.append(dist.P2POp(dist.irecv, t, peer, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    if str(device).startswith("cuda"):
        # req.wait() only orders torch's current stream after the NCCL transfer; the import
        # kernels run on the handle's own stream, so the host has to wait for the data
        torch.cuda.current_stream().synchronize()
    for which, t in inbox:
        import_cells(*layout[which], t)


# ---- the decomposed system ------------------------------------------------------------------------
class SlabSystem:
    """``Line1d.System_*`` / ``Line2d.System_*`` of global ``shape`` spread over the ranks of
    ``group``. The method names follow the reference; results that are scalars of the whole system
    (return codes, ``u_frame``, ``residual``, S) are identical on every rank."""

    def __init__(self, module: str, cls: str, *, halo: int = 16, group=None, device: int = -1,
                 **kw):
        import torch.distributed as dist

        self.group = group
        self.dist_on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.dist_on else 0
        self.world = dist.get_world_size(group) if self.dist_on else 1
        self.backend = dist.get_backend(group) if self.dist_on else "none"
        shape = [int(i) for i in kw.pop("shape")]
        self.shape = shape
        self.unit = shape[1] if len(shape) == 2 else 1  # cells per row
        self.rows_total = shape[0]
        self.size = int(np.prod(shape))
        self.halo = int(halo)
        self.lo, self.cnt = shard_realisations(self.rows_total, self.rank, self.world)
        if self.cnt < self.halo:
            raise ValueError("every rank must own at least `halo` rows")
        local_rows = self.cnt + 2 * self.halo
        local_shape = [local_rows] + shape[1:]
        first_row = (self.lo - self.halo) % self.rows_total
        ns = Line2d if module == "Line2d" else Line1d
        self.sys = getattr(ns, cls)(shape=local_shape, kernel=2, device=device,
                                    seed_first=first_row * self.unit, seed_period=self.size, **kw)
        self._h = self.sys._h
        self.own = (self.halo * self.unit, (self.halo + self.cnt) * self.unit)
        check(lib.fqsb_set_owned_range(self._h, self.own[0], self.own[1]))
        k = self.halo * self.unit
        self.layout = {
            "top": (self.own[0], k),
            "bottom": (self.own[1] - k, k),
            "top_halo": (0, k),
            "bottom_halo": (self.own[1], k),
        }
        self._mu, self._k_frame = float(kw["mu"]), float(kw["k_frame"])
        self._overdamped = "Nopassing" in cls
        self._buffers = {}
        # Steps per batch. Verlet: after k steps the garbage entering through the outermost halo
        # row has corrupted v,a of halo row k-1 but not yet its position, so the owned rows and
        # their forces are exact for k = halo. Jacobi sweeps: the residual of state k reads the
        # neighbouring halo row AT state k, which is exact only for k <= halo - 1.
        self.batch = self.halo - 1 if self._overdamped else self.halo
        if self.batch < 1:
            raise ValueError("halo too small")
        import torch

        # halo / reduction buffers live on the device unless the group can only move host memory
        self._tdev = (torch.device("cpu") if self.backend == "gloo"
                      else torch.device("cuda", torch.cuda.current_device()))

    # ---- plumbing
    def _export(self, first, count, tensor):
        check(lib.fqsb_export_cells(self._h, first, count, C.c_void_p(tensor.data_ptr()),
                                    int(tensor.is_cuda)))

    def _import(self, first, count, tensor):
        check(lib.fqsb_import_cells(self._h, first, count, C.c_void_p(tensor.data_ptr()),
                                    int(tensor.is_cuda)))

    def exchange(self):
        exchange_halos(self._export, self._import, self.layout, self.rank, self.world,
                       self.group, self._tdev, self._buffers)

    def _allreduce(self, arr: np.ndarray, op: str = "sum") -> np.ndarray:
        if self.world == 1:
            return arr
        import torch
        import torch.distributed as dist

        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(self._tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN if op == "min" else dist.ReduceOp.SUM,
                        group=self.group)
        return t.cpu().numpy()

    def _sums(self, what, direction=1, i_n=None):
        out = np.empty(4, dtype=np.float64)
        ptr = None
        if i_n is not None:
            i_n = np.ascontiguousarray(i_n, dtype=np.int64)
            ptr = i_n.ctypes.data
        check(lib.fqsb_reduce_sums(self._h, what, direction, ptr, out.ctypes.data))
        return out

    def _logged(self, k: int) -> np.ndarray:
        log = np.empty((k, NLOG), dtype=np.float64)
        check(lib.fqsb_logged_steps(self._h, k, log.ctypes.data))
        return log

    def _owned(self, arr):
        return arr.reshape(-1)[self.own[0]:self.own[1]]

    # ---- reference surface
    @property
    def u_frame(self):
        return self.sys.u_frame

    @u_frame.setter
    def u_frame(self, x):
        self.sys.u_frame = x

    @property
    def inc(self):
        return self.sys.inc

    @property
    def residual(self):
        s = self._allreduce(self._sums(1)[:2])
        return float(residual_from_sums(s[0], s[1]))

    @property
    def mean_f_frame(self):
        s = self._allreduce(self._sums(2)[:2])
        return float(s[1] / self.size)

    def index_at_align_owned(self):
        return self._owned(self.sys.chunk.index_at_align).copy()

    def owned(self, name):
        return self._owned(getattr(self.sys, name)).copy()

    def gather(self, values: np.ndarray) -> np.ndarray:
        """Concatenate the ranks' owned slices into the global flat array (every rank gets it)."""
        if self.world == 1:
            return values.copy()
        import torch
        import torch.distributed as dist

        counts = [shard_realisations(self.rows_total, r, self.world)[1] * self.unit
                  for r in range(self.world)]
        width = max(counts)
        pad = np.zeros(width, dtype=values.dtype)
        pad[: values.size] = values
        mine = torch.from_numpy(pad).to(self._tdev)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        return np.concatenate([p.cpu().numpy()[:c] for p, c in zip(parts, counts)])

    def avalanche(self, i_n_owned):
        """Global (S, A) since the owned reference indices ``i_n_owned``."""
        full = self.sys.chunk.index_at_align.reshape(-1).copy()
        full[self.own[0]:self.own[1]] = i_n_owned
        s = self._allreduce(self._sums(4, 1, full)[:2])
        return int(round(s[0])), int(round(s[1]))

    def timeSteps(self, n: int):
        n = int(n)
        while n > 0:
            k = min(n, self.batch)
            self.sys.timeSteps(k)
            self.exchange()
            n -= k

    def minimise(self, tol=1e-5, niter_tol=10, max_iter=int(1e9), max_iter_is_error=True):
        """detail.h:1676-1792 (dynamic or overdamped), decided per batch on identical global sums."""
        if not tol < 1.0:
            raise RuntimeError("assertion failed (tol < 1.0)")
        ring = StopList(niter_tol)
        done = 0
        while done < max_iter:
            k = int(min(self.batch, max_iter - done))
            check(lib.fqsb_snapshot(self._h))
            saved = ring.state()
            log = self._allreduce(self._logged(k))
            stop = first_stop(log, ring, tol)
            if stop:
                if stop < k:  # the criterion fired inside the batch: redo exactly `stop` steps
                    check(lib.fqsb_rollback(self._h))
                    ring.restore(saved)
                    log = self._allreduce(self._logged(stop))
                    assert first_stop(log, ring, tol) == stop
                self.exchange()
                self.sys.quench()
                return 0
            self.exchange()
            done += k
        if max_iter_is_error:
            raise RuntimeError("No convergence found")  # detail.h:1788
        return done + 1

    def maxUniformDisplacement(self, direction=1):
        s = self._sums(3, direction)
        off = self._allreduce(np.array([s[1]]))[0]
        mn = self._allreduce(np.array([s[3]]), "min")[0]
        return 0.0 if off > 0 else float(mn)

    def eventDrivenStep(self, eps, kick, direction=1):
        """detail.h:1933-1960; the displacement is agreed globally, applied locally (halos move
        with their originals, no exchange needed)."""
        if kick:
            du = eps if direction > 0 else -eps
        else:
            d = self.maxUniformDisplacement(direction)
            if d < 0.5 * eps:
                return 0.0
            du = d - 0.5 * eps if direction > 0 else 0.5 * eps - d
        du_frame = du * (self._k_frame + self._mu) / self._k_frame
        a = np.array([du], dtype=np.float64)
        b = np.array([du_frame], dtype=np.float64)
        check(lib.fqsb_advance_uniformly(self._h, a.ctypes.data, b.ctypes.data))
        return du_frame
