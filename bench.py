#!/usr/bin/env python
"""bench.py -- the headline measurement of the B200-native FrictionQPotSpringBlock integrator.

Workload (BASELINE.json configs[1]): Line1d System_Cuspy_Laplace ensemble, 16384 disorder
realisations x N = 4096 blocks, physics of examples/Line1d_Cuspy_Laplace.py. One "step" = one
``timeSteps(T)`` call over the whole ensemble (T = 1000 velocity-Verlet steps), started from a
kicked (avalanching) state. Metric: block-updates/s = realisations x N x T / time.
``--scaling weak`` (default): every rank integrates its own 16384 realisations (disjoint seeds, no
inter-GPU traffic); ``--scaling strong``: 16384 realisations in total, sharded over the ranks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every field).
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

N_BLOCKS = 4096
R_CONFIG = 16384  # realisations of BASELINE configs[1]
ALGO_BYTES_PER_UPDATE = 64.0  # u,v,a read + write (48) + y_left,y_right read (16); SURVEY 8(d)
# FP64-pipe instructions per block-update of k_resident<Cuspy,Laplace1d,unit> (SASS of the loop,
# cross-checked with ncu's instruction counts, profiles/r2g_ncu_resident_fixed.txt: 144 DADD +
# 81 DMUL + 20 DSETP per warp-step of 8 blocks; the reference's evaluation order, no FMA contraction)
FP64_INSTR_PER_UPDATE = 245.0 / 8.0
# DADD/DMUL issue rate a B200 SM sustains in a register-only loop (tools/fp64_peak.cu, measured:
# 1.934 of the nominal 2.0 warp-instructions per SM per clock)
FP64_PRACTICAL_ISSUE = 1.934 / 2.0
METRIC = "block_updates_per_s"
UNIT = "block-updates/s"


def physics(N):
    """examples/Line1d_Cuspy_Laplace.py:18-30 with k_frame = 1/N."""
    return dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, k_interactions=1.0,
                k_frame=1.0 / N, dt=0.1, shape=[N], distribution="random", parameters=[2.0],
                offset=-50)


def measured_peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(gpu_index)],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(",") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                for name, val in zip(names, r[3:7]):
                    if val.strip().lower() == "active":
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)),
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_config(T, scaling):
    """Identical for both arms: the workload both are quoted on. The reference arm integrates a
    bounded SAMPLE of its realisations per step (``cpu_baseline.sample``)."""
    return {
        "workload": "Line1d System_Cuspy_Laplace ensemble (BASELINE configs[1]): "
                    f"{R_CONFIG} realisations x N={N_BLOCKS}"
                    f"{' per GPU' if scaling == 'weak' else ' in total'}, one step = "
                    f"timeSteps({T}) of every realisation after minimise + eventDrivenStep + kick",
        "realisations": R_CONFIG, "blocks": N_BLOCKS, "inner_steps": T,
        "protocol": "minimise(); eventDrivenStep(1e-3, False); eventDrivenStep(1e-3, True); then "
                    "per step timeSteps(inner_steps), no further kicks",
        "physics": "m=1 eta=2sqrt(3)/10 mu=1 k=1 k_frame=1/N dt=0.1 random[2.0] offset=-50",
        "parallelism": "independent realisations sharded across GPUs, no collective",
        "l2": "state per GPU (3.8 GB) is larger than L2, no flush needed",
    }


# ---- CPU arm: the restated reference (oracle port) on the host cores ------------------------------
def cpu_ensemble(nsys, nthreads, fused):
    from oracle import oracle as orc

    kw = physics(N_BLOCKS)
    par = orc.make_params("Cuspy", "Laplace1d", 0, kw["shape"], kw["m"], kw["eta"], kw["mu"], 0.0,
                          kw["k_interactions"], 0.0, kw["k_frame"], kw["dt"], 0, "random",
                          kw["parameters"], kw["offset"], 5000)
    ens = orc.CpuEnsemble(par, nsys, nthreads)  # minimise + eventDrivenStep + kick per line
    # denormals flushed in the workers: an un-kicked line decays into denormal velocities after
    # ~1e4 steps, which x86 executes several times slower -- an artefact that would flatter the GPU
    ens.configure(fused=fused, ftz=True)
    return ens


def cpu_sample_text(nsys, cores, T, flavour):
    return (f"{nsys} of the {R_CONFIG} realisations x N={N_BLOCKS} per step (timeSteps({T}) each, "
            f"same protocol), one realisation per thread on {cores} threads, {flavour} flavour of "
            "oracle/fqsb_oracle.c (-O3 -march=native, denormals flushed)")


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on all host cores, same
    config / protocol / inner step count as the GPU arm, on a bounded sample of the realisations.
    The genuine library cannot be built (xtensor/prrng/GooseFEM absent), so this is the oracle
    port (kind "port"), pinned on the reference's goldens. Rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    T = args.inner
    nsys = 4 * cores
    out = {}
    for flavour in ("faithful", "fused"):
        ens = cpu_ensemble(nsys, cores, flavour == "fused")
        for _ in range(args.warmup):
            ens.time_steps(T)
        total = 0.0
        for _ in range(args.steps):
            sec, _cs = ens.time_steps(T)
            total += sec
        out[flavour] = (nsys * N_BLOCKS * T * args.steps / total, total)
        del ens
    value, total = out["faithful"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(T, args.scaling),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": cpu_sample_text(nsys, cores, T, "faithful (multi-pass, the "
                                                   "reference's cost structure)"),
                         "fused": {"value": out["fused"][0], "unit": UNIT,
                                   "sample": cpu_sample_text(nsys, cores, T, "fused single-pass "
                                                             "(best-case CPU)")}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- slab decomposition: ONE large system over all ranks (BASELINE configs #3, #5) ----------------
def slab_block(F, rank, world, local_rank, barrier, max_over_ranks, quick):
    """Strong scaling of one 4096 x 4096 interface (config #5, Verlet and no-passing) and one line
    of 2^20 blocks (config #3) over the `world` GPUs through fqsb_slab_* (halo rows as NVLink peer
    stores between the ranks' GPUs, mailboxes shared through CUDA IPC). Every case asserts parity
    with the single-handle run on rank 0: S, A of an event and the frame position."""
    from frictionqpotspringblock_b200.distributed import allgather_bytes
    from frictionqpotspringblock_b200.slab import SlabSystem

    out = {"members": world, "transport": "NVLink peer stores + release/acquire epoch flags; "
                                          "CUDA IPC mailboxes; no NCCL on the data path"}

    def protocol(s):
        s.u_frame = 1.0
        assert s.minimise(max_iter=20000) == 0
        s.mark_indices()
        s.eventDrivenStep(1e-3, False)
        s.eventDrivenStep(1e-3, True)
        assert s.minimise(max_iter=20000) == 0
        S, A = s.avalanche_since_mark()
        return int(S), int(A), float(s.u_frame)

    cases = [("config5_line2d_4096_verlet", "Line2d", "System_Cuspy_Laplace", [4096, 4096], 32,
              128 if quick else 512, dict(k_interactions=1.0)),
             ("config5_line2d_4096_nopassing", "Line2d", "System_Cuspy_Laplace_Nopassing",
              [4096, 4096], 33, 0, dict(k_interactions=1.0)),
             ("config3_line1d_2p20_quartic", "Line1d", "System_Cuspy_Quartic", [1 << 20], 64,
              1024 if quick else 4096, dict(a1=1.0, a2=1.0))]
    for name, module, cls, shape, halo, T, extra in cases:
        n = int(np.prod(shape))
        kw = dict(mu=1.0, k_frame=1.0 / n, shape=shape, seed=0, distribution="random",
                  parameters=[2.0], offset=-50, **extra)
        if "Nopassing" not in cls:
            kw.update(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, dt=0.1)
        res = {"shape": shape, "halo_rows": halo}
        try:
            s = SlabSystem(module, cls, halo=halo, rank=rank, world=world, device=local_rank,
                           allgather=allgather_bytes, **kw)
            barrier()
            t0 = time.perf_counter()
            got = protocol(s)
            barrier()
            res["protocol_s"] = max_over_ranks(time.perf_counter() - t0)
            res["event"] = {"S": got[0], "A": got[1], "u_frame": got[2]}
            res["minimise_steps"] = s.last_minimise_steps
            barrier()
            t0 = time.perf_counter()
            ret = s.minimise(tol=1e-300, max_iter=3 * s.batch, max_iter_is_error=False)
            barrier()
            sec = max_over_ranks(time.perf_counter() - t0)
            res["minimise_us_per_step"] = 1e6 * sec / (3 * s.batch)
            assert ret == 3 * s.batch + 1
            if T:
                s.timeSteps(2 * s.batch)
                barrier()
                t0 = time.perf_counter()
                s.timeSteps(T)
                barrier()
                sec = max_over_ranks(time.perf_counter() - t0)
                res.update(us_per_step=1e6 * sec / T, block_updates_per_s=n * T / sec)
            res["kernel"] = s.members[0].last_kernel
            res["info"] = s.info()
            del s
            if rank == 0:  # the same protocol on ONE handle
                one = getattr(getattr(F, module), cls)(device=local_rank, **kw)
                want = protocol(one)
                res["single_handle_event"] = {"S": want[0], "A": want[1], "u_frame": want[2]}
                res["parity"] = bool(want[:2] == got[:2] and
                                     np.isclose(want[2], got[2], rtol=1e-12, atol=0))
                assert res["parity"], (name, want, got)
                if T:
                    one.timeSteps(64)
                    t0 = time.perf_counter()
                    one.timeSteps(T)
                    sec1 = time.perf_counter() - t0
                    res["single_handle_us_per_step"] = 1e6 * sec1 / T
                    res["single_handle_kernel"] = one.last_kernel
                    res["speedup_vs_single_handle"] = res["single_handle_us_per_step"] / res["us_per_step"]
                del one
            barrier()
        except Exception as e:  # pragma: no cover
            res["error"] = f"{type(e).__name__}: {e}"
        out[name] = res
    return out


def other_configs(F, device, peak):
    """Short device-timed measurements of BASELINE configs #3, #4, #5 and of the thermal systems
    (SURVEY section 8f row N4); times are CUDA-event times of the stepping-kernel launches."""
    out = {}
    phys = dict(m=1.0, eta=2.0 * np.sqrt(3.0) / 10.0, mu=1.0, dt=0.1, distribution="random",
                parameters=[2.0], offset=-50, seed=0, device=device)

    def timed_steps(system, n, warm):
        system.timeSteps(warm)
        system.timeSteps(n)
        return system.last_kernel_seconds / n

    try:  # config #3: one Cuspy_Quartic line of 2^20 blocks, temporally blocked kernel (K2b)
        N = 1 << 20
        s = F.Line1d.System_Cuspy_Quartic(a1=1.0, a2=1.0, k_frame=1.0 / N, shape=[N], **phys)
        s.u_frame = 0.5
        sec = timed_steps(s, 2048, 256)
        out["config3_line_2p20_quartic"] = {
            "kernel": s.last_kernel, "us_per_step": 1e6 * sec, "block_updates_per_s": N / sec,
            "algorithmic_GBps": 64.0 * N / sec / 1e9,
            "note": "k steps per pass over memory: the algorithmic 64 B per block-update are not "
                    "DRAM traffic (FP64-pipe bound)"}
        del s
    except Exception as e:  # pragma: no cover
        out["config3_line_2p20_quartic"] = {"error": str(e)}
    try:  # config #4: LongRange alpha = 1.5, N = 8192 x 1024 realisations (DMMA Toeplitz GEMM)
        N, R = 8192, 1024
        s = F.Line1d.Ensemble_Cuspy_LongRange(k_interactions=1.0, alpha=1.5, k_frame=1.0 / N,
                                              shape=[N], nrealisations=R, **phys)
        s.u_frame = np.full(R, 0.5)
        sec = timed_steps(s, 4, 2)
        out["config4_longrange_8192x1024"] = {
            "kernel": s.last_kernel, "ms_per_step": 1e3 * sec,
            "fp64_TFLOPs": 2.0 * N * N * R / sec / 1e12,
            "block_updates_per_s": N * R / sec}
        del s
    except Exception as e:  # pragma: no cover
        out["config4_longrange_8192x1024"] = {"error": str(e)}
    try:  # config #5: 4096 x 4096 interface, Verlet step and no-passing sweeps
        rows = cols = 4096
        n = rows * cols
        p2 = dict(phys)
        s = F.Line2d.System_Cuspy_Laplace(k_interactions=1.0, k_frame=1.0 / n, shape=[rows, cols],
                                          **p2)
        s.u_frame = 1.0
        sec = timed_steps(s, 40, 5)
        out["config5_line2d_4096_verlet"] = {
            "kernel": s.last_kernel, "us_per_step": 1e6 * sec, "block_updates_per_s": n / sec,
            "algorithmic_GBps": 64.0 * n / sec / 1e9, "frac_of_hbm_peak": 64.0 * n / sec / 1e9 / peak}
        del s
        for key in ("m", "eta", "dt"):
            p2.pop(key)
        s = F.Line2d.System_Cuspy_Laplace_Nopassing(k_interactions=1.0, k_frame=1.0 / n,
                                                    shape=[rows, cols], **p2)
        s.u_frame = 1.0
        steps0 = s.step_count
        t0 = time.perf_counter()
        ret = s.minimise(max_iter=2000, max_iter_is_error=False)
        wall = time.perf_counter() - t0
        sweeps = int(s.step_count - steps0)
        # sweeps at the fixed point move no block between wells: the streaming rate of the kernel
        s.minimise(tol=1e-300, max_iter=60, max_iter_is_error=False)
        per = s.last_kernel_seconds / max(1, s.last_kernel_launches)
        out["config5_line2d_4096_nopassing"] = {
            "kernel": s.last_kernel, "minimise_ret": int(ret), "minimise_sweeps": sweeps,
            "minimise_wall_ms": 1e3 * wall, "us_per_sweep_at_fixed_point": 1e6 * per,
            "block_updates_per_s": n / per, "algorithmic_GBps": 32.0 * n / per / 1e9,
            "frac_of_hbm_peak": 32.0 * n / per / 1e9 / peak}
        del s
    except Exception as e:  # pragma: no cover
        out["config5_line2d_4096"] = {"error": str(e)}
    try:  # thermal systems (External = RandomNormalForcing), flowSteps
        for label, N, R, steps in (("thermal_example_N1000", 1000, 1, 2000),
                                   ("thermal_ensemble_1024x4096", 4096, 1024, 200)):
            rng = np.random.default_rng(0)
            s = F.Line1d.Ensemble_Cuspy_Laplace_RandomForcing(
                k_interactions=1.0, k_frame=1.0 / N, shape=[N], nrealisations=R, mean=0.0,
                stddev=0.05, seed_forcing=0, dinc_init=rng.integers(0, 100, N),
                dinc=100 * np.ones(N, dtype=np.int64), **phys)
            s.flowSteps(steps, 5e-2)
            s.flowSteps(steps, 5e-2)
            sec = s.last_kernel_seconds / steps
            out[label] = {"kernel": s.last_kernel, "us_per_step": 1e6 * sec,
                          "block_updates_per_s": N * R / sec}
            del s
    except Exception as e:  # pragma: no cover
        out["thermal"] = {"error": str(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 16384 realisations per GPU; strong: 16384 in total")
    ap.add_argument("--realisations", type=int, default=R_CONFIG,
                    help="per GPU (weak) / in total (strong)")
    ap.add_argument("--inner", type=int, default=1000, help="Verlet steps per timeSteps call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream", action="store_true")
    ap.add_argument("--no-events", action="store_true")
    ap.add_argument("--no-driven", action="store_true")
    ap.add_argument("--no-contracted", action="store_true")
    ap.add_argument("--no-slab", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short device-timed lines of BASELINE configs #3-#5 and the "
                         "thermal systems")
    ap.add_argument("--quick", action="store_true", help="shorter side measurements")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    import frictionqpotspringblock_b200 as F
    from frictionqpotspringblock_b200._capi import lib, check

    if F.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    N, T = N_BLOCKS, args.inner
    K, W = args.steps, args.warmup
    if args.scaling == "strong":
        from frictionqpotspringblock_b200.distributed import shard_realisations

        first, R = shard_realisations(args.realisations, rank, world)
        R_total = args.realisations
    else:
        R = args.realisations
        first = rank * R
        R_total = world * R
    kw = physics(N)
    ens = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, seed=first * N, device=local_rank, **kw)
    stream = torch.cuda.current_stream()
    ens.set_stream(stream.cuda_stream)
    n = R * N

    # ---- prepare: equilibrium, then an event-driven kick (examples/Line1d_Cuspy_Laplace.py:44-52)
    assert np.all(ens.minimise() == 0)
    ens.eventDrivenStep(1e-3, False)
    ens.eventDrivenStep(1e-3, True)

    # host staging for the end-to-end arm (pinned)
    hu = torch.empty(n, dtype=torch.float64, pin_memory=True)
    hv = torch.empty(n, dtype=torch.float64, pin_memory=True)
    ha = torch.empty(n, dtype=torch.float64, pin_memory=True)
    hout = torch.empty(n, dtype=torch.float64, pin_memory=True)
    hmean = np.empty(R, dtype=np.float64)
    for which, buf in ((0, hu), (1, hv), (2, ha)):
        check(lib.fqsb_get(ens._h, which, buf.data_ptr(), n))

    # ---- end-to-end arm: host buffers in, host buffers out, every copy inside the timed region.
    #      ONE handle, ONE call per step: fqsb_run_from_host = `system.u/v/a = ...; timeSteps(T);
    #      read back u and mean(f_frame)` with the copies of one chunk of realisations hidden
    #      behind the kernels of the others on the handle's internal streams.
    def e2e_step():
        check(lib.fqsb_run_from_host(ens._h, hu.data_ptr(), hv.data_ptr(), ha.data_ptr(), n, T,
                                     hout.data_ptr(), None, None, hmean.ctypes.data))

    for _ in range(max(1, W // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    barrier()
    e2e_sec = max_over_ranks(time.perf_counter() - t0)
    e2e_value = R_total * N * T * K / e2e_sec
    h2d = 3 * n * 8
    d2h = n * 8 + R * 8

    # the same through the separate public calls (copies and kernel of a step serialise)
    def separate_step():
        check(lib.fqsb_set_u(ens._h, hu.data_ptr(), n))
        check(lib.fqsb_set_v(ens._h, hv.data_ptr(), n))
        check(lib.fqsb_set_a(ens._h, ha.data_ptr(), n))
        check(lib.fqsb_time_steps(ens._h, T))
        check(lib.fqsb_get(ens._h, 0, hout.data_ptr(), n))
        check(lib.fqsb_mean_f_frame(ens._h, hmean.ctypes.data))

    separate_step()
    barrier()
    t0 = time.perf_counter()
    nsep = 2
    for _ in range(nsep):
        separate_step()
    barrier()
    sep_sec = max_over_ranks(time.perf_counter() - t0)

    # ---- device-resident arm: the state is already in HBM when the timed region starts
    check(lib.fqsb_set_u(ens._h, hu.data_ptr(), n))
    check(lib.fqsb_set_v(ens._h, hv.data_ptr(), n))
    check(lib.fqsb_set_a(ens._h, ha.data_ptr(), n))
    for _ in range(W):
        ens.timeSteps(T)
    ens.mark_indices()
    launches0 = ens.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    kernel_sec = 0.0
    kernel_launches = 0
    barrier()
    ev0.record(stream)
    for _ in range(K):
        ens.timeSteps(T)
        kernel_sec += ens.last_kernel_seconds
        kernel_launches += ens.last_kernel_launches
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    sec = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    gpu_launches = ens.launch_count - launches0
    value = R_total * N * T * K / sec
    kernel_name = ens.last_kernel
    hops_quiescent = float(np.sum(np.abs(ens.avalanche_since_mark()[0]))) / (R * N * T * K)

    # ---- roofline of the dominant kernel (k_resident). Its state never leaves the chip within a
    #      launch, so HBM is not its roof: it is bound by the FP64 pipe's ISSUE rate (64 lanes per
    #      SM per clock; 30.6 FP64-pipe instructions per block-update without FMA). The algorithmic
    #      HBM figure of SURVEY 8(d) is the sub-entry "hbm".
    peak, peak_src = measured_peaks()
    kernel_avg = kernel_sec / max(1, kernel_launches)
    sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
    fp64_peak = 148 * 64 * sm_clock * 1e6
    fp64_rate = FP64_INSTR_PER_UPDATE * R * N * T / kernel_avg
    hbm_ach = ALGO_BYTES_PER_UPDATE * R * N * T / kernel_avg / 1e9
    roofline = {
        "bound": "fp64_issue",
        "kernel": "k_resident<Cuspy,Laplace1d,B=8,T=512,full,unit>" if kernel_name == "resident"
        else kernel_name,
        "achieved": fp64_rate / 1e9, "peak": fp64_peak / 1e9, "unit": "G FP64-pipe instr/s",
        "frac": fp64_rate / fp64_peak,
        "peak_source": "148 SMs x 64 FP64 lanes x SM clock sampled during the timed region",
        "instr_per_block_update": FP64_INSTR_PER_UPDATE,
        "frac_of_measured_issue_peak": fp64_rate / (fp64_peak * FP64_PRACTICAL_ISSUE),
        "measured_issue_peak_source": "tools/fp64_peak.cu: 1.934 of 2.0 warp-instr/SM/clock "
                                      "(DMUL+DADD chains, 512 threads/SM)",
        "launch_ms": 1e3 * kernel_avg, "traffic": None,
        "hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": peak, "unit": "GB/s",
                "frac": hbm_ach / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_block_update": ALGO_BYTES_PER_UPDATE,
                "note": "algorithmic bytes of SURVEY 8(d); > 1 by construction: the state stays "
                        "on chip for all T steps of a launch (real DRAM traffic in `traffic`); "
                        "the kernel that streams HBM once per step is in roofline_stream"},
    }
    traffic_file = ROOT / "profiles" / "traffic.json"
    tr = {}
    if traffic_file.exists():
        try:
            tr = json.loads(traffic_file.read_text())
        except Exception:
            tr = {}
    if tr.get("k_resident_bytes_per_block_per_launch"):
        # state is read and written once per launch, whatever the step count
        roofline["traffic"] = tr["k_resident_bytes_per_block_per_launch"] * R * N
        roofline["traffic_source"] = tr.get("k_resident_source")

    # ---- driven variant: the frame moves at a finite velocity (flowSteps), blocks change wells
    #      all the time (hop_shared, the pcg32 refill and the index bookkeeping run every step)
    driven = None
    if not args.no_driven:
        v_frame = 1.0
        ens.u_frame = ens.u_frame + 0.35 / kw["k_frame"]  # well past the depinning force
        for _ in range(2):
            ens.flowSteps(T, v_frame)
        ens.mark_indices()
        Kd = 3
        barrier()
        ev0.record(stream)
        for _ in range(Kd):
            ens.flowSteps(T, v_frame)
        ev1.record(stream)
        barrier()
        dsec = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
        S_d, A_d = ens.avalanche_since_mark()
        driven = {"value": R_total * N * T * Kd / dsec, "unit": UNIT, "v_frame": v_frame,
                  "ms_per_step": 1e3 * dsec / Kd,
                  "hops_per_block_update": float(np.sum(S_d)) / (R * N * T * Kd),
                  "blocks_that_moved": float(np.mean(A_d)) / N,
                  "how": f"flowSteps({T}, {v_frame}) x {Kd} in sliding motion, device-timed"}

    # ---- opt-in contracted arithmetic (contracted=True = FQSB_KERNEL_FMA): the same kernel built
    #      with FMA contraction. NOT part of `value` (the default stays bit-identical to the
    #      reference arithmetic); parity of this mode: tests/test_gpu_contracted.py (landscape
    #      exact, trajectories to rounding, goldens S-exact).
    contracted = None
    del ens
    if not args.no_contracted:
        ensc = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, seed=first * N, device=local_rank,
                                               kernel=1, contracted=True, **kw)
        ensc.set_stream(stream.cuda_stream)
        assert np.all(ensc.minimise() == 0)
        ensc.eventDrivenStep(1e-3, False)
        ensc.eventDrivenStep(1e-3, True)
        for _ in range(2):
            ensc.timeSteps(T)
        csec, cl = 0.0, 0
        for _ in range(3):
            ensc.timeSteps(T)
            csec += ensc.last_kernel_seconds
            cl += ensc.last_kernel_launches
        csec = max_over_ranks(csec)
        ensc.eventDrivenStep(1e-3, False)
        ensc.eventDrivenStep(1e-3, True)
        steps0 = ensc.step_count  # (steps over all realisations of the handle)
        assert np.all(ensc.minimise() == 0)
        msec = max_over_ranks(ensc.last_kernel_seconds)
        msteps = float(ensc.step_count - steps0)
        contracted = {"value": R_total * N * T * 3 / csec, "unit": UNIT,
                      "ms_per_step": 1e3 * csec / 3,
                      "minimise_block_updates_per_s": world * msteps * N / msec,
                      "how": "the same ensemble, protocol and timeSteps(T) calls with "
                             "contracted=True (resident kernel compiled with -fmad=true: 20.6 "
                             "instead of 30.6 FP64-pipe instructions per block-update); "
                             "device-timed; results equal to rounding, not bit for bit"}
        del ensc

    # ---- the streaming kernel K1 on the same ensemble (one fused step per launch, HBM-bound)
    roofline_stream = None
    if not args.no_stream:
        ens2 = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, seed=first * N,
                                               device=local_rank, kernel=2, **kw)
        ens2.set_stream(stream.cuda_stream)
        ens2.u_frame = np.full(R, 1.0)
        Ts = 20
        for _ in range(3):
            ens2.timeSteps(Ts)
        ksec, kl = 0.0, 0
        for _ in range(3):
            ens2.timeSteps(Ts)
            ksec += ens2.last_kernel_seconds
            kl += ens2.last_kernel_launches
        per = ALGO_BYTES_PER_UPDATE * R * N
        ach = per / (ksec / kl) / 1e9
        roofline_stream = {
            "bound": "hbm", "kernel": "k_stream_1d<Cuspy,Laplace1d,unit>", "achieved": ach,
            "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
            "launch_ms": 1e3 * ksec / kl, "peak_source": peak_src,
            "block_updates_per_s": R_total * N / (ksec / kl),
        }
        if tr.get("k_stream_1d_bytes_per_block_update"):
            roofline_stream["traffic"] = tr["k_stream_1d_bytes_per_block_update"] * R * N
            roofline_stream["traffic_source"] = tr.get("k_stream_1d_source")
        del ens2

    # ---- quasistatic events/s (BASELINE metric, second part): one event = eventDrivenStep to
    #      the next instability + kick + minimise, on every realisation of a fresh ensemble; the
    #      avalanche sizes come from the device-side mark (R x 16 bytes per event over PCIe)
    events = None
    if not args.no_events:
        ens3 = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, seed=first * N,
                                               device=local_rank, **kw)
        ens3.set_stream(stream.cuda_stream)
        ens3.minimise()
        nev = 3
        S_sum = 0
        barrier()
        t0 = time.perf_counter()
        steps0 = ens3.step_count
        for _ in range(nev):
            ens3.mark_indices()
            ens3.eventDrivenStep(1e-3, False)
            ens3.eventDrivenStep(1e-3, True)
            ret = ens3.minimise()
            assert np.all(ret == 0)
            S_ev, A_ev = ens3.avalanche_since_mark()
            S_sum += int(np.sum(S_ev))
        barrier()
        ev_sec = max_over_ranks(time.perf_counter() - t0)
        # (outside the timed region) the device-side S of the last event against host arithmetic
        i_now = ens3.chunk.index_at_align
        ens3.mark_indices()
        ens3.eventDrivenStep(1e-3, False)
        ens3.eventDrivenStep(1e-3, True)
        ens3.minimise()
        S_ev, A_ev = ens3.avalanche_since_mark()
        i_new = ens3.chunk.index_at_align
        assert np.array_equal(S_ev, np.sum(i_new - i_now, axis=1))
        assert np.array_equal(A_ev, np.sum(i_new != i_now, axis=1))
        events = {"value": R_total * nev / ev_sec, "unit": "events/s",
                  "realisations_per_gpu": R, "events_per_realisation": nev,
                  "mean_minimise_steps": (ens3.step_count - steps0) / (R * nev),
                  "mean_S": S_sum / (R * nev),
                  "block_updates_per_s": sum_over_ranks((ens3.step_count - steps0) * N) / ev_sec,
                  "S_A_checked_against_host_arithmetic": True}
        del ens3, i_now, i_new

    # ---- strong scaling of the named config (16384 realisations in TOTAL over the ranks)
    strong = None
    if args.scaling == "weak":
        from frictionqpotspringblock_b200.distributed import shard_realisations

        f4, R4 = shard_realisations(R_CONFIG, rank, world)
        ens4 = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R4, seed=f4 * N, device=local_rank,
                                               **kw)
        ens4.set_stream(stream.cuda_stream)
        assert np.all(ens4.minimise() == 0)
        ens4.eventDrivenStep(1e-3, False)
        ens4.eventDrivenStep(1e-3, True)
        ens4.timeSteps(T)
        Ks = 3
        barrier()
        ev0.record(stream)
        for _ in range(Ks):
            ens4.timeSteps(T)
        ev1.record(stream)
        barrier()
        ssec = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
        strong = {"value": R_CONFIG * N * T * Ks / ssec, "unit": UNIT,
                  "realisations_total": R_CONFIG, "realisations_per_gpu": R4,
                  "ms_per_step": 1e3 * ssec / Ks}
        del ens4

    # ---- ONE large system decomposed over all ranks (configs #3, #5), parity asserted
    slab = None
    if not args.no_slab:
        slab = slab_block(F, rank, world, local_rank, barrier, max_over_ranks, args.quick)

    # ---- the other BASELINE configs and the thermal systems (rank 0, device-timed kernel
    #      launches, short): reported next to the headline, not part of `value`
    other = None
    if rank == 0 and not args.no_other_configs:
        other = other_configs(F, local_rank, peak)

    # ---- CPU baseline (rank 0): the oracle port on all host cores, bounded sample, same protocol
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        nsys = 2 * cores
        cpu = {}
        for flavour in ("faithful", "fused"):
            cens = cpu_ensemble(nsys, cores, flavour == "fused")
            sec_probe, _ = cens.time_steps(T)  # warm-up step, sizes the sample (~6 s per flavour)
            reps = int(min(40, max(2, 6.0 / max(sec_probe, 1e-3))))
            csec = 0.0
            for _ in range(reps):
                sec_rep, _ = cens.time_steps(T)
                csec += sec_rep
            cpu[flavour] = {"value": nsys * N * T * reps / csec, "unit": UNIT,
                            "sample": cpu_sample_text(nsys, cores, T, flavour) +
                            f", {reps} steps, {csec:.1f} s timed"}
            del cens
        cpu = {"value": cpu["faithful"]["value"], "unit": UNIT, "cores": cores, "kind": "port",
               "sample": cpu["faithful"]["sample"], "fused": cpu["fused"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": 1e3 * sec / K, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(T, args.scaling),
            "realisations_per_gpu": R, "hops_per_block_update": hops_quiescent,
            "roofline": roofline, "roofline_stream": roofline_stream, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_sec / K,
                    "handles": 1,
                    "how": "ONE handle, one fqsb_run_from_host call per step: u, v, a from pinned "
                           "host memory, timeSteps(T), u and mean f_frame per realisation back to "
                           "the host; the handle pipelines chunks of realisations over its "
                           "internal streams so copies overlap kernels",
                    "separate_calls": {"value": R_total * N * T * nsep / sep_sec,
                                       "ms_per_step": 1e3 * sep_sec / nsep,
                                       "how": "set_u, set_v, set_a, time_steps, get, mean_f_frame "
                                              "as six calls (copies and kernel serialise)"}},
            "gpu_launches": int(gpu_launches), "clocks": clocks,
            "driven": driven, "contracted_arithmetic": contracted, "strong_scaling": strong,
            "quasistatic_events": events, "slab": slab,
            "other_configs": other,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
