#!/usr/bin/env python
"""bench.py -- the headline measurement of the B200-native FrictionQPotSpringBlock integrator.

Workload (BASELINE.json configs[1]): Line1d System_Cuspy_Laplace ensemble, 16384 disorder
realisations x N = 4096 blocks per GPU (weak scaling: every rank integrates its own 16384
realisations, disjoint seeds, no inter-GPU traffic), physics of examples/Line1d_Cuspy_Laplace.py.
One "step" = one ``timeSteps(T)`` call over the whole ensemble (T velocity-Verlet steps), started
from a kicked (avalanching) state. Metric: block-updates/s = realisations x N x T / time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every field).
"""

from __future__ import annotations

import argparse
import ctypes as C
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
ALGO_BYTES_PER_UPDATE = 64.0  # u,v,a read + write (48) + y_left,y_right read (16); SURVEY 8(d)
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


def cpu_arm(nsys, nthreads, T, seconds_target=None):
    """The restated reference (oracle port) on the host cores: nsys lines of N_BLOCKS blocks."""
    from oracle import oracle as orc

    kw = physics(N_BLOCKS)
    par = orc.make_params("Cuspy", "Laplace1d", 0, kw["shape"], kw["m"], kw["eta"], kw["mu"], 0.0,
                          kw["k_interactions"], 0.0, kw["k_frame"], kw["dt"], 0, "random",
                          kw["parameters"], kw["offset"], 5000)
    return orc.CpuEnsemble(par, nsys, nthreads)


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on all host cores.
    The genuine library cannot be built (xtensor/prrng/GooseFEM absent), so this is the oracle
    port (kind "port"), pinned on the reference's goldens. Rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nsys = 2 * cores
    T = args.ref_inner
    ens = cpu_arm(nsys, cores, T)
    for _ in range(args.warmup):
        ens.time_steps(T)
    total = 0.0
    for k in range(args.steps):
        if k and k % 4 == 0:
            ens.kick()  # keep the lines active (untimed), as the GPU arm's kicked state
        sec, _cs = ens.time_steps(T)
        total += sec
    updates = nsys * N_BLOCKS * T * args.steps
    value = updates / total
    sample = (f"{nsys} realisations x N={N_BLOCKS} x timeSteps({T}) per step on {cores} threads "
              f"(one realisation per thread), after minimise + one kick")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, nsys, T, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, R, T, where):
    return {
        "workload": "Line1d System_Cuspy_Laplace ensemble (BASELINE configs[1]): "
                    f"{R} realisations x N={N_BLOCKS} per {'GPU' if where == 'gpu' else 'host'}, "
                    f"one step = timeSteps({T}) after minimise + eventDrivenStep kick",
        "realisations_per_gpu": R, "blocks": N_BLOCKS, "inner_steps": T,
        "physics": "m=1 eta=2sqrt(3)/10 mu=1 k=1 k_frame=1/N dt=0.1 random[2.0] offset=-50",
        "parallelism": "independent realisations sharded across GPUs, no collective",
        "l2": "state per GPU (3.8 GB) is larger than L2, no flush needed",
    }


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
    ap.add_argument("--realisations", type=int, default=16384, help="per GPU")
    ap.add_argument("--inner", type=int, default=1000, help="Verlet steps per timeSteps call")
    ap.add_argument("--ref-inner", type=int, default=500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream", action="store_true")
    ap.add_argument("--no-events", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short device-timed lines of BASELINE configs #3-#5 and the "
                         "thermal systems")
    ap.add_argument("--pipeline", type=int, default=8, help="handles of the pipelined e2e arm")
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

    R, N, T = args.realisations, N_BLOCKS, args.inner
    K, W = args.steps, args.warmup
    kw = physics(N)
    ens = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=R, seed=rank * R * N, device=local_rank,
                                          **kw)
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

    def e2e_step():
        # the calls a user of the reference makes: system.u/v/a = ...; timeSteps; read back
        check(lib.fqsb_set_u(ens._h, hu.data_ptr(), n))
        check(lib.fqsb_set_v(ens._h, hv.data_ptr(), n))
        check(lib.fqsb_set_a(ens._h, ha.data_ptr(), n))
        check(lib.fqsb_time_steps(ens._h, T))
        check(lib.fqsb_get(ens._h, 0, hout.data_ptr(), n))
        check(lib.fqsb_mean_f_frame(ens._h, hmean.ctypes.data))

    # ---- end-to-end arm (host buffers, copies inside the timed region), one handle: the three
    #      uploads, the kernel and the download of a step serialise
    for _ in range(max(1, W // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    barrier()
    e2e1_sec = max_over_ranks(time.perf_counter() - t0)
    e2e1_value = world * R * N * T * K / e2e1_sec
    h2d = 3 * n * 8
    d2h = n * 8 + R * 8

    # ---- end-to-end arm, pipelined: the same ensemble split over `npipe` handles driven by
    #      host threads through the same public calls (ctypes releases the GIL, every handle owns
    #      a stream), so one handle's PCIe copies overlap another handle's kernel. Same bytes,
    #      same steps, all copies inside the timed region.
    npipe = args.pipeline if R % max(1, args.pipeline) == 0 else 1
    e2e_sec, e2e_value = e2e1_sec, e2e1_value
    if npipe > 1:
        from concurrent.futures import ThreadPoolExecutor

        Rp = R // npipe
        npart = Rp * N
        parts = []
        for k in range(npipe):
            ek = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=Rp, device=local_rank,
                                                 seed=rank * R * N + k * Rp * N, **kw)
            assert np.all(ek.minimise() == 0)
            ek.eventDrivenStep(1e-3, False)
            ek.eventDrivenStep(1e-3, True)
            sl = slice(k * npart, (k + 1) * npart)
            for which, buf in ((0, hu), (1, hv), (2, ha)):
                check(lib.fqsb_get(ek._h, which, buf[sl].data_ptr(), npart))
            parts.append((ek, sl, np.empty(Rp, dtype=np.float64)))

        def part_step(part):
            ek, sl, mean = part
            torch.cuda.set_device(local_rank)
            check(lib.fqsb_set_u(ek._h, hu[sl].data_ptr(), npart))
            check(lib.fqsb_set_v(ek._h, hv[sl].data_ptr(), npart))
            check(lib.fqsb_set_a(ek._h, ha[sl].data_ptr(), npart))
            check(lib.fqsb_time_steps(ek._h, T))
            check(lib.fqsb_get(ek._h, 0, hout[sl].data_ptr(), npart))
            check(lib.fqsb_mean_f_frame(ek._h, mean.ctypes.data))

        def part_loop(arg):
            part, nsteps, delay = arg
            time.sleep(delay)  # stagger the handles so that copies and kernels interleave
            for _ in range(nsteps):
                part_step(part)

        with ThreadPoolExecutor(max_workers=npipe) as pool:
            list(pool.map(part_loop, [(p, max(1, W // 2), 0.0) for p in parts]))
            barrier()
            # offset of about one part's upload time (copies are ~1/4 of a single-handle step)
            stagger = 0.25 * e2e1_sec / K / npipe
            t0 = time.perf_counter()
            list(pool.map(part_loop, [(p, K, k * stagger) for k, p in enumerate(parts)]))
            barrier()
            e2e_sec = max_over_ranks(time.perf_counter() - t0)
        e2e_value = world * R * N * T * K / e2e_sec
        for ek, _sl, _m in parts:
            del ek
        del parts
        # restore the single-handle snapshot for the device-resident arm
        for which, buf in ((0, hu), (1, hv), (2, ha)):
            check(lib.fqsb_get(ens._h, which, buf.data_ptr(), n))

    # ---- device-resident arm: the state is already in HBM when the timed region starts
    check(lib.fqsb_set_u(ens._h, hu.data_ptr(), n))
    check(lib.fqsb_set_v(ens._h, hv.data_ptr(), n))
    check(lib.fqsb_set_a(ens._h, ha.data_ptr(), n))
    for _ in range(W):
        ens.timeSteps(T)
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
    value = world * R * N * T * K / sec
    kernel_name = ens.last_kernel

    # ---- roofline of the dominant kernel (k_resident): algorithmic bytes / kernel duration
    peak, peak_src = measured_peaks()
    per_launch_bytes = ALGO_BYTES_PER_UPDATE * R * N * T
    kernel_avg = kernel_sec / max(1, kernel_launches)
    achieved = per_launch_bytes / kernel_avg / 1e9
    roofline = {
        "bound": "hbm",
        "kernel": "k_resident<Cuspy,Laplace1d,B=8,T=512,full,unit>" if kernel_name == "resident"
        else kernel_name,
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": None, "peak_source": peak_src,
        "algorithmic_bytes_per_block_update": ALGO_BYTES_PER_UPDATE,
        "launch_ms": 1e3 * kernel_avg,
        "note": "resident kernel: state stays on chip for all T steps of a launch, so DRAM "
                "traffic is ~64/T B per block-update and the algorithmic-byte fraction exceeds 1; "
                "its real bound is the FP64 pipe (fp64_pipe); the HBM-streaming kernel (one step "
                "per pass over HBM) is reported in roofline_stream",
    }
    # the resident kernel's own bound: 32 FP64 pipe instructions per block-update (30 DADD/DMUL
    # in the reference's evaluation order without FMA + 2 DSETP), 64 lanes per SM per clock
    sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
    fp64_peak = 148 * 64 * sm_clock * 1e6
    roofline["fp64_pipe"] = {
        "ops_per_block_update": 32, "achieved_ops_per_s": 32 * R * N * T / kernel_avg,
        "peak_ops_per_s": fp64_peak, "frac": 32 * R * N * T / kernel_avg / fp64_peak,
        "peak_source": "148 SMs x 64 FP64 lanes x measured SM clock",
    }
    traffic_file = ROOT / "profiles" / "traffic.json"
    if traffic_file.exists():
        try:
            tr = json.loads(traffic_file.read_text())
            per_block = tr.get("k_resident_bytes_per_block_per_launch")
            if per_block:
                # state is read and written once per launch, whatever the step count
                roofline["traffic"] = per_block * R * N
                roofline["traffic_source"] = tr.get("k_resident_source")
        except Exception:
            pass

    # ---- the streaming kernel K1 on the same ensemble (one fused step per launch, HBM-bound)
    roofline_stream = None
    if not args.no_stream:
        del ens
        Rs = R
        ens2 = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=Rs, seed=rank * R * N,
                                               device=local_rank, kernel=2, **kw)
        ens2.set_stream(stream.cuda_stream)
        ens2.u_frame = np.full(Rs, 1.0)
        Ts = 20
        for _ in range(3):
            ens2.timeSteps(Ts)
        ksec, kl = 0.0, 0
        for _ in range(3):
            ens2.timeSteps(Ts)
            ksec += ens2.last_kernel_seconds
            kl += ens2.last_kernel_launches
        per = ALGO_BYTES_PER_UPDATE * Rs * N
        ach = per / (ksec / kl) / 1e9
        roofline_stream = {
            "bound": "hbm", "kernel": "k_stream_1d<Cuspy,Laplace1d,unit>", "achieved": ach,
            "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
            "launch_ms": 1e3 * ksec / kl,
            "block_updates_per_s": world * Rs * N / (ksec / kl),
        }
        if traffic_file.exists():
            try:
                tr = json.loads(traffic_file.read_text())
                per_update = tr.get("k_stream_1d_bytes_per_block_update")
                if per_update:
                    roofline_stream["traffic"] = per_update * Rs * N
                    roofline_stream["traffic_source"] = tr.get("k_stream_1d_source")
            except Exception:
                pass
        del ens2

    # ---- quasistatic events/s (BASELINE metric, second part): one event = eventDrivenStep to
    #      the next instability + kick + minimise, on every realisation of a fresh ensemble
    events = None
    if not args.no_events:
        Re = R  # the whole ensemble of the named config (a longer work queue hides the tail of
        #         the realisations with the longest avalanches)
        ens3 = F.Line1d.Ensemble_Cuspy_Laplace(nrealisations=Re, seed=rank * R * N,
                                               device=local_rank, **kw)
        ens3.set_stream(stream.cuda_stream)
        ens3.minimise()
        nev = 3
        barrier()
        t0 = time.perf_counter()
        steps0 = ens3.step_count
        for _ in range(nev):
            ens3.eventDrivenStep(1e-3, False)
            ens3.eventDrivenStep(1e-3, True)
            ret = ens3.minimise()
            assert np.all(ret == 0)
        barrier()
        ev_sec = max_over_ranks(time.perf_counter() - t0)
        events = {"value": world * Re * nev / ev_sec, "unit": "events/s",
                  "realisations_per_gpu": Re, "events_per_realisation": nev,
                  "mean_minimise_steps": (ens3.step_count - steps0) / (Re * nev),
                  "block_updates_per_s": world * (ens3.step_count - steps0) * N / ev_sec}
        del ens3

    # ---- the other BASELINE configs and the thermal systems (rank 0, device-timed kernel
    #      launches, short): reported next to the headline, not part of `value`
    other = None
    if rank == 0 and not args.no_other_configs:
        other = other_configs(F, local_rank, peak)

    # ---- CPU baseline (rank 0): the oracle port on all host cores, bounded sample
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        nsys = 2 * cores
        cens = cpu_arm(nsys, cores, args.ref_inner)
        cens.time_steps(50)
        sec_probe, _ = cens.time_steps(100)
        rate = nsys * N * 100 / sec_probe
        # ~12 s of CPU work in kick -> timeSteps(1000) cycles (an un-kicked line decays into
        # denormal velocities after ~1e4 steps, which x86 executes ~5x slower: not representative)
        Tc = 1000
        reps = int(min(200, max(2, 12.0 * rate / (nsys * N * Tc))))
        csec = 0.0
        for _ in range(reps):
            cens.kick()
            sec_rep, _ = cens.time_steps(Tc)
            csec += sec_rep
        cpu = {"value": nsys * N * Tc * reps / csec, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{reps} x [kick + timeSteps({Tc})] on {nsys} realisations x N={N}, "
                         f"{cores} threads, {csec:.1f} s timed, oracle/fqsb_oracle.c -O3 "
                         f"-march=native"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": 1e3 * sec / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, R, T, "gpu"),
            "roofline": roofline, "roofline_stream": roofline_stream, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_sec / K,
                    "handles": npipe,
                    "how": "set_u/set_v/set_a from pinned host memory, timeSteps(T), get u + "
                           "mean f_frame per realisation, through the C ABI; the ensemble is "
                           "split over `handles` handles driven by host threads so copies "
                           "overlap kernels",
                    "single_handle": {"value": e2e1_value, "ms_per_step": 1e3 * e2e1_sec / K}},
            "gpu_launches": int(gpu_launches), "clocks": clocks,
            "quasistatic_events": events,
            "other_configs": other,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
