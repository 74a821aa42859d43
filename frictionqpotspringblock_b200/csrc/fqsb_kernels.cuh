// fqsb_kernels.cuh -- sm_100a kernels of the integrator hot path.
//
//  k_resident<POT,INT,B,T>      K2: one CTA = one realisation, state on chip for the whole call
//                               (timeSteps / flowSteps / timeStepsUntilEvent / minimise /
//                               minimise_truncate), stop tests without a host round trip.
//  k_resident_nopassing<..>     K6: overdamped no-passing Jacobi sweeps, same residency.
//  k_stream_step<POT,INT>       K1: one fused Verlet step per launch over all blocks, streaming
//                               u,v,a,y_l,y_r once (64 B per block-update), last-CTA finalise.
//  k_stream_np                 K6 streaming (sweep fused with the previous residual).
//  k_init / k_align / k_forces / k_reduce_* / k_chunk_data / ...   K4, K5, K8 helpers.
#pragma once

#include "fqsb_device.cuh"

namespace fqsb {

// per-call progress, replicated in the registers of every thread that takes decisions
struct Prog {
    i64 steps, S, A, s_n, inc, qs_first, qs_last;
    int init;
};

__device__ __forceinline__ void prog_load(Prog& g, const Ctl& c)
{
    g.steps = c.steps;
    g.S = c.S;
    g.A = c.A;
    g.s_n = c.s_n;
    g.inc = c.inc;
    g.qs_first = c.qs_first;
    g.qs_last = c.qs_last;
    g.init = c.init;
}

__device__ __forceinline__ void prog_store(const Prog& g, Ctl& c)
{
    c.steps = g.steps;
    c.S = g.S;
    c.A = g.A;
    c.s_n = g.s_n;
    c.inc = g.inc;
    c.qs_first = g.qs_first;
    c.qs_last = g.qs_last;
    c.init = g.init;
}

// The per-step decisions of timeStepsUntilEvent (detail.h:1605-1619), minimise (1764-1784)
// and minimise_truncate (1858-1886), evaluated by a full warp (ring entry l lives in lane l).
__device__ __forceinline__ int step_decide(const RunArgs& A, Prog& g, RingEntry& ring, int lane,
                                           double sf, double sff, int hops, int dS, int dA)
{
    g.steps++;
    if (A.mode == MODE_FIXED) {
        return g.steps >= A.max_steps ? ST_EXHAUSTED : ST_RUNNING;
    }
    if (sf != sf) { // NaN forces <=> NaN positions (detail.h:1567)
        return ST_NAN;
    }
    if (A.mode == MODE_UNTIL_EVENT && hops > 0) {
        return ST_EVENT;
    }
    ring = ring_roll_insert(ring, ring_entry(sf, sff), A.niter_tol, lane);
    if (A.track) {
        g.S += dS;
        g.A += dA;
        if (g.S != g.s_n) {
            if (g.init) {
                g.init = 0;
                g.qs_first = g.inc;
            }
            g.qs_last = g.inc;
        }
        g.s_n = g.S;
    }
    if (ring_stop(ring, A.niter_tol, lane, A.tol2, A.tol2 * A.tol2)) {
        return ST_CONVERGED;
    }
    if (A.mode == MODE_TRUNCATE) {
        if (A.A_truncate > 0 && g.A >= A.A_truncate) {
            return ST_TRUNCATED;
        }
        if (A.S_truncate > 0 && g.S >= A.S_truncate) {
            return ST_TRUNCATED;
        }
    }
    return g.steps >= A.max_steps ? ST_EXHAUSTED : ST_RUNNING;
}

__device__ __forceinline__ RingEntry ring_load(const Ctl& c, const RunArgs& A, int lane)
{
    RingEntry e;
    const bool in = lane < A.niter_tol && lane < FQSB_RING;
    e.num = in ? c.ring[lane] : 0.0;
    e.den = in ? c.ring_den[lane] : 1.0;
    return e;
}

__device__ __forceinline__ void ring_store(Ctl& c, const RunArgs& A, int lane, RingEntry e)
{
    if (lane < A.niter_tol && lane < FQSB_RING) {
        c.ring[lane] = e.num;
        c.ring_den[lane] = e.den;
    }
}

// MODE_LOG: append the sums of the step that just finished; no decision on the device
__device__ __forceinline__ int step_log(const RunArgs& A, int r, Prog& g, double sf, double sff,
                                        double hops, double dS, double dA)
{
    double* e = A.log + ((size_t)r * A.max_steps + g.steps) * FQSB_NLOG;
    e[0] = sf;
    e[1] = sff;
    e[2] = hops;
    e[3] = dS;
    e[4] = dA;
    g.steps++;
    return g.steps >= A.max_steps ? ST_EXHAUSTED : ST_RUNNING;
}

// change of |i - i_n| and (i != i_n) when a block moves by `moved` wells
__device__ __forceinline__ void track_hop(const RunArgs& A, i64 gp, i64 i_before, int moved,
                                          int& dS, int& dA)
{
    if (A.track) {
        i64 in = A.i_n[gp];
        i64 b = i_before - in, a = b + moved;
        dS += (int)((a < 0 ? -a : a) - (b < 0 ? -b : b));
        dA += (int)(a != 0) - (int)(b != 0);
    }
}

// =============================================================================================
// K2: resident velocity-Verlet. grid = R CTAs (the hardware block scheduler is the work queue
// over realisations), T threads, thread t owns blocks p = t + j*T (j < B): global loads/stores
// are coalesced and shared-memory neighbour reads are conflict-free for every stencil.
//
// On-chip layout for the whole call: slips u in shared memory only (two buffers, 1-D lines
// ghost-padded so the periodic neighbours are plain offsets), v and a in registers, the wells
// (y_l, y_r) in registers or -- YSMEM -- in shared memory, pcg32 states in shared memory.
// DRAM is touched at entry/exit and on the rare well change (idx, i_n).
// =============================================================================================

// a block left its well (out of line: keeps the hot loop's register budget small). The global
// well index stays in DRAM: only the number of wells moved since the launch began is kept on chip
// (`sd`), so a forward move -- the common one, and in driven flow every block makes one every few
// steps -- touches no global memory at all; the index itself is read for the left-boundary check
// of a backward move and for the S / A bookkeeping of the tracking modes.
static __device__ __noinline__ int hop_shared(const Par& P, double un, double* yl, double* yr, u64* st,
                                       const i64* gidx, int* sd, int* underflow, int track,
                                       i64* i_before)
{
    double l = *yl, r = *yr;
    u64 s = *st;
    const int d0 = *sd;
    int moved = well_align_lazy(P, un, l, r, s, [gidx, d0]() { return *gidx + d0; }, underflow);
    *yl = l;
    *yr = r;
    *st = s;
    *sd = d0 + moved;
    if (track) {
        *i_before = *gidx + d0;
    }
    return moved;
}

// the same with generator state and well index updated in place in global memory (the thermal
// resident kernel, whose shared memory is taken by its forcing arrays)
static __device__ __noinline__ int hop_global(const Par& P, double un, double* yl, double* yr, u64* st,
                                       i64* gidx, int* underflow)
{
    double l = *yl, r = *yr;
    u64 s = *st;
    const i64 i0 = *gidx;
    int moved = well_align(P, un, l, r, s, i0, underflow);
    *yl = l;
    *yr = r;
    *st = s;
    *gidx = i0 + moved;
    return moved;
}

#ifndef FQSB_YMID
#define FQSB_YMID 1
#endif
#ifndef FQSB_YMID_FIXED
#define FQSB_YMID_FIXED 0 // ... in the fixed-step loops too: measured slower (see YMID below)
#endif
#ifndef FQSB_UREG
#define FQSB_UREG 1
#endif

// STOP = false: timeSteps / flowSteps only (MODE_FIXED); STOP = true: the stop modes. Two
// instantiations so that the bookkeeping of the stop modes costs the fixed-step loop no registers.
// HOPINL = true: the common well change (one well to the right on a `random` landscape) is taken
// inline. It pays whenever blocks change wells regularly (flowSteps, avalanches inside the stop
// modes: +60 % at 0.07 well changes per block-update) but costs the quiescent fixed-step loop
// ~5 %, so timeSteps() launches the variant without it.
// FMA only names the instantiations of the translation units that are compiled with FMA
// contraction (-fmad=true, FQSB_FMA_BUILD): same source, contracted by the compiler.
template <int POT, int INT, int B, int T, bool YSMEM, bool FULL, bool UNIT, bool STOP,
          bool HOPINL = STOP, bool FMA = false>
__global__ void __launch_bounds__(T)
    k_resident(const __grid_constant__ Par P, const __grid_constant__ State S,
               const __grid_constant__ RunArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool ONE_D = INT < INT_LAPLACE2D;
    constexpr int G = ONE_D ? 1 : 0; // ghost cells on each side of a 1-D line
    constexpr int NW = T / 32;
    // Ownership. Nearest-neighbour lines (NN1D): thread t owns the B CONSECUTIVE blocks
    // p = t*B + j, so B - 1 of the 2B stencil neighbours are the thread's own new slips
    // (registers) and only the two edge neighbours come from shared memory; the slips sit in
    // shared memory transposed (slot j*T + t: conflict-free). Everything else (2-D stencils,
    // LongRange): p = t + j*T, slot p (+ ghost), neighbours all from shared memory.
    constexpr bool NN1D = ONE_D && INT != INT_LONGRANGE1D;
    const int N = (int)P.N;
    const int NS = NN1D ? B * T : N + 2 * G; // slots of one slip buffer
    const int r = blockIdx.x;
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    auto POF = [&](int j) { return NN1D ? t * B + j : t + j * T; };            // block of (t, j)
    auto SLOT = [&](int q) { return NN1D ? (q % B) * T + q / B : q + G; };     // its slip slot

    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }

    double* us = reinterpret_cast<double*>(smem_raw);          // [2][NS]
    // stop modes: per-thread sums of the steps whose residual has not been reduced yet
    // ([FQSB_SKIP_K - 1][T] pairs; 16-byte aligned: 2 NS is even)
    double2* sback = reinterpret_cast<double2*>(us + 2 * (size_t)NS);
    u64* sst = reinterpret_cast<u64*>(sback + (STOP ? (FQSB_SKIP_K - 1) * T : 0)); // [N]
    double* syl = reinterpret_cast<double*>(sst + N);          // [N] if YSMEM
    double* syr = syl + (YSMEM ? N : 0);                       // [N] if YSMEM
    double* spref = syr + (YSMEM ? N : 0);                     // [N] if LongRange
    double* red = spref + (INT == INT_LONGRANGE1D ? N : 0);    // [2][NW][2]
    double* rdb = red + 4 * NW;                                // [K-1][2 NW] warp sums of parked steps
    int* redi = reinterpret_cast<int*>(rdb + (STOP ? (FQSB_SKIP_K - 1) * 2 * NW : 0)); // [2][NW][4]
    int* sdidx = redi + 8 * NW;                                // [N] wells moved in this launch

    const i64 base = (i64)r * P.N;
    double v[B], a[B];
    double yl[YSMEM ? 1 : B], yr[YSMEM ? 1 : B];
    // Cuspy potential: the midpoint of the well, 0.5 * (y_l + y_r), only changes with the well;
    // kept beside it, the force costs one subtraction instead of three operations (same bits).
    // In the fixed-step loop of this kernel it costs more in spills than it saves: 229 instead of
    // 245 FP64 instructions per warp-step but 10 instead of 4 local loads -- 5.57e11 -> 5.48e11
    // block-updates/s (flowSteps -10 %, contracted build -3 %), measured twice (round 2: -0.7 %
    // before the loop was trimmed). Hence STOP only (FQSB_YMID_FIXED = 0). k_blocked, with 4
    // blocks per thread and registers to spare, keeps the midpoint in every variant.
    constexpr bool YMID = FQSB_YMID && POT == POT_CUSPY && !YSMEM && (STOP || FQSB_YMID_FIXED);
    double ym[YMID ? B : 1];
    int ij[ONE_D ? 1 : B];

#pragma unroll
    for (int j = 0; j < B; ++j) {
        const int p = POF(j);
        const int pc = (FULL || p < N) ? p : N - 1;
        v[j] = S.v[base + pc];
        a[j] = S.a[base + pc];
        if (!YSMEM) {
            yl[j] = S.yl[base + pc];
            yr[j] = S.yr[base + pc];
            if (YMID) {
                ym[j] = 0.5 * (yl[j] + yr[j]);
            }
        }
        if (!ONE_D) {
            int i = pc / P.cols;
            ij[j] = (i << 16) | (pc - i * P.cols);
        }
        if (FULL || p < N) {
            us[SLOT(p)] = S.u[base + p];
            if (YSMEM) {
                syl[p] = S.yl[base + p];
                syr[p] = S.yr[base + p];
            }
            sst[p] = S.rng[base + p];
            sdidx[p] = 0;
            if (INT == INT_LONGRANGE1D) {
                spref[p] = S.pref[p];
            }
        }
    }
    __syncthreads(); // clamped threads read block N-1's slip, written by its owner

    double uf = S.u_frame[r];
    // (0.5*dt)*dt, detail.h:1549. The fixed-step loops take it as a kernel parameter (computed on
    // the host: the same two IEEE products), which costs them neither a register nor the two
    // multiplications the compiler otherwise repeats every step; the stop variants keep the
    // expression (their register allocation came out ~10 % slower with the parameter)
    const double c2 = STOP ? 0.5 * P.dt * P.dt : P.c2;
    int underflow = 0;
    int prev = 0; // buffer holding the current slips

    // Blocks beyond N (ragged last slice) are clamped onto block N-1: they redo its arithmetic
    // but never store, so the hot loop has no per-block branches and the B independent
    // floating-point chains of a thread interleave.

    // ---- positions (detail.h:1549): purely local; ghosts keep the line periodic
    // (the new slips also stay in the caller's registers: the well test and the forces of the
    // thread's own blocks need no reload after the barrier)
    // UREG (full nearest-neighbour lines): a thread carries the slips of its blocks in registers
    // from step to step. The fixed-step loop then only publishes the two edge blocks of a thread
    // for its neighbours; the stop modes still store all of them (a stop falls back to the slips
    // of the step before, which are only in shared memory by then).
    constexpr bool UREG = FQSB_UREG && NN1D && FULL;
    auto phase1 = [&](const int oprev, const int ocur, double (&un)[B]) {
        const double* uprev = us + oprev;
        double* ucur = us + ocur;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = POF(j);
            const int pc = (FULL || p < N) ? p : N - 1;
            un[j] = (UREG ? un[j] : uprev[SLOT(pc)]) + P.dt * v[j] + c2 * a[j];
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = POF(j);
            if (UREG && !STOP && j != 0 && j != B - 1) {
                continue;
            }
            if (FULL || p < N) {
                ucur[SLOT(p)] = un[j];
                if (ONE_D && !NN1D) {
                    if (j == 0 && t == 0) {
                        ucur[N + 1] = un[j];
                    }
                    if (FULL ? (j == B - 1 && t == T - 1) : (p == N - 1)) {
                        ucur[0] = un[j];
                    }
                }
            }
        }
    };

    // ---- well search (detail.h:144), forces at the new positions (detail.h:1380-1386) and
    //      the Verlet tail (detail.h:1552-1565)
    // phase 2a: the new positions of this thread's blocks and the well test -- no side effects,
    // so the stop modes run it BEFORE the decision about the previous step is known (its
    // shuffle chain is still in flight then)
    // Fixed-step loops: `need` is ONE flag per thread, an OR chain through the compares (a
    // per-block bit mask costs two integer instructions per block and step); the rare path below
    // repeats the test per block. The stop modes keep the bit mask (their register allocation
    // came out 9 % slower with the flag in the FMA build and no faster in the exact one).
    auto phase2a = [&](const double (&uc)[B], unsigned& need) {
        need = 0u;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = POF(j);
            const int pc = (FULL || p < N) ? p : N - 1;
            const double l = YSMEM ? syl[pc] : yl[j];
            const double rr = YSMEM ? syr[pc] : yr[j];
            if (STOP) {
                if ((FULL || p < N) && (uc[j] > rr || !(uc[j] > l))) {
                    need |= 1u << j;
                }
            }
            else {
                need |= (unsigned)((FULL || p < N) & ((uc[j] > rr) | !(uc[j] > l)));
            }
        }
    };
    auto phase2b = [&](const int ocur, double (&uc)[B], const unsigned need, auto accumulate,
                       double& sf, double& sff, int& hops, int& dS, int& dA) {
        const double* ucur = us + ocur;
        auto U = [&](int q) { return ucur[NN1D ? SLOT(q) : q + G]; };
        // NN1D && FULL: the slots of the two edge neighbours (periodic line)
        const int slot_left = (B - 1) * T + (t == 0 ? T - 1 : t - 1);
        const int slot_right = t == T - 1 ? 0 : t + 1;
        double wl[B], wr[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = POF(j);
            const int pc = (FULL || p < N) ? p : N - 1;
            wl[j] = YSMEM ? syl[pc] : yl[j];
            wr[j] = YSMEM ? syr[pc] : yr[j];
        }
        if (need) { // rare: some block of this thread left its well
#pragma unroll
            for (int j = 0; j < B; ++j) {
                if (STOP ? (((need >> j) & 1u) != 0u)
                         : ((FULL || POF(j) < N) && (uc[j] > wr[j] || !(uc[j] > wl[j])))) {
                    const int p = POF(j);
                    double l = wl[j], rr = wr[j];
                    i64 i_before = 0;
                    int moved = 0;
                    // inline fast path -- one well to the right on a `random` landscape, no S / A
                    // tracking: what nearly every well change of a driven line is. No call, no
                    // global memory; everything else goes out of line.
                    if (HOPINL && P.dist == DIST_RANDOM && !A.track && uc[j] > rr) {
                        const u64 st = sst[p];
                        const double r2 = FQSB_XADD(
                            rr, FQSB_XADD(FQSB_XMUL(pcg_double(st), P.dpar[0]), P.dpar[1]));
                        if (!(uc[j] > r2)) {
                            sst[p] = pcg_next(st);
                            sdidx[p] += 1;
                            l = rr;
                            rr = r2;
                            moved = 1;
                        }
                    }
                    if (moved == 0) {
                        moved = hop_shared(P, uc[j], &l, &rr, sst + p, S.idx + base + p,
                                           sdidx + p, &underflow, A.track, &i_before);
                    }
                    wl[j] = l;
                    wr[j] = rr;
                    if (YSMEM) {
                        syl[p] = l;
                        syr[p] = rr;
                    }
                    else {
                        yl[j] = l;
                        yr[j] = rr;
                        if (YMID) {
                            ym[j] = 0.5 * (l + rr);
                        }
                    }
                    hops += moved != 0;
                    track_hop(A, base + p, i_before, moved, dS, dA);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = POF(j);
            const int pc = (FULL || p < N) ? p : N - 1;
            const int qi = ONE_D ? 0 : (ij[ONE_D ? 0 : j] >> 16);
            const int qj = ONE_D ? 0 : (ij[ONE_D ? 0 : j] & 0xffff);
            double fi;
            if (NN1D && FULL) {
                const double ul = j > 0 ? uc[j > 0 ? j - 1 : 0] : ucur[slot_left];
                const double ur = j < B - 1 ? uc[j < B - 1 ? j + 1 : 0] : ucur[slot_right];
                auto UE = [&](int q) { return q < pc ? ul : ur; };
                fi = f_interactions<INT, false, UNIT>(P, UE, spref, pc, qi, qj, uc[j]);
            }
            else if (NN1D) { // ragged line: periodic indices resolved per block
                fi = f_interactions<INT, true, UNIT>(P, U, spref, pc, qi, qj, uc[j]);
            }
            else {
                fi = f_interactions<INT, !ONE_D, UNIT>(P, U, spref, pc, qi, qj, uc[j]);
            }
            double fp;
            if (YMID) { // detail.h:164-169
                fp = UNIT ? (ym[j] - uc[j]) : (ym[j] - uc[j]) * P.mu;
            }
            else {
                fp = f_potential<POT, UNIT>(P, uc[j], wl[j], wr[j]);
            }
            double ff = P.k_frame * (uf - uc[j]);
            double F = ff + fp + fi;
            double f = verlet_tail<UNIT>(P, F, v[j], a[j]);
            if (decltype(accumulate)::value) {
                // residual norms (detail.h:1512-1520): one explicit FMA per term. The sums are
                // reductions -- their association already differs from the reference's sequential
                // loop -- so the contraction changes nothing that is compared bit for bit.
                sf = (FULL || p < N) ? fma(f, f, sf) : sf;
                sff = (FULL || p < N) ? fma(ff, ff, sff) : sff;
            }
        }
    };
    auto phase2 = [&](const int ocur, double (&uc)[B], auto accumulate, double& sf, double& sff,
                      int& hops, int& dS, int& dA) {
        unsigned need;
        phase2a(uc, need);
        phase2b(ocur, uc, need, accumulate, sf, sff, hops, dS, dA);
    };

    int status = ST_RUNNING;
    i64 steps_done = ctl.steps;
    const i64 nloop = A.max_steps - steps_done < A.launch_steps ? A.max_steps - steps_done
                                                                : A.launch_steps;

    if (!STOP) {
        // timeSteps / flowSteps: no stop test, one barrier per step
        double sf = 0.0, sff = 0.0;
        int hops = 0, dS = 0, dA = 0;
        double uc[B];
        if (UREG) {
#pragma unroll
            for (int j = 0; j < B; ++j) {
                uc[j] = us[prev * NS + SLOT(POF(j))];
            }
        }
        for (i64 it = 0; it < nloop; ++it) {
            // (the fixed-step variant with the inline well change IS the flowSteps variant:
            // FQSB_TRY_MODE in fqsb_resident.cu; timeSteps' loop carries no frame update at all)
            if (HOPINL && A.flow) {
                uf += A.v_frame * P.dt; // detail.h:1642
            }
            phase1(prev * NS, (prev ^ 1) * NS, uc);
            __syncthreads();
            phase2((prev ^ 1) * NS, uc, std::false_type{}, sf, sff, hops, dS, dA);
            prev ^= 1;
        }
        if (UREG) { // (the write-back below reads the slips from shared memory)
            // the two edge slips are already there (published in phase 1; neighbours may still
            // be reading them)
#pragma unroll
            for (int j = 1; j < B - 1; ++j) {
                us[prev * NS + SLOT(POF(j))] = uc[j];
            }
        }
        steps_done += nloop;
        if (steps_done >= A.max_steps) {
            status = ST_EXHAUSTED;
        }
        if (t == 0) {
            ctl.inc += nloop; // detail.h:1541
            ctl.steps = steps_done;
            ctl.status = status;
            S.u_frame[r] = uf;
        }
    }
    else {
        // The decision is taken redundantly by every warp (one barrier per step), so whatever it
        // keeps alive competes with the 8 force chains for registers. Hence: the StopList ring is
        // a CIRCULAR buffer over the lanes (no roll), counters are 32-bit loop-relative, S / A run
        // as deltas, and everything that is only needed once (increment and step numbers, the
        // residual) is re-read from the control block after the loop. The activity timestamps
        // (detail.h:1770-1776) are written by thread 0 on the rare steps that change S.
        RingEntry ring = ring_load(ctl, A, lane); // lane l = logical entry l (oldest first)
        const int nring = A.niter_tol;
        int head = 0; // lane holding the OLDEST entry = the next one to be overwritten
        int dS_run = 0, dA_run = 0;
        // truncation thresholds relative to the totals at entry (detail.h:1879-1885)
        const i64 A_left = A.A_truncate > 0 ? A.A_truncate - ctl.A : (i64)1 << 40;
        const i64 S_left = A.S_truncate > 0 ? A.S_truncate - ctl.S : (i64)1 << 40;
        const bool fresh = steps_done == 0 && ctl.S != 0; // "s != s_n" at the first step (s_n = 0)
        const int nl = (int)nloop;
        const double tol4 = A.tol2 * A.tol2;
        int its = 0;
        // One barrier per step: after the forces of step s, the positions of step s+1 are
        // computed speculatively (purely local, into the other slip buffer); the barrier that
        // publishes them also publishes the partial sums of step s, whose stop decision is
        // taken right after it. A stop discards the speculative positions.
        // One barrier per step, and the decision off the critical path: after the forces of step
        // s the positions of step s+1 are computed speculatively (purely local, into the other
        // slip buffer); the barrier that publishes them also publishes the partial sums of step
        // s. Their reduction (a chain of dependent shuffles) is ISSUED right after the barrier but
        // consumed only after the side-effect-free first part of step s+1 (phase2a), so its
        // latency hides behind useful work. A stop discards the speculative positions.
        double gsf = 0.0, gsff = 0.0; // sums of the step whose decision is pending
        // Residual sampling (plain minimise): both criteria need ALL niter_tol newest residuals
        // below tol, so a residual >= tol at step s rules out a stop at steps s .. s+niter_tol-1.
        // While the residual is that large it is therefore only reduced every kskip <= niter_tol
        // steps; in between a thread just parks its two partial sums in shared memory (no
        // shuffles, no cross-warp sum, no decision). A sample below tol first evaluates the parked
        // steps (every residual after the last one >= tol enters the ring, in order: the ring is
        // then exact wherever it can matter) and switches back to per-step decisions.
        const int kskip = (A.mode == MODE_MINIMISE && !A.track)
                              ? (nring < FQSB_SKIP_K ? nring : FQSB_SKIP_K) : 1;
        bool skip = false;    // sampling mode
        bool pending = false; // gsf, gsff hold the sums of the step before
        int nback = 0;        // parked steps
        // ring insert: the newest entry replaces the oldest
        auto ring_push = [&](const RingEntry& e) {
            if (lane == head) {
                ring = e;
            }
            head = head + 1 == nring ? 0 : head + 1;
        };
        // evaluate the parked steps (CTA-uniform; rare)
        auto flush_parked = [&]() {
#pragma unroll 1
            for (int b = 0; b < nback; ++b) {
                const double2 x = sback[b * T + t];
                const bool odd = lane & 1;
                double keep = odd ? x.y : x.x;
                keep += __shfl_xor_sync(0xffffffffu, odd ? x.x : x.y, 1);
#pragma unroll
                for (int o = 16; o > 1; o >>= 1) {
                    keep += __shfl_xor_sync(0xffffffffu, keep, o);
                }
                if (lane < 2) {
                    rdb[b * 2 * NW + 2 * warp + lane] = keep;
                }
            }
            __syncthreads();
#pragma unroll 1
            for (int b = 0; b < nback; ++b) {
                double s1, s2;
                if (2 * NW == 32) {
                    warp_sum_interleaved(rdb[b * 2 * NW + lane], s1, s2);
                }
                else {
                    s1 = warp_sum(lane < NW ? rdb[b * 2 * NW + 2 * lane] : 0.0);
                    s2 = warp_sum(lane < NW ? rdb[b * 2 * NW + 2 * lane + 1] : 0.0);
                }
                ring_push(ring_entry(s1, s2));
            }
            nback = 0;
        };
        double uc[B]; // slips of the step being computed (from phase 1, in registers)
        if (UREG) {
#pragma unroll
            for (int j = 0; j < B; ++j) {
                uc[j] = us[prev * NS + SLOT(POF(j))];
            }
        }
        if (nl > 0) {
            phase1(prev * NS, (prev ^ 1) * NS, uc);
            __syncthreads();
        }
        for (int it = 0;; ++it) {
            unsigned need = 0u;
            if (it < nl) {
                phase2a(uc, need);
            }
            if (it > 0) {
                its = it;
            }
            if (pending) { // ---- decision about step `it` (1-based), detail.h:1605-1619, 1764-1784
                const int* ri = redi + ((it - 1) & 1) * 4 * NW;
                const double sf = gsf, sff = gsff;
                if (sf != sf) { // NaN forces <=> NaN positions (detail.h:1567)
                    status = ST_NAN;
                    break;
                }
                if (A.mode == MODE_UNTIL_EVENT) { // detail.h:1609
                    const int hops = __reduce_add_sync(0xffffffffu, lane < NW ? ri[4 * lane] : 0);
                    if (hops > 0) {
                        status = ST_EVENT;
                        break;
                    }
                }
                const RingEntry e = ring_entry(sf, sff);
                const bool below = e.num < A.tol2 * e.den;
                if (skip) { // a sampled step
                    if (below) {
                        flush_parked();
                        skip = false;
                    }
                    else {
                        nback = 0; // no window that could stop contains the parked steps
                    }
                }
                ring_push(e);
                if (A.track) { // detail.h:1768-1778, 1863-1872
                    const int dS = __reduce_add_sync(0xffffffffu, lane < NW ? ri[4 * lane + 1] : 0);
                    const int dA = __reduce_add_sync(0xffffffffu, lane < NW ? ri[4 * lane + 2] : 0);
                    dS_run += dS;
                    dA_run += dA;
                    // s != s_n  <=>  S changed in this step (s_n starts at 0 before the first step)
                    if ((dS != 0 || (fresh && it == 1)) && t == 0) {
                        const i64 inc_now = ctl.inc + its; // (ctl.inc is advanced after the loop)
                        if (ctl.init) {
                            ctl.init = 0;
                            ctl.qs_first = inc_now;
                        }
                        ctl.qs_last = inc_now;
                    }
                }
                // Both criteria need EVERY entry below tol (all_less(tol) resp. all_less(tol^2)),
                // the newest included: while it is not, nothing can stop (no shuffles, one compare)
                if (below) {
                    // lane l's successor in time is lane l + 1 (cyclically), except for the newest
                    // entry, whose cyclic neighbour is the oldest
                    const int succ = lane + 1 == nring ? 0 : lane + 1;
                    RingEntry nxt;
                    nxt.num = __shfl_sync(0xffffffffu, ring.num, succ & 31);
                    nxt.den = __shfl_sync(0xffffffffu, ring.den, succ & 31);
                    const bool in = lane < nring;
                    // std::is_sorted(..., greater): never r_{k+1} > r_k
                    const bool desc =
                        !in || succ == head || !(nxt.num * ring.den > ring.num * nxt.den);
                    const bool less1 = !in || (ring.num < A.tol2 * ring.den); // all_less(tol)
                    const bool less2 = !in || (ring.num < tol4 * ring.den);   // all_less(tol^2)
                    const bool descending = __all_sync(0xffffffffu, desc);
                    const bool all1 = __all_sync(0xffffffffu, less1);
                    const bool all2 = __all_sync(0xffffffffu, less2);
                    if ((descending && all1) || all2) {
                        status = ST_CONVERGED;
                        break;
                    }
                }
                else if (kskip > 1) {
                    skip = true;
                    nback = 0;
                }
                if (A.mode == MODE_TRUNCATE) { // detail.h:1879-1885
                    if (dA_run >= A_left || dS_run >= S_left) {
                        status = ST_TRUNCATED;
                        break;
                    }
                }
            }
            if (it >= nl) {
                break;
            }
            double sf = 0.0, sff = 0.0;
            int hops = 0, dS = 0, dA = 0;
            phase2b((prev ^ 1) * NS, uc, need, std::true_type{}, sf, sff, hops, dS, dA);
            prev ^= 1;

            // ---- residual + index-change reductions (detail.h:1512-1520, 1609, 1863-1864)
            double* rd = red + (it & 1) * 2 * NW;
            int* ri = redi + (it & 1) * 4 * NW;
            pending = !(skip && nback < kskip - 1);
            if (!pending) { // sampling mode, not a sampled step: park the thread's sums
                sback[nback * T + t] = make_double2(sf, sff);
                ++nback;
            }
            else {
                // one butterfly for the pair: even lanes collect sf, odd lanes sff; lanes 0 / 1
                // end up with the warp's sums (no broadcast needed)
                const bool odd = lane & 1;
                double keep = odd ? sff : sf;
                keep += __shfl_xor_sync(0xffffffffu, odd ? sf : sff, 1);
#pragma unroll
                for (int o = 16; o > 1; o >>= 1) {
                    keep += __shfl_xor_sync(0xffffffffu, keep, o);
                }
                if (lane < 2) {
                    rd[2 * warp + lane] = keep;
                }
                if (A.mode == MODE_UNTIL_EVENT || A.track) {
                    hops = __reduce_add_sync(0xffffffffu, hops);
                    if (A.track) {
                        dS = __reduce_add_sync(0xffffffffu, dS);
                        dA = __reduce_add_sync(0xffffffffu, dA);
                    }
                    if (lane == 0) {
                        ri[4 * warp] = hops;
                        ri[4 * warp + 1] = dS;
                        ri[4 * warp + 2] = dA;
                    }
                }
            }
            phase1(prev * NS, (prev ^ 1) * NS, uc); // speculative
            __syncthreads();
            // issue the reduction of this step's partials; consumed in the next iteration
            if (pending) {
                if (2 * NW == 32) {
                    warp_sum_interleaved(rd[lane], gsf, gsff);
                }
                else {
                    gsf = warp_sum(lane < NW ? rd[2 * lane] : 0.0);
                    gsff = warp_sum(lane < NW ? rd[2 * lane + 1] : 0.0);
                }
            }
        }
        if (nback > 0) { // parked steps at the end of the launch: the ring must know them
            flush_parked();
        }
        const i64 steps_now = ctl.steps + its; // (re-read: not kept in registers over the loop)
        if (status == ST_RUNNING && steps_now >= A.max_steps) {
            status = ST_EXHAUSTED;
        }
        if (t < 32) {
            // back to logical order (oldest first): lane l holds logical entry (l - head) mod n
            if (lane < nring && lane < FQSB_RING) {
                const int k = lane - head + (lane < head ? nring : 0);
                ctl.ring[k] = ring.num;
                ctl.ring_den[k] = ring.den;
            }
            const int newest = head == 0 ? nring - 1 : head - 1;
            const double last_num = __shfl_sync(0xffffffffu, ring.num, newest & 31);
            const double last_den = __shfl_sync(0xffffffffu, ring.den, newest & 31);
            if (lane == 0) {
                ctl.steps = steps_now;
                ctl.inc += its; // detail.h:1541
                ctl.S += dS_run;
                ctl.A += dA_run;
                ctl.s_n = ctl.S;
                ctl.status = status;
                if (its > 0) {
                    ctl.residual = residual_from_sums(last_num, last_den);
                }
            }
        }
    }

    // ---- write back; quench() on convergence (detail.h:1527-1532,1781)
    const double* ufin = us + (size_t)prev * NS;
    bool nan = false;
#pragma unroll
    for (int j = 0; j < B; ++j) {
        const int p = POF(j);
        if (p < N) {
            const bool q = status == ST_CONVERGED;
            const double uu = ufin[SLOT(p)];
            S.u[base + p] = uu;
            S.v[base + p] = q ? 0.0 : v[j];
            S.a[base + p] = q ? 0.0 : a[j];
            nan |= uu != uu;
            S.yl[base + p] = YSMEM ? syl[p] : yl[j];
            S.yr[base + p] = YSMEM ? syr[p] : yr[j];
            S.rng[base + p] = sst[p];
            if (sdidx[p] != 0) {
                S.idx[base + p] += sdidx[p];
            }
        }
    }
    if (nan) {
        S.err[1] = 1;
    }
    if (underflow) {
        S.err[0] = 1;
    }
}

// =============================================================================================
// K6 resident: overdamped no-passing sweeps (detail.h:1694-1753); Jacobi: every block reads
// only the previous sweep's neighbours. 2-D Laplace is the new generalisation (4 neighbours).
// =============================================================================================
template <int INT, int B, int T>
__global__ void __launch_bounds__(T) k_resident_nopassing(const Par P, const State S,
                                                          const RunArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = (int)P.N;
    const int r = blockIdx.x;
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    constexpr int NW = T / 32;
    constexpr bool TWO_D = INT == INT_LAPLACE2D;

    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }
    double* us = reinterpret_cast<double*>(smem_raw);      // [2][N]
    u64* sst = reinterpret_cast<u64*>(us + 2 * (size_t)N); // [N]
    double* red = reinterpret_cast<double*>(sst + N);      // [NW][2]

    const i64 base = (i64)r * P.N;
    double u[B], yl[B], yr[B];
    int didx[B], ij[B];
#pragma unroll
    for (int j = 0; j < B; ++j) {
        const int p = t + j * T;
        didx[j] = 0;
        ij[j] = 0;
        if (p < N) {
            u[j] = S.u[base + p];
            yl[j] = S.yl[base + p];
            yr[j] = S.yr[base + p];
            sst[p] = S.rng[base + p];
            us[p] = u[j];
            if (TWO_D) {
                int i = p / P.cols;
                ij[j] = (i << 16) | (p - i * P.cols);
            }
        }
        else {
            u[j] = 0.0;
            yl[j] = -1.7976931348623157e308;
            yr[j] = 1.7976931348623157e308;
        }
    }
    Prog g;
    prog_load(g, ctl);
    RingEntry ring = ring_load(ctl, A, lane);
    const double uf = S.u_frame[r];
    double last_sf = 0.0, last_sff = 0.0;
    const double k = P.k1, kf = P.k_frame, mu = P.mu;
    const double denom = (TWO_D ? 4 : 2) * k + kf + mu;
    const double rdenom = 1.0 / denom;
    int status = ST_RUNNING, underflow = 0, cur = 0;
    const i64 nloop = A.max_steps - g.steps < A.launch_steps ? A.max_steps - g.steps
                                                             : A.launch_steps;
    __syncthreads();

    for (i64 it = 0; it < nloop; ++it) {
        const double* uold = us + (size_t)cur * N;
        double* unew = us + (size_t)(cur ^ 1) * N;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = t + j * T;
            if (p < N) {
                double uneigh;
                if (!TWO_D) { // detail.h:1715-1723
                    int l = p == 0 ? N - 1 : p - 1, rr = p == N - 1 ? 0 : p + 1;
                    uneigh = uold[l] + uold[rr];
                }
                else {
                    const int R = P.rows, C = P.cols, i = ij[j] >> 16, jj = ij[j] & 0xffff;
                    int im = (i == 0 ? R - 1 : i - 1) * C, ip = (i == R - 1 ? 0 : i + 1) * C;
                    int jm = jj == 0 ? C - 1 : jj - 1, jp = jj == C - 1 ? 0 : jj + 1;
                    uneigh = uold[im + jj] + uold[ip + jj] + uold[i * C + jm] + uold[i * C + jp];
                }
                double un;
                for (;;) { // detail.h:1728-1738
                    double umin = 0.5 * (yl[j] + yr[j]);
                    un = div_by_invariant(k * uneigh + kf * uf + mu * umin, denom, rdenom);
                    if (!(un > yr[j] || !(un > yl[j])) || un != un) {
                        break;
                    }
                    u64 st = sst[p];
                    int moved = well_align(P, un, yl[j], yr[j], st,
                                           S.idx[base + p] + didx[j], &underflow);
                    sst[p] = st;
                    didx[j] += moved;
                    if (moved == 0) {
                        break;
                    }
                }
                u[j] = un;
                unew[p] = un;
            }
        }
        __syncthreads();
        // f = f_pot + f_int + f_frame at the new positions (detail.h:1740-1745)
        double sf = 0.0, sff = 0.0;
        auto U = [&](int q) { return unew[q]; };
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = t + j * T;
            if (p < N) {
                const double uc = u[j];
                double umin = 0.5 * (yl[j] + yr[j]);
                double ff = kf * (uf - uc);
                double fp = mu * (umin - uc);
                double fi = f_interactions<INT>(P, U, nullptr, p, ij[j] >> 16, ij[j] & 0xffff, uc);
                double f = fp + fi + ff;
                sf += f * f;
                sff += ff * ff;
            }
        }
        sf = warp_sum(sf);
        sff = warp_sum(sff);
        if (lane == 0) {
            red[2 * warp] = sf;
            red[2 * warp + 1] = sff;
        }
        __syncthreads();
        sf = warp_sum(lane < NW ? red[2 * lane] : 0.0);
        sff = warp_sum(lane < NW ? red[2 * lane + 1] : 0.0);
        cur ^= 1;
        last_sf = sf;
        last_sff = sff;
        status = step_decide(A, g, ring, lane, sf, sff, 0, 0, 0);
        if (status != ST_RUNNING) {
            break;
        }
    }

    bool nan = false;
#pragma unroll
    for (int j = 0; j < B; ++j) {
        const int p = t + j * T;
        if (p < N) {
            S.u[base + p] = u[j];
            if (status == ST_CONVERGED) { // quench(), detail.h:1749
                S.v[base + p] = 0.0;
                S.a[base + p] = 0.0;
            }
            nan |= u[j] != u[j];
            if (didx[j] != 0) {
                S.yl[base + p] = yl[j];
                S.yr[base + p] = yr[j];
                S.idx[base + p] += didx[j];
                S.rng[base + p] = sst[p];
            }
        }
    }
    if (nan) {
        S.err[1] = 1;
    }
    if (underflow) {
        S.err[0] = 1;
    }
    if (t < 32) {
        ring_store(ctl, A, lane, ring);
        if (lane == 0) {
            prog_store(g, ctl);
            ctl.status = status;
            ctl.residual = residual_from_sums(last_sf, last_sff);
        }
    }
}

// =============================================================================================
// block-level reduction helper (fixed order)
// =============================================================================================
template <int NV>
__device__ __forceinline__ void block_sum(double (&x)[NV], double* scratch /* [32*NV] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        x[k] = warp_sum(x[k]);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            scratch[warp * NV + k] = x[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        x[k] = warp_sum(lane < nw ? scratch[lane * NV + k] : 0.0);
    }
}

// =============================================================================================
// K1: streaming velocity-Verlet, one fused step per launch over all blocks of all realisations.
// Reads u,v,a,y_l,y_r once and writes u,v,a once: 64 B of DRAM traffic per block-update. The
// new state goes to the other buffer set (neighbouring CTAs still need the old one); `flip` is
// the parity of the launch within the call. In the stop modes the last CTA of a realisation to
// finish reduces the per-CTA partials in index order and takes the step's decision, so queued
// launches after the stop are no-ops.
// =============================================================================================
#define FQSB_NPART 8

// per-CTA partials -> last-CTA-done finalise of the step (one warp)
__device__ __forceinline__ void stream_finalise(const Par& P, const State& S, const RunArgs& A,
                                                int r, int flip, double uf, double (&acc)[2],
                                                int hops, int dS, int dA, double* scratch,
                                                int* iscratch, int* s_last, const int fin = 1)
{
    Ctl& ctl = S.ctl[r];
    block_sum<2>(acc, scratch);
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        hops = __reduce_add_sync(0xffffffffu, hops);
        dS = __reduce_add_sync(0xffffffffu, dS);
        dA = __reduce_add_sync(0xffffffffu, dA);
        if (lane == 0) {
            iscratch[warp * 4] = hops;
            iscratch[warp * 4 + 1] = dS;
            iscratch[warp * 4 + 2] = dA;
        }
        __syncthreads();
        hops = __reduce_add_sync(0xffffffffu, lane < nw ? iscratch[lane * 4] : 0);
        dS = __reduce_add_sync(0xffffffffu, lane < nw ? iscratch[lane * 4 + 1] : 0);
        dA = __reduce_add_sync(0xffffffffu, lane < nw ? iscratch[lane * 4 + 2] : 0);
    }
    if ((fin & 3) == 2) {
        // slab batches (fqsb_slab.inl): the per-CTA partials of step (fin >> 2) go to their own
        // slot and are added up once per batch -- no ticket, no last-CTA tail between two steps
        if (threadIdx.x == 0) {
            double* slot = A.log + ((size_t)(fin >> 2) * gridDim.x + blockIdx.x) * FQSB_NPART;
            slot[0] = acc[0];
            slot[1] = acc[1];
            slot[2] = (double)hops;
            slot[3] = (double)dS;
            slot[4] = (double)dA;
        }
        return;
    }
    double* part = S.part + ((size_t)r * gridDim.x + blockIdx.x) * FQSB_NPART;
    if (threadIdx.x == 0) {
        part[0] = acc[0];
        part[1] = acc[1];
        part[2] = (double)hops;
        part[3] = (double)dS;
        part[4] = (double)dA;
        __threadfence();
        unsigned int ticket = atomicAdd(&ctl.count, 1u);
        *s_last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (!*s_last) {
        return;
    }
    // the last CTA of the realisation: all its threads gather the per-CTA partials (thread t
    // takes CTAs t, t + blockDim, ... in order), then one fixed-order block reduction
    __threadfence();
    double tot[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const volatile double* all = S.part + (size_t)r * gridDim.x * FQSB_NPART;
    for (int c = threadIdx.x; c < (int)gridDim.x; c += blockDim.x) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            tot[k] += all[c * FQSB_NPART + k];
        }
    }
    block_sum<5>(tot, scratch);
    if (threadIdx.x >= 32) {
        return;
    }
    const int lane = threadIdx.x;
    const double sf = tot[0], sff = tot[1], dh = tot[2], ds = tot[3], da = tot[4];
    Prog g;
    prog_load(g, ctl);
    g.inc++;
    RingEntry ring = ring_load(ctl, A, lane);
    int status;
    if (A.mode == MODE_LOG) {
        status = lane == 0 ? step_log(A, r, g, sf, sff, dh, ds, da) : ST_RUNNING;
    }
    else {
        status = step_decide(A, g, ring, lane, sf, sff, (int)dh, (int)ds, (int)da);
        ring_store(ctl, A, lane, ring);
    }
    if (lane == 0) {
        prog_store(g, ctl);
        ctl.residual = residual_from_sums(sf, sff);
        ctl.flip = flip ^ 1;
        ctl.count = 0u;
        S.u_frame[r] = uf;
        ctl.status = status;
    }
}

// ---- 1-D 3-point stencils: 256 threads x 8 blocks, double2 streams, halo through smem --------
#define FQSB_ST_THREADS 256
#define FQSB_ST_J 4
#define FQSB_ST_SLAB (2 * FQSB_ST_THREADS)
#define FQSB_ST_TILE (FQSB_ST_J * FQSB_ST_SLAB)

template <int POT, int INT, bool UNIT>
__global__ void __launch_bounds__(FQSB_ST_THREADS, 2)
    k_stream_1d(const __grid_constant__ Par P, const __grid_constant__ State S,
                const __grid_constant__ RunArgs A, const int flip, const int finalise)
{
    constexpr int J = FQSB_ST_J, SLAB = FQSB_ST_SLAB;
    __shared__ __align__(16) double sun[J][SLAB + 4]; // [1] left halo, [2..2+SLAB) data, then right
    __shared__ double scratch[32 * 5];
    __shared__ int iscratch[32 * 4];
    __shared__ int s_last;
    const int r = blockIdx.y;
    const int t = threadIdx.x;
    Ctl& ctl = S.ctl[r];
    const int status = ctl.status; // consumed after the state loads are in flight
    const int N = (int)P.N;
    const i64 base = (i64)r * P.N;
    const double* __restrict__ ui = (flip ? S.u2 : S.u) + base;
    const double* __restrict__ vi = (flip ? S.v2 : S.v) + base;
    const double* __restrict__ ai = (flip ? S.a2 : S.a) + base;
    double* __restrict__ uo = (flip ? S.u : S.u2) + base;
    double* __restrict__ vo = (flip ? S.v : S.v2) + base;
    double* __restrict__ ao = (flip ? S.a : S.a2) + base;
    const int tile0 = blockIdx.x * FQSB_ST_TILE;
    const double c2 = 0.5 * P.dt * P.dt;

    double2 u2[J], v2[J], a2[J], l2[J], r2[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int p0 = tile0 + j * SLAB + 2 * t;
        if (p0 < N) {
            u2[j] = *reinterpret_cast<const double2*>(ui + p0);
            v2[j] = *reinterpret_cast<const double2*>(vi + p0);
            a2[j] = *reinterpret_cast<const double2*>(ai + p0);
            l2[j] = *reinterpret_cast<const double2*>(S.yl + base + p0);
            r2[j] = *reinterpret_cast<const double2*>(S.yr + base + p0);
        }
        else {
            u2[j] = v2[j] = a2[j] = l2[j] = r2[j] = make_double2(0.0, 0.0);
        }
    }
    double uf = S.u_frame[r];
    if (A.flow) {
        uf += A.v_frame * P.dt; // detail.h:1642
    }

    // ---- positions (detail.h:1549)
#pragma unroll
    for (int j = 0; j < J; ++j) {
        u2[j].x = u2[j].x + P.dt * v2[j].x + c2 * a2[j].x;
        u2[j].y = u2[j].y + P.dt * v2[j].y + c2 * a2[j].y;
        if (tile0 + j * SLAB + 2 * t < N) { // (the slot after a ragged slab holds its right halo)
            *reinterpret_cast<double2*>(&sun[j][2 + 2 * t]) = u2[j];
        }
    }
    if (t < 2 * J) { // the two periodic neighbours just outside each slab
        const int j = t >> 1, side = t & 1;
        const int s0 = tile0 + j * SLAB;
        if (s0 < N) {
            const int cnt = N - s0 < SLAB ? N - s0 : SLAB;
            int q = side ? s0 + cnt : s0 - 1;
            q = q < 0 ? N - 1 : (q >= N ? 0 : q);
            sun[j][side ? 2 + cnt : 1] = ui[q] + P.dt * vi[q] + c2 * ai[q];
        }
    }
    if (status != ST_RUNNING) {
        return;
    }
    __syncthreads();

    double acc[2] = {0.0, 0.0};
    int hops = 0, dS = 0, dA = 0, underflow = 0;
    bool nan = false;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int p0 = tile0 + j * SLAB + 2 * t;
        if (p0 < N) {
            const double* row = &sun[j][2];
            auto U = [&](int q) { return row[q]; };
            double uc[2] = {u2[j].x, u2[j].y};
            double wl[2] = {l2[j].x, l2[j].y};
            double wr[2] = {r2[j].x, r2[j].y};
            double vv[2] = {v2[j].x, v2[j].y};
            double aa[2] = {a2[j].x, a2[j].y};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (uc[e] > wr[e] || !(uc[e] > wl[e])) { // rare: well change, detail.h:144
                    const i64 gp = base + p0 + e;
                    u64 st = S.rng[gp];
                    i64 i_before = S.idx[gp];
                    int moved = well_align(P, uc[e], wl[e], wr[e], st, i_before, &underflow);
                    S.rng[gp] = st;
                    S.idx[gp] = i_before + moved;
                    S.yl[gp] = wl[e];
                    S.yr[gp] = wr[e];
                    if (p0 + e >= A.own_lo && p0 + e < A.own_hi) {
                        hops += moved != 0;
                        track_hop(A, gp, i_before, moved, dS, dA);
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double fi = f_interactions<INT, false, UNIT>(P, U, nullptr, 2 * t + e, 0, 0, uc[e]);
                double fp = f_potential<POT, UNIT>(P, uc[e], wl[e], wr[e]);
                double ff = P.k_frame * (uf - uc[e]);
                double F = ff + fp + fi;
                double f = verlet_tail<UNIT>(P, F, vv[e], aa[e]);
                const bool own = p0 + e >= A.own_lo && p0 + e < A.own_hi;
                acc[0] += own ? f * f : 0.0;
                acc[1] += own ? ff * ff : 0.0;
                nan |= uc[e] != uc[e];
            }
            *reinterpret_cast<double2*>(uo + p0) = u2[j];
            *reinterpret_cast<double2*>(vo + p0) = make_double2(vv[0], vv[1]);
            *reinterpret_cast<double2*>(ao + p0) = make_double2(aa[0], aa[1]);
        }
    }
    if (nan) {
        S.err[1] = 1;
    }
    if (underflow) {
        S.err[0] = 1;
    }
    if (finalise) {
        stream_finalise(P, S, A, r, flip, uf, acc, hops, dS, dA, scratch, iscratch, &s_last, finalise);
    }
}

// ---- 2-D stencils (5- and 9-point): row-marching strips ---------------------------------------
// A CTA owns a strip of FQSB_S2_TX columns and marches over Par::s2_ty rows (a per-handle choice:
// with a fixed 32 rows the 4096 x 4096 interface of config #5 made 1024 CTAs = 3.46 waves of the
// 296 resident ones, i.e. the last wave ran at 46 % occupancy). The new positions
// of rows i-1, i, i+1 live in a 4-slot ring in shared memory (one barrier per row); every cell's
// u,v,a,y_l,y_r is loaded exactly once (plus the 2 halo rows per strip and 2 halo columns per
// row), as double2, one row ahead of its use.
#define FQSB_S2_THREADS 256
#define FQSB_S2_TX (2 * FQSB_S2_THREADS)
#define FQSB_S2_TY 32 // default rows per CTA (Par::s2_ty == 0)
// resident CTAs per SM of the no-passing sweep kernel. With 4 (64 registers) the three rows of
// look-ahead spilled to local memory and the kernel stalled on those reloads (ncu: long
// scoreboard 9.7 of 16 stall cycles per issue, 27 % of the DRAM peak).
#ifndef FQSB_S2_NP_CTAS
#define FQSB_S2_NP_CTAS 2
#endif
#ifndef FQSB_S2_NP_DU
#define FQSB_S2_NP_DU 3 // rows of slips in flight ahead of the row being swept (4 / 3: slower)
#define FQSB_S2_NP_DW 2 // rows of wells in flight
#endif

template <int INT, bool UNIT>
__global__ void __launch_bounds__(FQSB_S2_THREADS, 2)
    k_stream_2d(const __grid_constant__ Par P, const __grid_constant__ State S,
                const __grid_constant__ RunArgs A, const int flip, const int finalise)
{
    constexpr int TX = FQSB_S2_TX;
    const int TY = P.s2_ty > 0 ? P.s2_ty : FQSB_S2_TY;
    __shared__ __align__(16) double sun[4][TX + 4]; // [1] left halo, [2..2+TX) data, then right
    __shared__ double scratch[32 * 5];
    __shared__ int iscratch[32 * 4];
    __shared__ int s_last;
    const int t = threadIdx.x;
    const int r = blockIdx.y;
    Ctl& ctl = S.ctl[r];
    const int status = ctl.status;
    const int R_ = P.rows, C_ = P.cols;
    const int strips = (C_ + TX - 1) / TX;
    const int strip = blockIdx.x % strips, band = blockIdx.x / strips;
    const int c0 = strip * TX, row0 = band * TY;
    const int cnt = C_ - c0 < TX ? C_ - c0 : TX; // columns of this strip (even)
    const int nrow = R_ - row0 < TY ? R_ - row0 : TY;
    const i64 base = (i64)r * P.N;
    const double* __restrict__ ui = (flip ? S.u2 : S.u) + base;
    const double* __restrict__ vi = (flip ? S.v2 : S.v) + base;
    const double* __restrict__ ai = (flip ? S.a2 : S.a) + base;
    double* __restrict__ uo = (flip ? S.u : S.u2) + base;
    double* __restrict__ vo = (flip ? S.v : S.v2) + base;
    double* __restrict__ ao = (flip ? S.a : S.a2) + base;
    const double c2 = 0.5 * P.dt * P.dt;
    const bool act = 2 * t < cnt;
    const int col = c0 + 2 * t;
    // periodic halo columns of this strip
    const int hcol = t == 0 ? (c0 == 0 ? C_ - 1 : c0 - 1) : (c0 + cnt == C_ ? 0 : c0 + cnt);
    double uf = S.u_frame[r];
    if (A.flow) {
        uf += A.v_frame * P.dt;
    }
    if (status != ST_RUNNING) {
        return;
    }

    // raw state of global row `gr` (wrapped): own pair + this thread's halo column
    struct RowRegs {
        double2 u, v, a;
        double hu, hv, ha;
    };
    auto load_row = [&](int gr, RowRegs& x) {
        const int wr = gr < 0 ? gr + R_ : (gr >= R_ ? gr - R_ : gr);
        const i64 rowoff = (i64)wr * C_;
        if (act) {
            x.u = *reinterpret_cast<const double2*>(ui + rowoff + col);
            x.v = *reinterpret_cast<const double2*>(vi + rowoff + col);
            x.a = *reinterpret_cast<const double2*>(ai + rowoff + col);
        }
        if (t < 2) {
            const i64 q = rowoff + hcol;
            x.hu = ui[q];
            x.hv = vi[q];
            x.ha = ai[q];
        }
    };
    // new positions (detail.h:1549) of a loaded row -> ring slot; returns the own pair
    auto positions = [&](const RowRegs& x, int slot, double2& un) {
        if (act) {
            un.x = x.u.x + P.dt * x.v.x + c2 * x.a.x;
            un.y = x.u.y + P.dt * x.v.y + c2 * x.a.y;
            *reinterpret_cast<double2*>(&sun[slot][2 + 2 * t]) = un;
        }
        if (t < 2) {
            sun[slot][t == 0 ? 1 : 2 + cnt] = x.hu + P.dt * x.hv + c2 * x.ha;
        }
    };

    double acc[2] = {0.0, 0.0};
    int hops = 0, dS = 0, dA = 0, underflow = 0;
    bool nan = false;
    RowRegs cur, nxt, nn;
    cur.u = cur.v = cur.a = nxt.u = nxt.v = nxt.a = nn.u = nn.v = nn.a = make_double2(0.0, 0.0);
    cur.hu = cur.hv = cur.ha = nxt.hu = nxt.hv = nxt.ha = nn.hu = nn.hv = nn.ha = 0.0;
    double2 un_c = make_double2(0.0, 0.0), un_n = un_c, un_dummy = un_c;
    double2 l2 = un_c, r2 = un_c, l2n = un_c, r2n = un_c;
    {
        RowRegs top;
        top.u = top.v = top.a = make_double2(0.0, 0.0);
        top.hu = top.hv = top.ha = 0.0;
        load_row(row0 - 1, top);
        load_row(row0, cur);
        load_row(row0 + 1, nxt);
        if (act) {
            l2 = *reinterpret_cast<const double2*>(S.yl + base + (i64)row0 * C_ + col);
            r2 = *reinterpret_cast<const double2*>(S.yr + base + (i64)row0 * C_ + col);
        }
        positions(top, (row0 + 3) & 3, un_dummy);
        positions(cur, row0 & 3, un_c);
    }

    for (int i = row0; i < row0 + nrow; ++i) {
        // prefetch: raw state two rows ahead and the wells one row ahead land during this row
        if (i + 2 <= row0 + nrow) {
            load_row(i + 2, nn);
        }
        if (act && i + 1 < row0 + nrow) {
            l2n = *reinterpret_cast<const double2*>(S.yl + base + (i64)(i + 1) * C_ + col);
            r2n = *reinterpret_cast<const double2*>(S.yr + base + (i64)(i + 1) * C_ + col);
        }
        positions(nxt, (i + 1) & 3, un_n);
        const i64 rowoff = (i64)i * C_;
        const double2 v_c = cur.v, a_c = cur.a;
        __syncthreads();
        if (act) {
            const double* up = &sun[(i + 3) & 3][2]; // row i-1
            const double* mid = &sun[i & 3][2];
            const double* dn = &sun[(i + 1) & 3][2];
            double uc[2] = {un_c.x, un_c.y};
            double wl[2] = {l2.x, l2.y};
            double wr[2] = {r2.x, r2.y};
            double vv[2] = {v_c.x, v_c.y};
            double aa[2] = {a_c.x, a_c.y};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (uc[e] > wr[e] || !(uc[e] > wl[e])) { // rare: well change, detail.h:144
                    const i64 gp = base + rowoff + col + e;
                    u64 st = S.rng[gp];
                    i64 i_before = S.idx[gp];
                    int moved = well_align(P, uc[e], wl[e], wr[e], st, i_before, &underflow);
                    S.rng[gp] = st;
                    S.idx[gp] = i_before + moved;
                    S.yl[gp] = wl[e];
                    S.yr[gp] = wr[e];
                    if (rowoff + col + e >= A.own_lo && rowoff + col + e < A.own_hi) {
                        hops += moved != 0;
                        track_hop(A, gp, i_before, moved, dS, dA);
                    }
                }
            }
            const bool own = rowoff + col >= A.own_lo && rowoff + col < A.own_hi; // whole rows
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int lc = 2 * t + e;
                double fi;
                if (INT == INT_LAPLACE2D) { // detail.h:557-582 (same operand order)
                    double lap = up[lc] + dn[lc] + mid[lc - 1] + mid[lc + 1] - 4 * uc[e];
                    fi = UNIT ? lap : lap * P.k1;
                }
                else { // QuarticGradient2d, detail.h:700-711
                    const double mk4_3 = P.k2 / 3.0;
                    const double mk4_23 = 2.0 * mk4_3;
                    const double u_pj = dn[lc], u_mj = up[lc], u_cp = mid[lc + 1], u_cm = mid[lc - 1];
                    double l = u_pj + u_mj + u_cp + u_cm - 4 * uc[e];
                    double dudx = 0.5 * (u_pj - u_mj);
                    double dudy = 0.5 * (u_cp - u_cm);
                    double d2udxdy = 0.25 * (dn[lc + 1] - dn[lc - 1] - up[lc + 1] + up[lc - 1]);
                    double d2udx2 = u_pj - 2 * uc[e] + u_mj;
                    double d2udy2 = u_cp - 2 * uc[e] + u_cm;
                    fi = l * (P.k1 + mk4_3) + mk4_23 * (dudx * dudx * d2udx2 + dudy * dudy * d2udy2 +
                                                        2.0 * dudx * dudy * d2udxdy);
                }
                double fp = f_potential<POT_CUSPY, UNIT>(P, uc[e], wl[e], wr[e]);
                double ff = P.k_frame * (uf - uc[e]);
                double F = ff + fp + fi;
                double f = verlet_tail<UNIT>(P, F, vv[e], aa[e]);
                acc[0] += own ? f * f : 0.0;
                acc[1] += own ? ff * ff : 0.0;
                nan |= uc[e] != uc[e];
            }
            *reinterpret_cast<double2*>(uo + rowoff + col) = un_c;
            *reinterpret_cast<double2*>(vo + rowoff + col) = make_double2(vv[0], vv[1]);
            *reinterpret_cast<double2*>(ao + rowoff + col) = make_double2(aa[0], aa[1]);
        }
        un_c = un_n;
        cur = nxt;
        nxt = nn;
        l2 = l2n;
        r2 = r2n;
    }
    if (nan) {
        S.err[1] = 1;
    }
    if (underflow) {
        S.err[0] = 1;
    }
    if (finalise) {
        stream_finalise(P, S, A, r, flip, uf, acc, hops, dS, dA, scratch, iscratch, &s_last, finalise);
    }
}

// ---- generic fallback (2-D stencils, LongRange, odd N): one block per thread, neighbours'
//      new positions recomputed from global memory (served by L1/L2)
template <int POT, int INT, bool THERMAL = false>
__global__ void __launch_bounds__(256)
    k_stream_step(const __grid_constant__ Par P, const __grid_constant__ State S,
                  const __grid_constant__ RunArgs A, const int flip, const int finalise)
{
    __shared__ double scratch[32 * 5];
    __shared__ int iscratch[32 * 4];
    __shared__ int s_last;
    const int r = blockIdx.y;
    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }
    const int N = (int)P.N;
    const i64 base = (i64)r * P.N;
    const double* __restrict__ ui = (flip ? S.u2 : S.u) + base;
    const double* __restrict__ vi = (flip ? S.v2 : S.v) + base;
    const double* __restrict__ ai = (flip ? S.a2 : S.a) + base;
    double* __restrict__ uo = (flip ? S.u : S.u2) + base;
    double* __restrict__ vo = (flip ? S.v : S.v2) + base;
    double* __restrict__ ao = (flip ? S.a : S.a2) + base;
    const double c2 = 0.5 * P.dt * P.dt;
    double uf = S.u_frame[r];
    if (A.flow) {
        uf += A.v_frame * P.dt;
    }
    auto UN = [&](int q) { return ui[q] + P.dt * vi[q] + c2 * ai[q]; };

    double acc[2] = {0.0, 0.0};
    int hops = 0, dS = 0, dA = 0, underflow = 0;
    bool nan = false;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
        double un = UN(p);
        double yl = S.yl[base + p], yr = S.yr[base + p];
        if (un > yr || !(un > yl)) {
            u64 st = S.rng[base + p];
            i64 i_before = S.idx[base + p];
            int moved = well_align(P, un, yl, yr, st, i_before, &underflow);
            S.rng[base + p] = st;
            S.idx[base + p] = i_before + moved;
            S.yl[base + p] = yl;
            S.yr[base + p] = yr;
            if (p >= A.own_lo && p < A.own_hi) {
                hops += moved != 0;
                track_hop(A, base + p, i_before, moved, dS, dA);
            }
        }
        int i = 0, j = 0;
        if (INT == INT_LAPLACE2D || INT == INT_QUARTICGRADIENT2D) {
            i = p / P.cols;
            j = p - i * P.cols;
        }
        double fi = f_interactions<INT>(P, UN, S.pref, p, i, j, un);
        double fp = f_potential<POT>(P, un, yl, yr);
        double ff = P.k_frame * (uf - un);
        double F = ff + fp + fi;
        double v = vi[p], a = ai[p];
        double f = THERMAL ? verlet_tail_thermal(P, F, S.f_thermal[base + p], v, a)
                           : verlet_tail(P, F, v, a);
        uo[p] = un;
        vo[p] = v;
        ao[p] = a;
        const bool own = p >= A.own_lo && p < A.own_hi;
        acc[0] += own ? f * f : 0.0;
        acc[1] += own ? ff * ff : 0.0;
        nan |= un != un;
    }
    if (nan) {
        S.err[1] = 1;
    }
    if (underflow) {
        S.err[0] = 1;
    }
    if (finalise) {
        stream_finalise(P, S, A, r, flip, uf, acc, hops, dS, dA, scratch, iscratch, &s_last, finalise);
    }
}

// =============================================================================================
// K6 streaming: overdamped no-passing sweeps (detail.h:1694-1753), one fused launch per sweep.
// Launch l evaluates the residual of its INPUT configuration (= the result of sweep l-1, so the
// stop decision of sweep l-1 is taken by this launch's last CTA) and, in the same pass over the
// data, performs sweep l into the other buffer: u is read once and written once, the wells are
// read once -- 32 B per block-update. When the decision says stop, the configuration to keep is
// the input buffer; the wells the discarded sweep moved are re-aligned by the host (k_align).
// =============================================================================================
// last-CTA decision of a no-passing launch: decides sweep l-1 from the residual of the input
__device__ __forceinline__ void np_finalise(const Par& P, const State& S, const RunArgs& A,
                                            const int r, const int flip, const int first,
                                            const int do_sweep)
{
    Ctl& ctl = S.ctl[r];
    const int lane = threadIdx.x;
    int status = ST_RUNNING;
    if (!first) {
        double sf = 0.0, sff = 0.0;
        const volatile double* all = S.part + (size_t)r * gridDim.x * FQSB_NPART;
        for (int c = lane; c < (int)gridDim.x; c += 32) {
            sf += all[c * FQSB_NPART];
            sff += all[c * FQSB_NPART + 1];
        }
        sf = warp_sum(sf);
        sff = warp_sum(sff);
        Prog g;
        prog_load(g, ctl);
        RingEntry ring = ring_load(ctl, A, lane);
        if (A.mode == MODE_LOG) {
            status = lane == 0 ? step_log(A, r, g, sf, sff, 0.0, 0.0, 0.0) : ST_RUNNING;
            status = __shfl_sync(0xffffffffu, status, 0);
        }
        else {
            status = step_decide(A, g, ring, lane, sf, sff, 0, 0, 0);
            ring_store(ctl, A, lane, ring);
        }
        if (lane == 0) {
            prog_store(g, ctl);
            ctl.residual = residual_from_sums(sf, sff);
        }
    }
    if (lane == 0) {
        // a stop keeps the input buffer (the sweep of this launch is discarded)
        ctl.flip = (status == ST_RUNNING && do_sweep) ? flip ^ 1 : flip;
        ctl.count = 0u;
        ctl.status = status;
    }
}

template <int INT>
__global__ void __launch_bounds__(256)
    k_stream_np(const __grid_constant__ Par P, const __grid_constant__ State S,
                const __grid_constant__ RunArgs A, const int flip, const int first,
                const int sweep_arg)
{
    // sweep_arg: bit 0 = perform the sweep; bits 1.. = 1 + log slot of a slab batch (0: none)
    const int do_sweep = sweep_arg & 1, slot = sweep_arg >> 1;
    __shared__ double scratch[32 * 2];
    __shared__ int s_last;
    const int r = blockIdx.y;
    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }
    constexpr bool TWO_D = INT == INT_LAPLACE2D;
    const int N = (int)P.N;
    const i64 base = (i64)r * P.N;
    const double* __restrict__ uold = (flip ? S.u2 : S.u) + base;
    double* __restrict__ unew = (flip ? S.u : S.u2) + base;
    const double uf = S.u_frame[r];
    const double k = P.k1, kf = P.k_frame, mu = P.mu;
    const double denom = (TWO_D ? 4 : 2) * k + kf + mu;
    const double rdenom = 1.0 / denom;
    int underflow = 0;
    bool nan = false;
    double acc[2] = {0.0, 0.0};
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
        const double uc = uold[p];
        double uneigh, lap;
        if (!TWO_D) { // detail.h:1715-1723 and 480-486
            int l = p == 0 ? N - 1 : p - 1, rr = p == N - 1 ? 0 : p + 1;
            const double ul = uold[l], ur = uold[rr];
            uneigh = ul + ur;
            lap = ul - 2 * uc + ur;
        }
        else { // detail.h:557-582
            const int R = P.rows, C = P.cols, i = p / C, jj = p - i * C;
            int im = (i == 0 ? R - 1 : i - 1) * C, ip = (i == R - 1 ? 0 : i + 1) * C;
            int jm = jj == 0 ? C - 1 : jj - 1, jp = jj == C - 1 ? 0 : jj + 1;
            const double a = uold[im + jj], b = uold[ip + jj], c = uold[i * C + jm],
                         d = uold[i * C + jp];
            uneigh = a + b + c + d;
            lap = a + b + c + d - 4 * uc;
        }
        double yl = S.yl[base + p], yr = S.yr[base + p];
        // residual of the input configuration: f = f_pot + f_int + f_frame (detail.h:1740-1745)
        {
            double umin = 0.5 * (yl + yr);
            double ff = kf * (uf - uc);
            double fp = mu * (umin - uc);
            double fi = lap * k;
            double f = fp + fi + ff;
            const bool own = p >= A.own_lo && p < A.own_hi;
            acc[0] += own ? f * f : 0.0;
            acc[1] += own ? ff * ff : 0.0;
            nan |= uc != uc;
        }
        if (!do_sweep) { // last launch of a logged batch: residual only
            continue;
        }
        // the sweep (detail.h:1728-1741)
        double un;
        int total = 0;
        u64 st = 0;
        i64 i0 = 0;
        bool loaded = false;
        for (;;) {
            double umin = 0.5 * (yl + yr);
            un = div_by_invariant(k * uneigh + kf * uf + mu * umin, denom, rdenom);
            if (!(un > yr || !(un > yl)) || un != un) {
                break;
            }
            if (!loaded) {
                st = S.rng[base + p];
                i0 = S.idx[base + p];
                loaded = true;
            }
            int moved = well_align(P, un, yl, yr, st, i0 + total, &underflow);
            total += moved;
            if (moved == 0) {
                break;
            }
        }
        if (loaded) {
            S.rng[base + p] = st;
            S.idx[base + p] = i0 + total;
            S.yl[base + p] = yl;
            S.yr[base + p] = yr;
        }
        unew[p] = un;
    }
    if (underflow) {
        S.err[0] = 1;
    }
    if (nan) {
        S.err[1] = 1;
    }
    block_sum<2>(acc, scratch);
    if (slot) {
        // slab batches (fqsb_slab.inl): this launch's residual partials belong to sweep slot - 1;
        // they are added up once per batch -- no ticket, no last-CTA tail between two sweeps
        if (threadIdx.x == 0 && !first) {
            double* e = A.log + ((size_t)(slot - 1) * gridDim.x + blockIdx.x) * FQSB_NPART;
            e[0] = acc[0];
            e[1] = acc[1];
            e[2] = 0.0;
            e[3] = 0.0;
            e[4] = 0.0;
        }
        return;
    }
    double* part = S.part + ((size_t)r * gridDim.x + blockIdx.x) * FQSB_NPART;
    if (threadIdx.x == 0) {
        part[0] = acc[0];
        part[1] = acc[1];
        __threadfence();
        unsigned int ticket = atomicAdd(&ctl.count, 1u);
        s_last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) {
        return;
    }
    __threadfence();
    np_finalise(P, S, A, r, flip, first, do_sweep);
}

// ---- no-passing sweep on a 2-D lattice: row-marching strips like k_stream_2d (u only) -----------
// Same fused launch semantics as k_stream_np (residual of the input configuration + next sweep).
template <int CTAS> // resident CTAs per SM the register budget is sized for
__global__ void __launch_bounds__(FQSB_S2_THREADS, CTAS)
    k_stream_np_2d(const __grid_constant__ Par P, const __grid_constant__ State S,
                   const __grid_constant__ RunArgs A, const int flip, const int first,
                   const int sweep_arg)
{
    // sweep_arg: bit 0 = perform the sweep; bits 1.. = 1 + log slot of a slab batch (0: none)
    const int do_sweep = sweep_arg & 1, slot = sweep_arg >> 1;
    constexpr int TX = FQSB_S2_TX;
    const int TY = P.s2_ty_np > 0 ? P.s2_ty_np : FQSB_S2_TY;
    __shared__ __align__(16) double su[4][TX + 4]; // [1] left halo, [2..2+TX) data, then right
    __shared__ double scratch[32 * 2];
    __shared__ int s_last;
    const int t = threadIdx.x;
    const int r = blockIdx.y;
    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }
    const int R_ = P.rows, C_ = P.cols;
    const int strips = (C_ + TX - 1) / TX;
    const int strip = blockIdx.x % strips, band = blockIdx.x / strips;
    const int c0 = strip * TX, row0 = band * TY;
    const int cnt = C_ - c0 < TX ? C_ - c0 : TX;
    const int nrow = R_ - row0 < TY ? R_ - row0 : TY;
    const i64 base = (i64)r * P.N;
    const double* __restrict__ uold = (flip ? S.u2 : S.u) + base;
    double* __restrict__ unew = (flip ? S.u : S.u2) + base;
    const bool act = 2 * t < cnt;
    const int col = c0 + 2 * t;
    const int hcol = t == 0 ? (c0 == 0 ? C_ - 1 : c0 - 1) : (c0 + cnt == C_ ? 0 : c0 + cnt);
    const double uf = S.u_frame[r];
    const double k = P.k1, kf = P.k_frame, mu = P.mu;
    const double denom = 4 * k + kf + mu;
    const double rdenom = 1.0 / denom;

    struct Row {
        double2 u;
        double h;
    };
    auto load_row = [&](int gr, Row& x) {
        const int wr = gr < 0 ? gr + R_ : (gr >= R_ ? gr - R_ : gr);
        const i64 rowoff = (i64)wr * C_;
        if (act) {
            x.u = *reinterpret_cast<const double2*>(uold + rowoff + col);
        }
        if (t < 2) {
            x.h = uold[rowoff + hcol];
        }
    };
    auto publish = [&](const Row& x, int slot) {
        if (act) {
            *reinterpret_cast<double2*>(&su[slot][2 + 2 * t]) = x.u;
        }
        if (t < 2) {
            su[slot][t == 0 ? 1 : 2 + cnt] = x.h;
        }
    };

    // software pipeline in registers: slips are loaded DU rows ahead, wells DW rows ahead of their
    // use (a sweep moves only 32 B per block, so a shallow look-ahead leaves too little in
    // flight). q[k] = slips of row i + k, wl/wr[k] = wells of row i + k.
    constexpr int DU = FQSB_S2_NP_DU, DW = FQSB_S2_NP_DW;
    Row q[DU + 1];
    double2 wlq[DW + 1], wrq[DW + 1];
#pragma unroll
    for (int k = 0; k <= DU; ++k) {
        q[k].u = make_double2(0.0, 0.0);
        q[k].h = 0.0;
    }
#pragma unroll
    for (int k = 0; k <= DW; ++k) {
        wlq[k] = wrq[k] = make_double2(0.0, 0.0);
    }
    auto load_wells = [&](int gr, double2& l, double2& rr) {
        if (act && gr < row0 + nrow) {
            l = *reinterpret_cast<const double2*>(S.yl + base + (i64)gr * C_ + col);
            rr = *reinterpret_cast<const double2*>(S.yr + base + (i64)gr * C_ + col);
        }
    };
    {
        Row top;
        top.u = make_double2(0.0, 0.0);
        top.h = 0.0;
        load_row(row0 - 1, top);
#pragma unroll
        for (int k = 0; k < DU; ++k) {
            if (k <= nrow) {
                load_row(row0 + k, q[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < DW; ++k) {
            load_wells(row0 + k, wlq[k], wrq[k]);
        }
        publish(top, (row0 + 3) & 3);
        publish(q[0], row0 & 3);
    }

    double acc[2] = {0.0, 0.0};
    int underflow = 0;
    bool nan = false;
    for (int i = row0; i < row0 + nrow; ++i) {
        if (i + DU <= row0 + nrow) {
            load_row(i + DU, q[DU]);
        }
        load_wells(i + DW, wlq[DW], wrq[DW]);
        publish(q[1], (i + 1) & 3);
        const i64 rowoff = (i64)i * C_;
        __syncthreads();
        if (act) {
            const double* up = &su[(i + 3) & 3][2];
            const double* mid = &su[i & 3][2];
            const double* dn = &su[(i + 1) & 3][2];
            double ucv[2] = {q[0].u.x, q[0].u.y};
            double wl[2] = {wlq[0].x, wlq[0].y};
            double wr[2] = {wrq[0].x, wrq[0].y};
            double out[2];
            const bool own = rowoff + col >= A.own_lo && rowoff + col < A.own_hi;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int lc = 2 * t + e;
                const double uc = ucv[e];
                const double a = up[lc], b = dn[lc], c = mid[lc - 1], d = mid[lc + 1];
                const double uneigh = a + b + c + d;           // detail.h:1715-1723 (2-D)
                const double lap = a + b + c + d - 4 * uc;     // detail.h:557-582
                {
                    double umin = 0.5 * (wl[e] + wr[e]);
                    double ff = kf * (uf - uc);
                    double fp = mu * (umin - uc);
                    double fi = lap * k;
                    double f = fp + fi + ff;
                    acc[0] += own ? f * f : 0.0;
                    acc[1] += own ? ff * ff : 0.0;
                    nan |= uc != uc;
                }
                double un = uc;
                if (do_sweep) {
                    int total = 0;
                    u64 st = 0;
                    i64 i0 = 0;
                    bool loaded = false;
                    const i64 gp = base + rowoff + col + e;
                    for (;;) { // detail.h:1728-1738
                        double umin = 0.5 * (wl[e] + wr[e]);
                        un = div_by_invariant(k * uneigh + kf * uf + mu * umin, denom, rdenom);
                        if (!(un > wr[e] || !(un > wl[e])) || un != un) {
                            break;
                        }
                        if (!loaded) {
                            st = S.rng[gp];
                            i0 = S.idx[gp];
                            loaded = true;
                        }
                        int moved = well_align(P, un, wl[e], wr[e], st, i0 + total, &underflow);
                        total += moved;
                        if (moved == 0) {
                            break;
                        }
                    }
                    if (loaded) {
                        S.rng[gp] = st;
                        S.idx[gp] = i0 + total;
                        S.yl[gp] = wl[e];
                        S.yr[gp] = wr[e];
                    }
                }
                out[e] = un;
            }
            if (do_sweep) {
                *reinterpret_cast<double2*>(unew + rowoff + col) = make_double2(out[0], out[1]);
            }
        }
#pragma unroll
        for (int k = 0; k < DU; ++k) {
            q[k] = q[k + 1];
        }
#pragma unroll
        for (int k = 0; k < DW; ++k) {
            wlq[k] = wlq[k + 1];
            wrq[k] = wrq[k + 1];
        }
    }
    if (underflow) {
        S.err[0] = 1;
    }
    if (nan) {
        S.err[1] = 1;
    }
    block_sum<2>(acc, scratch);
    if (slot) {
        // slab batches (fqsb_slab.inl): this launch's residual partials belong to sweep slot - 1;
        // they are added up once per batch -- no ticket, no last-CTA tail between two sweeps
        if (threadIdx.x == 0 && !first) {
            double* e = A.log + ((size_t)(slot - 1) * gridDim.x + blockIdx.x) * FQSB_NPART;
            e[0] = acc[0];
            e[1] = acc[1];
            e[2] = 0.0;
            e[3] = 0.0;
            e[4] = 0.0;
        }
        return;
    }
    double* part = S.part + ((size_t)r * gridDim.x + blockIdx.x) * FQSB_NPART;
    if (threadIdx.x == 0) {
        part[0] = acc[0];
        part[1] = acc[1];
        __threadfence();
        unsigned int ticket = atomicAdd(&ctl.count, 1u);
        s_last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) {
        return;
    }
    __threadfence();
    np_finalise(P, S, A, r, flip, first, do_sweep);
}

// ---- no-passing sweep on a 2-D lattice, rows staged in shared memory by the TMA engine ----------
// Same arithmetic and launch semantics as k_stream_np_2d. Instead of holding the look-ahead in
// registers (3 rows of slips + 2 rows of wells per thread: 128 registers at 2 CTAs per SM, and
// still only ~57 KB in flight per SM), one thread issues cp.async.bulk copies of whole rows
// (slips, y_l, y_r of the strip: 3 x 4 KB) into a ring of S stages, S-3 rows ahead of their use;
// each stage completes on its own mbarrier. The staged slips ARE the stencil source: row i reads
// rows i-1, i, i+1 of the ring (the periodic halo columns are added by two threads), so the
// separate ring of published rows and its stores disappear too. A stage is refilled after the
// barrier that follows its last use (row q is read by rows q-1, q, q+1).
// Measured on 4096 x 4096 (tools/line2d.py, sweeps at the fixed point): 135 -> 106 us per sweep.
// Neither more stages, nor a third CTA per SM (5 stages), nor a fully barrier-free variant (warp-
// specialised producer + per-stage empty/full mbarriers, halo columns as 16-byte bulk copies)
// moved it further (106-114 us): at 5.0 TB/s the sweep sits on the same DRAM plateau as the
// register-staged 2-D Verlet step (5.3 TB/s).
#define FQSB_S2_BULK_STAGES 8

struct BulkStage {
    double u[FQSB_S2_TX + 4]; // [1] left halo, [2 .. 2 + cnt) the strip, [2 + cnt] right halo
    double yl[FQSB_S2_TX];
    double yr[FQSB_S2_TX];
};

inline size_t stream_np_2d_bulk_smem(int stages) { return sizeof(BulkStage) * (size_t)stages; }

template <int NS, int CTAS> // stages of the ring, resident CTAs per SM
__global__ void __launch_bounds__(FQSB_S2_THREADS, CTAS)
    k_stream_np_2d_bulk(const __grid_constant__ Par P, const __grid_constant__ State S,
                        const __grid_constant__ RunArgs A, const int flip, const int first,
                        const int sweep_arg)
{
    // sweep_arg: bit 0 = perform the sweep; bits 1.. = 1 + log slot of a slab batch (0: none)
    const int do_sweep = sweep_arg & 1, slot = sweep_arg >> 1;
    constexpr int TX = FQSB_S2_TX;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BulkStage* stage = reinterpret_cast<BulkStage*>(smem_raw);
    __shared__ __align__(8) u64 full[NS];
    __shared__ double scratch[32 * 2];
    __shared__ int s_last;
    const int TY = P.s2_ty_np > 0 ? P.s2_ty_np : FQSB_S2_TY;
    const int t = threadIdx.x;
    const int r = blockIdx.y;
    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }
    const int R_ = P.rows, C_ = P.cols;
    const int strips = (C_ + TX - 1) / TX;
    const int strip = blockIdx.x % strips, band = blockIdx.x / strips;
    const int c0 = strip * TX, row0 = band * TY;
    const int cnt = C_ - c0 < TX ? C_ - c0 : TX;
    const int nrow = R_ - row0 < TY ? R_ - row0 : TY;
    const i64 base = (i64)r * P.N;
    const double* __restrict__ uold = (flip ? S.u2 : S.u) + base;
    double* __restrict__ unew = (flip ? S.u : S.u2) + base;
    const bool act = 2 * t < cnt;
    const int col = c0 + 2 * t;
    const int hcol = t == 0 ? (c0 == 0 ? C_ - 1 : c0 - 1) : (c0 + cnt == C_ ? 0 : c0 + cnt);
    const double uf = S.u_frame[r];
    const double k = P.k1, kf = P.k_frame, mu = P.mu;
    const double denom = 4 * k + kf + mu;
    const double rdenom = 1.0 / denom;
    const unsigned row_bytes = (unsigned)cnt * 8u;

    // band row rr = -1 .. nrow (global row row0 + rr, wrapped) lives in stage (rr + 1) % NS
    auto wrapped = [&](int rr) {
        const int gr = row0 + rr;
        return gr < 0 ? gr + R_ : (gr >= R_ ? gr - R_ : gr);
    };
    auto issue = [&](int rr) { // one thread: slips of row rr, and its wells if it is swept here
        BulkStage& st = stage[(rr + 1) % NS];
        u64* bar = &full[(rr + 1) % NS];
        const i64 off = (i64)wrapped(rr) * C_ + c0;
        const bool wells = rr >= 0 && rr < nrow;
        mbar_arrive_expect_tx(bar, wells ? 3u * row_bytes : row_bytes);
        bulk_copy_g2s(&st.u[2], uold + off, row_bytes, bar);
        if (wells) {
            bulk_copy_g2s(st.yl, S.yl + base + off, row_bytes, bar);
            bulk_copy_g2s(st.yr, S.yr + base + off, row_bytes, bar);
        }
    };
    if (t == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1u);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (t == 0) {
        for (int rr = -1; rr < NS - 1 && rr <= nrow; ++rr) {
            issue(rr);
        }
    }
    // halo columns of the rows: two threads, plain loads two rows ahead
    double h0 = 0.0, h1 = 0.0, h2 = 0.0;
    if (t < 2) {
        h0 = uold[(i64)wrapped(-1) * C_ + hcol];
        h1 = uold[(i64)wrapped(0) * C_ + hcol];
        if (1 <= nrow) {
            h2 = uold[(i64)wrapped(1) * C_ + hcol];
        }
    }
    const int hidx = t == 0 ? 1 : 2 + cnt;
    // rows -1 and 0 must be complete (with halos) before the first row is swept
    mbar_wait(&full[0], 0u);
    mbar_wait(&full[1 % NS], 0u);
    if (t < 2) {
        stage[0].u[hidx] = h0;
        stage[1 % NS].u[hidx] = h1;
    }

    double acc[2] = {0.0, 0.0};
    int underflow = 0;
    bool nan = false;
    for (int i = 0; i < nrow; ++i) {
        const int sm = (i + 1) % NS, sp = (i + 2) % NS, su_ = i % NS;
        double h3 = 0.0;
        if (t < 2 && i + 2 <= nrow) {
            h3 = uold[(i64)wrapped(i + 2) * C_ + hcol];
        }
        // row i+1 landed? (fill number (i + 2) / NS of its stage)
        mbar_wait(&full[sp], (unsigned)(((i + 2) / NS) & 1));
        if (t < 2) {
            stage[sp].u[hidx] = h2;
        }
        __syncthreads(); // halos of row i+1 visible; every thread is done with row i-1
        if (t == FQSB_S2_THREADS - 32 && i >= 1) { // (not warp 0: it already feeds the halos)
            const int rr = i - 2 + NS; // refill the stage of row i-2 (last read by row i-1)
            if (rr <= nrow) {
                issue(rr);
            }
        }
        const i64 rowoff = (i64)(row0 + i) * C_;
        if (act) {
            const double* up = &stage[su_].u[2];
            const double* mid = &stage[sm].u[2];
            const double* dn = &stage[sp].u[2];
            const double2 ucp = *reinterpret_cast<const double2*>(&mid[2 * t]);
            const double2 wlp = *reinterpret_cast<const double2*>(&stage[sm].yl[2 * t]);
            const double2 wrp = *reinterpret_cast<const double2*>(&stage[sm].yr[2 * t]);
            double ucv[2] = {ucp.x, ucp.y};
            double wl[2] = {wlp.x, wlp.y};
            double wr[2] = {wrp.x, wrp.y};
            double out[2], uneigh[2];
            const bool own = rowoff + col >= A.own_lo && rowoff + col < A.own_hi;
            // straight-line part for both cells (their chains interleave): residual of the
            // input configuration and the common case of the sweep, the block stays in its well
            unsigned hop = 0u;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int lc = 2 * t + e;
                const double uc = ucv[e];
                const double a = up[lc], b = dn[lc], c = mid[lc - 1], d = mid[lc + 1];
                uneigh[e] = a + b + c + d;                 // detail.h:1715-1723 (2-D)
                const double lap = a + b + c + d - 4 * uc; // detail.h:557-582
                const double umin = 0.5 * (wl[e] + wr[e]);
                const double ff = kf * (uf - uc);
                const double fp = mu * (umin - uc);
                const double fi = lap * k;
                const double f = fp + fi + ff;
                acc[0] += own ? f * f : 0.0;
                acc[1] += own ? ff * ff : 0.0;
                nan |= uc != uc;
                const double un = div_by_invariant(k * uneigh[e] + kf * uf + mu * umin, denom, rdenom);
                out[e] = do_sweep ? un : uc;
                hop |= (do_sweep && (un > wr[e] || !(un > wl[e])) && !(un != un)) ? (1u << e) : 0u;
            }
            if (hop) { // rare: a block leaves its well, detail.h:1728-1738
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (!((hop >> e) & 1u)) {
                        continue;
                    }
                    const i64 gp = base + rowoff + col + e;
                    u64 st = S.rng[gp];
                    const i64 i0 = S.idx[gp];
                    int total = 0;
                    double un = out[e];
                    for (;;) {
                        int moved = well_align(P, un, wl[e], wr[e], st, i0 + total, &underflow);
                        total += moved;
                        if (moved == 0) {
                            break;
                        }
                        double umin = 0.5 * (wl[e] + wr[e]);
                        un = div_by_invariant(k * uneigh[e] + kf * uf + mu * umin, denom, rdenom);
                        if (!(un > wr[e] || !(un > wl[e])) || un != un) {
                            break;
                        }
                    }
                    S.rng[gp] = st;
                    S.idx[gp] = i0 + total;
                    S.yl[gp] = wl[e];
                    S.yr[gp] = wr[e];
                    out[e] = un;
                }
            }
            if (do_sweep) {
                *reinterpret_cast<double2*>(unew + rowoff + col) = make_double2(out[0], out[1]);
            }
        }
        h2 = h3;
    }
    if (underflow) {
        S.err[0] = 1;
    }
    if (nan) {
        S.err[1] = 1;
    }
    block_sum<2>(acc, scratch);
    if (slot) {
        // slab batches (fqsb_slab.inl): this launch's residual partials belong to sweep slot - 1;
        // they are added up once per batch -- no ticket, no last-CTA tail between two sweeps
        if (threadIdx.x == 0 && !first) {
            double* e = A.log + ((size_t)(slot - 1) * gridDim.x + blockIdx.x) * FQSB_NPART;
            e[0] = acc[0];
            e[1] = acc[1];
            e[2] = 0.0;
            e[3] = 0.0;
            e[4] = 0.0;
        }
        return;
    }
    double* part = S.part + ((size_t)r * gridDim.x + blockIdx.x) * FQSB_NPART;
    if (threadIdx.x == 0) {
        part[0] = acc[0];
        part[1] = acc[1];
        __threadfence();
        unsigned int ticket = atomicAdd(&ctl.count, 1u);
        s_last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) {
        return;
    }
    __threadfence();
    np_finalise(P, S, A, r, flip, first, do_sweep);
}

// ---- 2-D velocity-Verlet step with rows staged in shared memory by the TMA engine --------------
// Same arithmetic as k_stream_2d. The raw state of a row (u, v, a, y_l, y_r of the strip: 5 x 4 KB)
// is brought into a ring of NS stages by cp.async.bulk copies that one thread issues NS-2 rows
// ahead; each stage completes on its own mbarrier. Row q's stage is read when the new positions
// of row q are computed (iteration q-1) and by the Verlet tail of row q (iteration q), and is
// refilled after the barrier of iteration q+1. The look-ahead costs no registers (k_stream_2d
// holds three rows of u, v, a in registers: 126 registers per thread).
struct VerletStage {
    double u[FQSB_S2_TX];
    double v[FQSB_S2_TX];
    double a[FQSB_S2_TX];
    double yl[FQSB_S2_TX];
    double yr[FQSB_S2_TX];
};

inline size_t stream_2d_bulk_smem(int stages) { return sizeof(VerletStage) * (size_t)stages; }

template <int INT, bool UNIT, int NS>
__global__ void __launch_bounds__(FQSB_S2_THREADS, 2)
    k_stream_2d_bulk(const __grid_constant__ Par P, const __grid_constant__ State S,
                     const __grid_constant__ RunArgs A, const int flip, const int finalise)
{
    constexpr int TX = FQSB_S2_TX;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    VerletStage* stage = reinterpret_cast<VerletStage*>(smem_raw);
    __shared__ __align__(16) double sun[4][TX + 4]; // [1] left halo, [2..2+TX) data, then right
    __shared__ __align__(8) u64 full[NS];
    __shared__ double scratch[32 * 5];
    __shared__ int iscratch[32 * 4];
    __shared__ int s_last;
    const int TY = P.s2_ty > 0 ? P.s2_ty : FQSB_S2_TY;
    const int t = threadIdx.x;
    const int r = blockIdx.y;
    Ctl& ctl = S.ctl[r];
    const int status = ctl.status;
    const int R_ = P.rows, C_ = P.cols;
    const int strips = (C_ + TX - 1) / TX;
    const int strip = blockIdx.x % strips, band = blockIdx.x / strips;
    const int c0 = strip * TX, row0 = band * TY;
    const int cnt = C_ - c0 < TX ? C_ - c0 : TX; // columns of this strip (even)
    const int nrow = R_ - row0 < TY ? R_ - row0 : TY;
    const i64 base = (i64)r * P.N;
    const double* __restrict__ ui = (flip ? S.u2 : S.u) + base;
    const double* __restrict__ vi = (flip ? S.v2 : S.v) + base;
    const double* __restrict__ ai = (flip ? S.a2 : S.a) + base;
    double* __restrict__ uo = (flip ? S.u : S.u2) + base;
    double* __restrict__ vo = (flip ? S.v : S.v2) + base;
    double* __restrict__ ao = (flip ? S.a : S.a2) + base;
    const double c2 = 0.5 * P.dt * P.dt;
    const bool act = 2 * t < cnt;
    const int col = c0 + 2 * t;
    const int hcol = t == 0 ? (c0 == 0 ? C_ - 1 : c0 - 1) : (c0 + cnt == C_ ? 0 : c0 + cnt);
    const int hidx = t == 0 ? 1 : 2 + cnt;
    double uf = S.u_frame[r];
    if (A.flow) {
        uf += A.v_frame * P.dt;
    }
    if (status != ST_RUNNING) {
        return;
    }
    const unsigned row_bytes = (unsigned)cnt * 8u;
    auto wrapped = [&](int rr) {
        const int gr = row0 + rr;
        return gr < 0 ? gr + R_ : (gr >= R_ ? gr - R_ : gr);
    };
    // band row rr = -1 .. nrow lives in stage (rr + 1) % NS
    auto issue = [&](int rr) {
        VerletStage& st = stage[(rr + 1) % NS];
        u64* bar = &full[(rr + 1) % NS];
        const i64 off = (i64)wrapped(rr) * C_ + c0;
        const bool wells = rr >= 0 && rr < nrow;
        mbar_arrive_expect_tx(bar, (wells ? 5u : 3u) * row_bytes);
        bulk_copy_g2s(st.u, ui + off, row_bytes, bar);
        bulk_copy_g2s(st.v, vi + off, row_bytes, bar);
        bulk_copy_g2s(st.a, ai + off, row_bytes, bar);
        if (wells) {
            bulk_copy_g2s(st.yl, S.yl + base + off, row_bytes, bar);
            bulk_copy_g2s(st.yr, S.yr + base + off, row_bytes, bar);
        }
    };
    if (t == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1u);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (t == 0) {
        for (int rr = -1; rr < NS - 1 && rr <= nrow; ++rr) {
            issue(rr);
        }
    }
    // raw state of this thread's halo column in band row rr (plain loads, threads 0 and 1); the
    // new position (detail.h:1549) is formed one row later, so the loads are never waited for
    // in the row that issues them (warp 0 would hold up the barrier of every row)
    struct Halo {
        double u, v, a;
    };
    auto halo_load = [&](int rr) {
        const i64 q = (i64)wrapped(rr) * C_ + hcol;
        Halo h;
        h.u = ui[q];
        h.v = vi[q];
        h.a = ai[q];
        return h;
    };
    auto halo_position = [&](const Halo& h) { return h.u + P.dt * h.v + c2 * h.a; };
    // new positions of band row rr from its stage -> ring slot; returns the own pair
    auto positions = [&](int rr, double2& un) {
        const VerletStage& st = stage[(rr + 1) % NS];
        if (act) {
            const double2 u2 = *reinterpret_cast<const double2*>(&st.u[2 * t]);
            const double2 v2 = *reinterpret_cast<const double2*>(&st.v[2 * t]);
            const double2 a2 = *reinterpret_cast<const double2*>(&st.a[2 * t]);
            un.x = u2.x + P.dt * v2.x + c2 * a2.x;
            un.y = u2.y + P.dt * v2.y + c2 * a2.y;
            *reinterpret_cast<double2*>(&sun[(row0 + rr) & 3][2 + 2 * t]) = un;
        }
    };
    Halo hm = {0.0, 0.0, 0.0}, h0 = hm, h1 = hm, h2 = hm; // rows -1, 0, then rows i+1, i+2
    if (t < 2) {
        hm = halo_load(-1);
        h0 = halo_load(0);
        h1 = halo_load(1);
    }
    double2 un_c = make_double2(0.0, 0.0), un_n = un_c, un_dummy = un_c;
    mbar_wait(&full[0], 0u);
    mbar_wait(&full[1 % NS], 0u);
    positions(-1, un_dummy);
    positions(0, un_c);
    if (t < 2) {
        sun[(row0 + 3) & 3][hidx] = halo_position(hm);
        sun[row0 & 3][hidx] = halo_position(h0);
    }

    double acc[2] = {0.0, 0.0};
    int hops = 0, dS = 0, dA = 0, underflow = 0;
    bool nan = false;
    for (int i = 0; i < nrow; ++i) {
        const int gi = row0 + i;
        if (t < 2 && i + 2 <= nrow) {
            h2 = halo_load(i + 2);
        }
        // row i+1 landed? (fill number (i + 2) / NS of its stage)
        mbar_wait(&full[(i + 2) % NS], (unsigned)(((i + 2) / NS) & 1));
        positions(i + 1, un_n);
        if (t < 2) {
            sun[(gi + 1) & 3][hidx] = halo_position(h1); // loaded one row ago
        }
        __syncthreads(); // positions of row i+1 visible; every thread is done with row i-1
        if (t == FQSB_S2_THREADS - 32) { // (not warp 0: it already feeds the halo columns)
            const int rr = i - 1 + NS; // refill the stage of row i-1 (row -1: read in the prologue)
            if (rr <= nrow) {
                issue(rr);
            }
        }
        const i64 rowoff = (i64)gi * C_;
        if (act) {
            const VerletStage& st = stage[(i + 1) % NS];
            const double* up = &sun[(gi + 3) & 3][2]; // row i-1
            const double* mid = &sun[gi & 3][2];
            const double* dn = &sun[(gi + 1) & 3][2];
            const double2 v_c = *reinterpret_cast<const double2*>(&st.v[2 * t]);
            const double2 a_c = *reinterpret_cast<const double2*>(&st.a[2 * t]);
            const double2 l2 = *reinterpret_cast<const double2*>(&st.yl[2 * t]);
            const double2 r2 = *reinterpret_cast<const double2*>(&st.yr[2 * t]);
            double uc[2] = {un_c.x, un_c.y};
            double wl[2] = {l2.x, l2.y};
            double wr[2] = {r2.x, r2.y};
            double vv[2] = {v_c.x, v_c.y};
            double aa[2] = {a_c.x, a_c.y};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (uc[e] > wr[e] || !(uc[e] > wl[e])) { // rare: well change, detail.h:144
                    const i64 gp = base + rowoff + col + e;
                    u64 st_ = S.rng[gp];
                    i64 i_before = S.idx[gp];
                    int moved = well_align(P, uc[e], wl[e], wr[e], st_, i_before, &underflow);
                    S.rng[gp] = st_;
                    S.idx[gp] = i_before + moved;
                    S.yl[gp] = wl[e];
                    S.yr[gp] = wr[e];
                    if (rowoff + col + e >= A.own_lo && rowoff + col + e < A.own_hi) {
                        hops += moved != 0;
                        track_hop(A, gp, i_before, moved, dS, dA);
                    }
                }
            }
            const bool own = rowoff + col >= A.own_lo && rowoff + col < A.own_hi; // whole rows
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int lc = 2 * t + e;
                double fi;
                if (INT == INT_LAPLACE2D) { // detail.h:557-582 (same operand order)
                    double lap = up[lc] + dn[lc] + mid[lc - 1] + mid[lc + 1] - 4 * uc[e];
                    fi = UNIT ? lap : lap * P.k1;
                }
                else { // QuarticGradient2d, detail.h:700-711
                    const double mk4_3 = P.k2 / 3.0;
                    const double mk4_23 = 2.0 * mk4_3;
                    const double u_pj = dn[lc], u_mj = up[lc], u_cp = mid[lc + 1], u_cm = mid[lc - 1];
                    double l = u_pj + u_mj + u_cp + u_cm - 4 * uc[e];
                    double dudx = 0.5 * (u_pj - u_mj);
                    double dudy = 0.5 * (u_cp - u_cm);
                    double d2udxdy = 0.25 * (dn[lc + 1] - dn[lc - 1] - up[lc + 1] + up[lc - 1]);
                    double d2udx2 = u_pj - 2 * uc[e] + u_mj;
                    double d2udy2 = u_cp - 2 * uc[e] + u_cm;
                    fi = l * (P.k1 + mk4_3) + mk4_23 * (dudx * dudx * d2udx2 + dudy * dudy * d2udy2 +
                                                        2.0 * dudx * dudy * d2udxdy);
                }
                double fp = f_potential<POT_CUSPY, UNIT>(P, uc[e], wl[e], wr[e]);
                double ff = P.k_frame * (uf - uc[e]);
                double F = ff + fp + fi;
                double f = verlet_tail<UNIT>(P, F, vv[e], aa[e]);
                acc[0] += own ? f * f : 0.0;
                acc[1] += own ? ff * ff : 0.0;
                nan |= uc[e] != uc[e];
            }
            *reinterpret_cast<double2*>(uo + rowoff + col) = un_c;
            *reinterpret_cast<double2*>(vo + rowoff + col) = make_double2(vv[0], vv[1]);
            *reinterpret_cast<double2*>(ao + rowoff + col) = make_double2(aa[0], aa[1]);
        }
        un_c = un_n;
        h1 = h2;
    }
    if (nan) {
        S.err[1] = 1;
    }
    if (underflow) {
        S.err[0] = 1;
    }
    if (finalise) {
        stream_finalise(P, S, A, r, flip, uf, acc, hops, dS, dA, scratch, iscratch, &s_last, finalise);
    }
}

} // namespace fqsb
