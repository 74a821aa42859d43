// fqsb_thermal.cuh -- K2t: resident velocity-Verlet for the thermal systems
// (External = RandomNormalForcing, detail.h:881-1000; Line1d.h:261-330, 486-556; Particles.h).
//
// Same residency as k_resident (one CTA = one realisation, slips in shared memory, v, a and the
// wells in registers, pcg32 well generators in shared memory) plus, on chip for the whole call:
// the random force of every block (fth), its next-draw increment relative to the start of the
// launch (nrel), its draw period (dincs) and a table of LCG jump coefficients.
//
// timeStep() starts with m_inc++; updated_inc() (detail.h:1541-1544): every block with
// inc >= next[p] takes, IN BLOCK ORDER, the next draw of the realisation's single pcg32 stream.
// The schedule does not depend on the dynamics, so the draws of step s+1 are prepared while
// step s is integrated and cost no extra barrier:
//   before the barrier of step s : every warp publishes how many of its blocks are due at s+1
//                                  (one count per (slice j, warp); block order = (j, warp, lane))
//   after the barrier            : every warp scans the (j, warp) counts (<= 128 of them) with
//                                  shuffles, a due block gets rank = prefix + popc(ballot below
//                                  its lane), jumps the stream by `rank` through the coefficient
//                                  table and draws; the stream advances by the total.
// A thread only ever writes the fth / nrel entries of its own blocks, after it has used the old
// value in the forces of step s.
#pragma once

#include "fqsb_kernels.cuh"

namespace fqsb {

#define FQSB_TH_JUMPS 256 // tabulated LCG jumps (ranks beyond it take the O(log n) loop)

// coefficients of an n-fold LCG step: s -> am * s + ap
__device__ __forceinline__ void pcg_jump_coeffs(u64 delta, u64 inc, u64& am, u64& ap)
{
    u64 cur_mult = FQSB_PCG_MULT, cur_plus = inc;
    am = 1ULL;
    ap = 0ULL;
    while (delta > 0) {
        if (delta & 1ULL) {
            am *= cur_mult;
            ap = ap * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1ULL) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1ULL;
    }
}

// rare path of the draw, out of line: normal(mu, sigma) from the uniform r (prrng)
static __device__ __noinline__ double thermal_normal(double mean, double sigma_sqrt2, double r)
{
    return mean + sigma_sqrt2 * erfinv(2.0 * r - 1.0);
}

template <int INT, int B, int T>
__global__ void __launch_bounds__(T)
    k_resident_thermal(const __grid_constant__ Par P, const __grid_constant__ State S,
                       const __grid_constant__ RunArgs A, const __grid_constant__ Thermal TH)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = T / 32;
    constexpr int K = B * NW;        // (slice, warp) counts, in block order
    constexpr int E = (K + 31) / 32; // counts scanned per lane
    const int N = (int)P.N;
    const int NS = N + 2;
    const int r = blockIdx.x;
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;

    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }

    double* us = reinterpret_cast<double*>(smem_raw);        // [2][NS]
    u64* sst = reinterpret_cast<u64*>(us + 2 * (size_t)NS);  // [N] well generators
    double* fth = reinterpret_cast<double*>(sst + N);        // [N] random forces
    u64* jm = reinterpret_cast<u64*>(fth + N);               // [FQSB_TH_JUMPS] jump multipliers
    u64* jp = jm + FQSB_TH_JUMPS;                            // [FQSB_TH_JUMPS] jump increments
    int* nrel = reinterpret_cast<int*>(jp + FQSB_TH_JUMPS);  // [N] next - inc0 (clamped)
    int* dincs = nrel + N;                                   // [N] dinc (clamped)
    int* cnt = dincs + N;                                    // [2][K]

    const i64 base = (i64)r * P.N;
    const i64 inc0 = ctl.inc;
    double v[B], a[B], yl[B], yr[B];

    constexpr int FAR = 1 << 30;
#pragma unroll
    for (int j = 0; j < B; ++j) {
        const int p = t + j * T;
        const int pc = p < N ? p : N - 1;
        v[j] = S.v[base + pc];
        a[j] = S.a[base + pc];
        yl[j] = S.yl[base + pc];
        yr[j] = S.yr[base + pc];
        if (p < N) {
            us[p + 1] = S.u[base + p];
            sst[p] = S.rng[base + p];
            // the first updated_inc() of the call copies the external's array (detail.h:942)
            fth[p] = TH.f_ext[base + p];
            const i64 d = TH.next[base + p] - inc0;
            nrel[p] = d > FAR ? FAR : (d < -FAR ? -FAR : (int)d);
            const i64 di = TH.dinc[base + p];
            dincs[p] = di > FAR ? FAR : (di < -FAR ? -FAR : (int)di);
        }
    }
    for (int k = t; k < FQSB_TH_JUMPS; k += T) {
        u64 am, ap;
        pcg_jump_coeffs((u64)k, TH.inc_rng, am, ap);
        jm[k] = am;
        jp[k] = ap;
    }
    u64 st0 = TH.state[r]; // stream state ahead of the draws of the step being prepared

    double uf = S.u_frame[r];
    const double c2 = 0.5 * P.dt * P.dt; // (0.5*dt)*dt, detail.h:1549
    int underflow = 0;
    int prev = 0;

    i64 steps_done = ctl.steps;
    const i64 nloop = A.max_steps - steps_done < A.launch_steps ? A.max_steps - steps_done
                                                                : A.launch_steps;

    // ---- which of my blocks are due at relative increment `rel` (bit j), counts published
    auto publish_due = [&](const int rel, const int buf) -> unsigned {
        unsigned due = 0u;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = t + j * T;
            const bool d = p < N && rel >= nrel[p];
            const unsigned bal = __ballot_sync(0xffffffffu, d);
            if (lane == 0) {
                cnt[buf * K + j * NW + warp] = __popc(bal);
            }
            due |= d ? (1u << j) : 0u;
        }
        return due;
    };

    // ---- the draws of one updated_inc() (detail.h:931-943) for my due blocks
    auto draw_due = [&](const unsigned due, const int buf) {
        // exclusive prefix of the K counts, E per lane
        int c[E], pfx[E];
        int tot = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = lane * E + e;
            c[e] = k < K ? cnt[buf * K + k] : 0;
            pfx[e] = tot;
            tot += c[e];
        }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) {
                incl += y;
            }
        }
        const int excl = incl - tot;
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) {
            return;
        }
        unsigned mine = 0u;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const bool d = (due >> j) & 1u;
            const unsigned bal = __ballot_sync(0xffffffffu, d);
            if (bal) { // warp-uniform
                const int k = j * NW + warp;
                const int e = k % E;
                int sel = pfx[0];
#pragma unroll
                for (int q = 1; q < E; ++q) {
                    sel = e == q ? pfx[q] : sel;
                }
                const int before = __shfl_sync(0xffffffffu, excl + sel, k / E);
                if (d) {
                    const int p = t + j * T;
                    const int rank = before + __popc(bal & ((1u << lane) - 1u));
                    const u64 st = rank < FQSB_TH_JUMPS ? jm[rank] * st0 + jp[rank]
                                                        : pcg_advance_inc(st0, (u64)rank, TH.inc_rng);
                    fth[p] = pcg_double(st); // the uniform; turned into the force below
                    // m_next += m_dinc (detail.h:938): on chip while both fit 31 bits (the
                    // exact value is written back at the end), else in global memory
                    const int nr = nrel[p], di = dincs[p];
                    if (nr > -FAR && di > -FAR && di < FAR) {
                        nrel[p] = nr + di;
                    }
                    else {
                        const i64 nx = TH.next[base + p] + TH.dinc[base + p];
                        TH.next[base + p] = nx;
                        const i64 dd = nx - inc0;
                        nrel[p] = dd > FAR ? FAR : (dd < -FAR ? -FAR : (int)dd);
                    }
                    mine |= 1u << j;
                }
            }
        }
        while (mine) { // one pass per due block of the busiest lane (usually one)
            const int j = __ffs(mine) - 1;
            mine &= mine - 1u;
            const int p = t + j * T;
            fth[p] = thermal_normal(TH.mean, TH.sigma_sqrt2, fth[p]);
        }
        st0 = total < FQSB_TH_JUMPS ? jm[total] * st0 + jp[total]
                                    : pcg_advance_inc(st0, (u64)total, TH.inc_rng);
    };

    // ---- positions (detail.h:1549); ghosts keep the line periodic
    auto phase1 = [&](const int oprev, const int ocur) {
        const double* uprev = us + oprev;
        double* ucur = us + ocur;
        double un[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = t + j * T;
            const int pc = p < N ? p : N - 1;
            un[j] = uprev[pc + 1] + P.dt * v[j] + c2 * a[j];
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = t + j * T;
            if (p < N) {
                ucur[p + 1] = un[j];
                if (p == 0) {
                    ucur[N + 1] = un[j];
                }
                if (p == N - 1) {
                    ucur[0] = un[j];
                }
            }
        }
    };

    // ---- well search, forces at the new positions, Verlet tail with the random force
    auto phase2 = [&](const int ocur) {
        const double* ucur = us + ocur;
        auto U = [&](int q) { return ucur[q + 1]; };
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = t + j * T;
            const int pc = p < N ? p : N - 1;
            const double uc = ucur[pc + 1];
            if (p < N && (uc > yr[j] || !(uc > yl[j]))) { // rare: the block left its well
                double l = yl[j], rr = yr[j];
                i64 i_before;
                hop_shared(P, uc, &l, &rr, sst + p, S.idx + base + p, &underflow, &i_before);
                yl[j] = l;
                yr[j] = rr;
            }
            const double fi = f_interactions<INT, false, false>(P, U, nullptr, pc, 0, 0, uc);
            const double fp = f_potential<POT_CUSPY, false>(P, uc, yl[j], yr[j]);
            const double ff = P.k_frame * (uf - uc);
            const double F = ff + fp + fi;
            verlet_tail_thermal(P, F, fth[pc], v[j], a[j]);
        }
    };

    // draws of the first step (inc0 + 1)
    unsigned due = 0u;
    if (nloop > 0) {
        due = publish_due(1, 0);
    }
    __syncthreads(); // state, tables and counts in place
    if (nloop > 0) {
        draw_due(due, 0);
    }
    for (i64 it = 0; it < nloop; ++it) {
        if (A.flow) {
            uf += A.v_frame * P.dt; // detail.h:1642
        }
        phase1(prev * NS, (prev ^ 1) * NS);
        const bool more = it + 1 < nloop;
        const int nbuf = (int)((it + 1) & 1);
        if (more) {
            due = publish_due((int)it + 2, nbuf);
        }
        __syncthreads();
        phase2((prev ^ 1) * NS);
        prev ^= 1;
        if (more) {
            draw_due(due, nbuf);
        }
    }
    steps_done += nloop;
    if (t == 0) {
        ctl.inc = inc0 + nloop; // detail.h:1541
        ctl.steps = steps_done;
        ctl.status = steps_done >= A.max_steps ? ST_EXHAUSTED : ST_RUNNING;
        S.u_frame[r] = uf;
        TH.state[r] = st0;
    }

    const double* ufin = us + (size_t)prev * NS;
    bool nan = false;
#pragma unroll
    for (int j = 0; j < B; ++j) {
        const int p = t + j * T;
        if (p < N) {
            const double uu = ufin[p + 1];
            S.u[base + p] = uu;
            S.v[base + p] = v[j];
            S.a[base + p] = a[j];
            nan |= uu != uu;
            S.yl[base + p] = yl[j];
            S.yr[base + p] = yr[j];
            S.rng[base + p] = sst[p];
            if (nloop > 0) {
                TH.f_ext[base + p] = fth[p];
                TH.f_sys[base + p] = fth[p];
                const int nr = nrel[p];
                if (nr > -FAR && nr < FAR) { // exact on chip (see draw_due)
                    TH.next[base + p] = inc0 + nr;
                }
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

} // namespace fqsb
