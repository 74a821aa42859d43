// fqsb_thermal.cuh -- K2t: resident velocity-Verlet for the thermal systems
// (External = RandomNormalForcing, detail.h:881-1000; Line1d.h:261-330, 486-556; Particles.h).
//
// Same residency as k_resident (one CTA = one realisation, slips in shared memory, v, a and the
// wells in registers) plus, on chip for the whole call: the random force of every block (fth),
// its next-draw increment relative to the start of the launch (nrel), its draw period (dincs)
// and a table of LCG jump coefficients. The well generators stay in global memory (they are
// touched on a well change only), which makes room for the arrays below.
//
// timeStep() starts with m_inc++; updated_inc() (detail.h:1541-1544): every block with
// inc >= next[p] takes, IN BLOCK ORDER, the next draw of the realisation's single pcg32 stream.
// The schedule does not depend on the dynamics, so the draws of step s+1 are made while step s
// is integrated, by ONE EXTRA WARP (warp specialisation: T integrator threads + 32):
//   integrator warps, before the barrier of step s: test their blocks against the schedule and
//       publish one ballot word per (slice j, warp) -- block order is (j, warp, lane);
//   producer warp, after that barrier: scans the popcounts of the <= 128 words, writes the due
//       blocks in rank order into a list, then walks the list 32 draws at a time: LCG jump by
//       the rank (tabulated coefficients), uniform -> normal (erf_inv_dev), value into fnew[parity];
//   integrator warps, after the barrier of step s+1: copy fnew into fth for the blocks they
//       know to be due, and integrate.
// The erf_inv chain thus runs beside the integration instead of in it, once per
// 32 draws instead of once per warp that owns a due block, and the step keeps its one barrier.
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

template <int INT, int B, int T, bool FULL, bool UNIT>
__global__ void __launch_bounds__(T + 32)
    k_resident_thermal(const __grid_constant__ Par P, const __grid_constant__ State S,
                       const __grid_constant__ RunArgs A, const __grid_constant__ Thermal TH)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = T / 32;       // integrator warps; warp NW is the producer
    constexpr int K = B * NW;        // (slice, warp) ballot words, in block order
    constexpr int E = (K + 31) / 32; // words scanned per producer lane
    const int N = (int)P.N;
    const int NS = N + 2;
    const int r = blockIdx.x;
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const bool producer = warp == NW;

    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }

    double* us = reinterpret_cast<double*>(smem_raw);         // [2][NS]
    double* fth = us + 2 * (size_t)NS;                        // [N] random forces in use
    double* fnew = fth + N;                                   // [2][N] draws of the coming step
    u64* jm = reinterpret_cast<u64*>(fnew + 2 * (size_t)N);   // [FQSB_TH_JUMPS] jump multipliers
    u64* jp = jm + FQSB_TH_JUMPS;                             // [FQSB_TH_JUMPS] jump increments
    int* nrel = reinterpret_cast<int*>(jp + FQSB_TH_JUMPS);   // [N] next - inc0 (clamped)
    int* dincs = nrel + N;                                    // [N] dinc (clamped)
    unsigned* bw = reinterpret_cast<unsigned*>(dincs + N);    // [2][K] ballot words
    unsigned short* dl = reinterpret_cast<unsigned short*>(bw + 2 * K); // [N] due list

    const i64 base = (i64)r * P.N;
    const i64 inc0 = ctl.inc;
    constexpr int FAR = 1 << 30;
    const i64 steps0 = ctl.steps;
    const i64 nloop = A.max_steps - steps0 < A.launch_steps ? A.max_steps - steps0
                                                            : A.launch_steps;

    if (producer) {
        // ================================ producer warp ======================================
        u64 st0 = TH.state[r]; // stream state ahead of the draws of the step being prepared
        auto produce = [&](const int par) {
            unsigned w[E];
            int pfx[E];
            int tot = 0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int k = lane * E + e;
                w[e] = k < K ? bw[par * K + k] : 0u;
                pfx[e] = tot;
                tot += __popc(w[e]);
            }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) {
                    incl += y;
                }
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0) {
                return;
            }
            const int excl = incl - tot;
#pragma unroll
            for (int e = 0; e < E; ++e) { // due blocks in rank order
                const int k = lane * E + e;
                const int p0 = (k / NW) * T + (k % NW) * 32;
                int rank = excl + pfx[e];
                unsigned word = w[e];
                while (word) {
                    const int b = __ffs(word) - 1;
                    word &= word - 1u;
                    dl[rank++] = (unsigned short)(p0 + b);
                }
            }
            __syncwarp();
            for (int k = lane; k < total; k += 32) { // detail.h:937, 32 draws at a time
                const int p = dl[k];
                const u64 st = k < FQSB_TH_JUMPS ? jm[k] * st0 + jp[k]
                                                 : pcg_advance_inc(st0, (u64)k, TH.inc_rng);
                fnew[(size_t)par * N + p] =
                    TH.mean + TH.sigma_sqrt2 * erf_inv_dev(2.0 * pcg_double(st) - 1.0);
            }
            st0 = total < FQSB_TH_JUMPS ? jm[total] * st0 + jp[total]
                                        : pcg_advance_inc(st0, (u64)total, TH.inc_rng);
            __syncwarp();
        };
        __syncthreads(); // state, tables and the ballots of the first step are in place
        if (nloop > 0) {
            produce(0);
        }
        __syncthreads();
        for (i64 it = 0; it < nloop; ++it) {
            __syncthreads();
            if (it + 1 < nloop) {
                produce((int)((it + 1) & 1));
            }
        }
        if (lane == 0) {
            TH.state[r] = st0;
        }
        return;
    }

    // ================================== integrator warps ======================================
    double v[B], a[B], yl[B], yr[B];
#pragma unroll
    for (int j = 0; j < B; ++j) {
        const int p = t + j * T;
        const int pc = (FULL || p < N) ? p : N - 1;
        v[j] = S.v[base + pc];
        a[j] = S.a[base + pc];
        yl[j] = S.yl[base + pc];
        yr[j] = S.yr[base + pc];
        if (FULL || p < N) {
            us[p + 1] = S.u[base + p];
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

    double uf = S.u_frame[r];
    const double c2 = 0.5 * P.dt * P.dt; // (0.5*dt)*dt, detail.h:1549
    int underflow = 0;
    int prev = 0;

    // ---- which of my blocks are due at relative increment `rel` (bit j): one ballot word per
    //      (slice, warp) for the producer; m_next += m_dinc (detail.h:938) for the due ones
    auto publish_due = [&](const int rel, const int par) -> unsigned {
        unsigned due = 0u;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = t + j * T;
            const bool d = (FULL || p < N) && rel >= nrel[(FULL || p < N) ? p : 0];
            const unsigned bal = __ballot_sync(0xffffffffu, d);
            if (lane == 0) {
                bw[par * K + j * NW + warp] = bal;
            }
            if (d) {
                // on chip while both fit 31 bits (the exact value is written back at the
                // end), else in global memory
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
                due |= 1u << j;
            }
        }
        return due;
    };

    // ---- positions (detail.h:1549); ghosts keep the line periodic
    auto phase1 = [&](const int oprev, const int ocur) {
        const double* uprev = us + oprev;
        double* ucur = us + ocur;
        double un[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = t + j * T;
            const int pc = (FULL || p < N) ? p : N - 1;
            un[j] = uprev[pc + 1] + P.dt * v[j] + c2 * a[j];
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = t + j * T;
            if (FULL || p < N) {
                ucur[p + 1] = un[j];
                if (j == 0 && t == 0) {
                    ucur[N + 1] = un[j];
                }
                if (FULL ? (j == B - 1 && t == T - 1) : (p == N - 1)) {
                    ucur[0] = un[j];
                }
            }
        }
    };

    // ---- well search, forces at the new positions, Verlet tail with the random force
    // (with the producer warp the register budget is 96: the B = 8 configuration runs its blocks
    // in two groups of four so that the temporaries of only four force chains are live)
    constexpr int G = B > 4 ? 4 : B;
    auto phase2_group = [&](const double* ucur, const int j0) {
        auto U = [&](int q) { return ucur[q + 1]; };
#pragma unroll
        for (int j = j0; j < j0 + G; ++j) {
            const int p = t + j * T;
            const int pc = (FULL || p < N) ? p : N - 1;
            const double uc = ucur[pc + 1];
            if ((FULL || p < N) && (uc > yr[j] || !(uc > yl[j]))) { // rare: left its well
                double l = yl[j], rr = yr[j];
                hop_global(P, uc, &l, &rr, S.rng + base + p, S.idx + base + p, &underflow);
                yl[j] = l;
                yr[j] = rr;
            }
            const double fi = f_interactions<INT, false, UNIT>(P, U, nullptr, pc, 0, 0, uc);
            const double fp = f_potential<POT_CUSPY, UNIT>(P, uc, yl[j], yr[j]);
            const double ff = P.k_frame * (uf - uc);
            const double F = ff + fp + fi;
            // (a clamped thread must not read block N-1's entry: its owner may be rewriting it)
            const double ft = (FULL || p < N) ? fth[(FULL || p < N) ? p : 0] : 0.0;
            verlet_tail_thermal<UNIT>(P, F, ft, v[j], a[j]);
        }
    };
    auto phase2 = [&](const int ocur) {
        const double* ucur = us + ocur;
#pragma unroll
        for (int j0 = 0; j0 < B; j0 += G) {
            phase2_group(ucur, j0);
            if (j0 + G < B) {
                asm volatile("" ::: "memory"); // keep the groups apart
            }
        }
    };

    unsigned due_cur = 0u; // my blocks redrawn at the step about to be integrated
    if (nloop > 0) {
        due_cur = publish_due(1, 0);
    }
    __syncthreads(); // -> producer: draws of the first step
    __syncthreads();
    for (i64 it = 0; it < nloop; ++it) {
        if (A.flow) {
            uf += A.v_frame * P.dt; // detail.h:1642
        }
        phase1(prev * NS, (prev ^ 1) * NS);
        unsigned due_next = 0u;
        if (it + 1 < nloop) {
            due_next = publish_due((int)it + 2, (int)((it + 1) & 1));
        }
        __syncthreads();
        if (due_cur) { // std::copy(m_f_thermal..., f), detail.h:942, for the redrawn blocks
            const double* fn = fnew + (size_t)(it & 1) * N;
#pragma unroll
            for (int j = 0; j < B; ++j) {
                if ((due_cur >> j) & 1u) {
                    fth[t + j * T] = fn[t + j * T];
                }
            }
        }
        phase2((prev ^ 1) * NS);
        prev ^= 1;
        due_cur = due_next;
    }
    const i64 steps_done = steps0 + nloop;
    if (t == 0) {
        ctl.inc = inc0 + nloop; // detail.h:1541
        ctl.steps = steps_done;
        ctl.status = steps_done >= A.max_steps ? ST_EXHAUSTED : ST_RUNNING;
        S.u_frame[r] = uf;
    }

    const double* ufin = us + (size_t)prev * NS;
    bool nan = false;
#pragma unroll
    for (int j = 0; j < B; ++j) {
        const int p = t + j * T;
        if (FULL || p < N) {
            const double uu = ufin[p + 1];
            S.u[base + p] = uu;
            S.v[base + p] = v[j];
            S.a[base + p] = a[j];
            nan |= uu != uu;
            S.yl[base + p] = yl[j];
            S.yr[base + p] = yr[j];
            if (nloop > 0) {
                TH.f_ext[base + p] = fth[p];
                TH.f_sys[base + p] = fth[p];
                const int nr = nrel[p];
                if (nr > -FAR && nr < FAR) { // exact on chip (see publish_due)
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
