// fqsb_kernels.cuh -- sm_100a kernels of the integrator hot path.
//
//  k_resident<POT,INT,B,T>      K2: one CTA = one realisation, state on chip for the whole call
//                               (timeSteps / flowSteps / timeStepsUntilEvent / minimise /
//                               minimise_truncate), stop tests without a host round trip.
//  k_resident_nopassing<..>     K6: overdamped no-passing Jacobi sweeps, same residency.
//  k_stream_step<POT,INT>       K1: one fused Verlet step per launch over all blocks, streaming
//                               u,v,a,y_l,y_r once (64 B per block-update), last-CTA finalise.
//  k_stream_sweep / _residual   K6 streaming.
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
__device__ __forceinline__ int step_decide(const RunArgs& A, Prog& g, double& ring, int lane,
                                           double sf, double sff, int hops, int dS, int dA,
                                           double* res_out)
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
    double res = residual_from_sums(sf, sff);
    *res_out = res;
    ring = ring_roll_insert(ring, res, A.niter_tol, lane);
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
    if (ring_stop(ring, A.niter_tol, lane, A.tol, A.tol2)) {
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
// Shared memory: us[2][N] slips (double-buffered), sst[N] pcg32 states, pref[N] (LongRange),
// reduction scratch.
// =============================================================================================
template <int POT, int INT, int B, int T>
__global__ void __launch_bounds__(T) k_resident(const Par P, const State S, const RunArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = (int)P.N;
    const int r = blockIdx.x;
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    constexpr int NW = T / 32;

    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }

    double* us = reinterpret_cast<double*>(smem_raw);              // [2][N]
    u64* sst = reinterpret_cast<u64*>(us + 2 * (size_t)N);         // [N]
    double* spref = reinterpret_cast<double*>(sst + N);            // [N] if LongRange
    double* red = spref + (INT == INT_LONGRANGE1D ? N : 0);        // [NW][2]
    int* redi = reinterpret_cast<int*>(red + 2 * NW);              // [NW][4]

    const i64 base = (i64)r * P.N;
    double u[B], v[B], a[B], yl[B], yr[B];
    int didx[B];
    int ij[B];

#pragma unroll
    for (int j = 0; j < B; ++j) {
        const int p = t + j * T;
        didx[j] = 0;
        ij[j] = 0;
        if (p < N) {
            u[j] = S.u[base + p];
            v[j] = S.v[base + p];
            a[j] = S.a[base + p];
            yl[j] = S.yl[base + p];
            yr[j] = S.yr[base + p];
            sst[p] = S.rng[base + p];
            if (INT == INT_LONGRANGE1D) {
                spref[p] = S.pref[p];
            }
            if (INT == INT_LAPLACE2D || INT == INT_QUARTICGRADIENT2D) {
                int i = p / P.cols;
                ij[j] = (i << 16) | (p - i * P.cols);
            }
        }
        else {
            u[j] = v[j] = a[j] = 0.0;
            yl[j] = -1.7976931348623157e308;
            yr[j] = 1.7976931348623157e308;
        }
    }

    Prog g;
    prog_load(g, ctl);
    double ring = (lane < A.niter_tol && lane < FQSB_RING) ? ctl.ring[lane] : 0.0;
    double uf = S.u_frame[r];
    double res_last = ctl.residual;
    const double c2 = 0.5 * P.dt * P.dt; // (0.5*dt)*dt, detail.h:1549
    int status = ST_RUNNING;
    int underflow = 0;
    int cur = 0;
    const i64 nloop = A.max_steps - g.steps < A.launch_steps ? A.max_steps - g.steps
                                                             : A.launch_steps;

    for (i64 it = 0; it < nloop; ++it) {
        g.inc++; // detail.h:1541
        if (A.flow) {
            uf += A.v_frame * P.dt; // detail.h:1642
        }
        double* ucur = us + (size_t)cur * N;
        int hops = 0, dS = 0, dA = 0;

        // ---- positions (detail.h:1549) + well search (detail.h:144)
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = t + j * T;
            if (p < N) {
                double un = u[j] + P.dt * v[j] + c2 * a[j];
                u[j] = un;
                ucur[p] = un;
                if (un > yr[j] || !(un > yl[j])) {
                    u64 st = sst[p];
                    i64 i_before = S.idx[base + p] + didx[j];
                    int moved = well_align(P, un, yl[j], yr[j], st, i_before, &underflow);
                    sst[p] = st;
                    didx[j] += moved;
                    hops += moved != 0;
                    track_hop(A, base + p, i_before, moved, dS, dA);
                }
            }
        }
        __syncthreads();

        // ---- forces at the new positions (detail.h:1380-1386) + Verlet tail (1552-1565)
        double sf = 0.0, sff = 0.0;
        auto U = [&](int q) { return ucur[q]; };
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = t + j * T;
            if (p < N) {
                const double uc = u[j];
                double fi = f_interactions<INT>(P, U, spref, p, ij[j] >> 16, ij[j] & 0xffff, uc);
                double fp = f_potential<POT>(P, uc, yl[j], yr[j]);
                double ff = P.k_frame * (uf - uc);
                double F = ff + fp + fi;
                double f = verlet_tail(P, F, v[j], a[j]);
                sf += f * f;
                sff += ff * ff;
            }
        }

        if (A.mode == MODE_FIXED) {
            // no stop test: one barrier per step thanks to the double-buffered slips
            cur ^= 1;
            g.steps++;
            continue;
        }

        // ---- residual + index-change reductions (detail.h:1512-1520, 1609, 1863-1864)
        sf = warp_sum(sf);
        sff = warp_sum(sff);
        hops = warp_sum(hops);
        if (A.track) {
            dS = warp_sum(dS);
            dA = warp_sum(dA);
        }
        if (lane == 0) {
            red[2 * warp] = sf;
            red[2 * warp + 1] = sff;
            redi[4 * warp] = hops;
            redi[4 * warp + 1] = dS;
            redi[4 * warp + 2] = dA;
        }
        __syncthreads();
        sf = lane < NW ? red[2 * lane] : 0.0;
        sff = lane < NW ? red[2 * lane + 1] : 0.0;
        hops = lane < NW ? redi[4 * lane] : 0;
        dS = lane < NW ? redi[4 * lane + 1] : 0;
        dA = lane < NW ? redi[4 * lane + 2] : 0;
        sf = warp_sum(sf);
        sff = warp_sum(sff);
        hops = warp_sum(hops);
        if (A.track) {
            dS = warp_sum(dS);
            dA = warp_sum(dA);
        }
        status = step_decide(A, g, ring, lane, sf, sff, hops, dS, dA, &res_last);
        if (status != ST_RUNNING) {
            break;
        }
        // the second barrier above also orders this step's reads of `ucur` before the next
        // step's writes, so alternating buffers is not required here (kept for uniformity)
        cur ^= 1;
    }

    if (A.mode == MODE_FIXED && g.steps >= A.max_steps) {
        status = ST_EXHAUSTED;
    }

    // ---- write back; quench() on convergence (detail.h:1527-1532,1781)
    bool nan = false;
#pragma unroll
    for (int j = 0; j < B; ++j) {
        const int p = t + j * T;
        if (p < N) {
            const bool q = status == ST_CONVERGED;
            S.u[base + p] = u[j];
            S.v[base + p] = q ? 0.0 : v[j];
            S.a[base + p] = q ? 0.0 : a[j];
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
        if (lane < A.niter_tol && lane < FQSB_RING) {
            ctl.ring[lane] = ring;
        }
        if (lane == 0) {
            prog_store(g, ctl);
            ctl.status = status;
            ctl.residual = res_last;
            S.u_frame[r] = uf;
        }
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
    double ring = (lane < A.niter_tol && lane < FQSB_RING) ? ctl.ring[lane] : 0.0;
    const double uf = S.u_frame[r];
    double res_last = ctl.residual;
    const double k = P.k1, kf = P.k_frame, mu = P.mu;
    const double denom = (TWO_D ? 4 : 2) * k + kf + mu;
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
                    un = (k * uneigh + kf * uf + mu * umin) / denom;
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
        status = step_decide(A, g, ring, lane, sf, sff, 0, 0, 0, &res_last);
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
        if (lane < A.niter_tol && lane < FQSB_RING) {
            ctl.ring[lane] = ring;
        }
        if (lane == 0) {
            prog_store(g, ctl);
            ctl.status = status;
            ctl.residual = res_last;
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
// K1: streaming velocity-Verlet, one step per launch. grid = (tiles, R), 256 threads.
// Reads u,v,a (+ the neighbours' through L1/L2), y_l, y_r; writes u,v,a to the other buffer
// set (neighbouring CTAs still need the old values): 64 B of DRAM traffic per block-update.
// The last CTA of a realisation to finish reduces the per-CTA partials in index order and
// takes the step's stop decision, so queued launches after the stop are no-ops.
// =============================================================================================
#define FQSB_NPART 8

template <int POT, int INT>
__global__ void __launch_bounds__(256) k_stream_step(const Par P, const State S, const RunArgs A)
{
    __shared__ double scratch[32 * 2];
    __shared__ int iscratch[32 * 4];
    __shared__ int s_last;
    const int r = blockIdx.y;
    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }
    const int N = (int)P.N;
    const i64 base = (i64)r * P.N;
    const int flip = ctl.flip;
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
            hops += moved != 0;
            track_hop(A, base + p, i_before, moved, dS, dA);
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
        double f = verlet_tail(P, F, v, a);
        uo[p] = un;
        vo[p] = v;
        ao[p] = a;
        acc[0] += f * f;
        acc[1] += ff * ff;
        nan |= un != un;
    }
    if (nan) {
        S.err[1] = 1;
    }
    if (underflow) {
        S.err[0] = 1;
    }

    // ---- per-CTA partials
    block_sum<2>(acc, scratch);
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        hops = warp_sum(hops);
        dS = warp_sum(dS);
        dA = warp_sum(dA);
        if (lane == 0) {
            iscratch[warp * 4] = hops;
            iscratch[warp * 4 + 1] = dS;
            iscratch[warp * 4 + 2] = dA;
        }
        __syncthreads();
        hops = warp_sum(lane < nw ? iscratch[lane * 4] : 0);
        dS = warp_sum(lane < nw ? iscratch[lane * 4 + 1] : 0);
        dA = warp_sum(lane < nw ? iscratch[lane * 4 + 2] : 0);
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
        s_last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) {
        return;
    }
    // ---- finalise the step (one warp of the last CTA)
    __threadfence();
    const int lane = threadIdx.x;
    double sf = 0.0, sff = 0.0, dh = 0.0, ds = 0.0, da = 0.0;
    const volatile double* all = S.part + (size_t)r * gridDim.x * FQSB_NPART;
    for (int c = lane; c < (int)gridDim.x; c += 32) {
        sf += all[c * FQSB_NPART];
        sff += all[c * FQSB_NPART + 1];
        dh += all[c * FQSB_NPART + 2];
        ds += all[c * FQSB_NPART + 3];
        da += all[c * FQSB_NPART + 4];
    }
    sf = warp_sum(sf);
    sff = warp_sum(sff);
    dh = warp_sum(dh);
    ds = warp_sum(ds);
    da = warp_sum(da);
    Prog g;
    prog_load(g, ctl);
    g.inc++;
    double ring = (lane < A.niter_tol && lane < FQSB_RING) ? ctl.ring[lane] : 0.0;
    double res_last = ctl.residual;
    int status = step_decide(A, g, ring, lane, sf, sff, (int)dh, (int)ds, (int)da, &res_last);
    if (lane < A.niter_tol && lane < FQSB_RING) {
        ctl.ring[lane] = ring;
    }
    if (lane == 0) {
        prog_store(g, ctl);
        ctl.residual = res_last;
        ctl.flip = flip ^ 1;
        ctl.count = 0u;
        S.u_frame[r] = uf;
        ctl.status = status;
    }
}

// =============================================================================================
// K6 streaming: sweep (u -> u2) then residual of the new configuration + stop decision.
// =============================================================================================
template <int INT>
__global__ void __launch_bounds__(256) k_stream_sweep(const Par P, const State S, const RunArgs A)
{
    const int r = blockIdx.y;
    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }
    constexpr bool TWO_D = INT == INT_LAPLACE2D;
    const int N = (int)P.N;
    const i64 base = (i64)r * P.N;
    const int flip = ctl.flip;
    const double* __restrict__ uold = (flip ? S.u2 : S.u) + base;
    double* __restrict__ unew = (flip ? S.u : S.u2) + base;
    const double uf = S.u_frame[r];
    const double k = P.k1, kf = P.k_frame, mu = P.mu;
    const double denom = (TWO_D ? 4 : 2) * k + kf + mu;
    int underflow = 0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
        double uneigh;
        if (!TWO_D) {
            int l = p == 0 ? N - 1 : p - 1, rr = p == N - 1 ? 0 : p + 1;
            uneigh = uold[l] + uold[rr];
        }
        else {
            const int R = P.rows, C = P.cols, i = p / C, jj = p - i * C;
            int im = (i == 0 ? R - 1 : i - 1) * C, ip = (i == R - 1 ? 0 : i + 1) * C;
            int jm = jj == 0 ? C - 1 : jj - 1, jp = jj == C - 1 ? 0 : jj + 1;
            uneigh = uold[im + jj] + uold[ip + jj] + uold[i * C + jm] + uold[i * C + jp];
        }
        double yl = S.yl[base + p], yr = S.yr[base + p];
        double un;
        int total = 0;
        u64 st = 0;
        bool loaded = false;
        const i64 i0 = S.idx[base + p];
        for (;;) {
            double umin = 0.5 * (yl + yr);
            un = (k * uneigh + kf * uf + mu * umin) / denom;
            if (!(un > yr || !(un > yl)) || un != un) {
                break;
            }
            if (!loaded) {
                st = S.rng[base + p];
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
}

template <int INT>
__global__ void __launch_bounds__(256) k_stream_sweep_residual(const Par P, const State S,
                                                               const RunArgs A)
{
    __shared__ double scratch[32 * 2];
    __shared__ int s_last;
    const int r = blockIdx.y;
    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }
    const int N = (int)P.N;
    const i64 base = (i64)r * P.N;
    const int flip = ctl.flip;
    const double* __restrict__ un = (flip ? S.u : S.u2) + base; // the sweep's output
    const double uf = S.u_frame[r];
    auto U = [&](int q) { return un[q]; };
    double acc[2] = {0.0, 0.0};
    bool nan = false;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
        int i = 0, j = 0;
        if (INT == INT_LAPLACE2D) {
            i = p / P.cols;
            j = p - i * P.cols;
        }
        const double uc = un[p];
        double umin = 0.5 * (S.yl[base + p] + S.yr[base + p]);
        double ff = P.k_frame * (uf - uc);
        double fp = P.mu * (umin - uc);
        double fi = f_interactions<INT>(P, U, nullptr, p, i, j, uc);
        double f = fp + fi + ff;
        acc[0] += f * f;
        acc[1] += ff * ff;
        nan |= uc != uc;
    }
    if (nan) {
        S.err[1] = 1;
    }
    block_sum<2>(acc, scratch);
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
    const int lane = threadIdx.x;
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
    double ring = (lane < A.niter_tol && lane < FQSB_RING) ? ctl.ring[lane] : 0.0;
    double res_last = ctl.residual;
    int status = step_decide(A, g, ring, lane, sf, sff, 0, 0, 0, &res_last);
    if (lane < A.niter_tol && lane < FQSB_RING) {
        ctl.ring[lane] = ring;
    }
    if (lane == 0) {
        prog_store(g, ctl);
        ctl.residual = res_last;
        ctl.flip = flip ^ 1;
        ctl.count = 0u;
        ctl.status = status;
    }
}

} // namespace fqsb
