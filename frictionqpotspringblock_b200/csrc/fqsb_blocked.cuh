// fqsb_blocked.cuh -- K2b: temporally blocked velocity-Verlet for 1-D lines that do not fit one CTA.
//
// A line of N > 4096 blocks (BASELINE config #3: N = 2^20) cannot live in the shared memory of
// one SM, and streaming it once per step (k_stream_1d) is bound by memory traffic: 64 B per
// block-update. The 3-point stencils (detail.h:474-487, 639-652, 778-804) only couple nearest
// neighbours, so information travels one block per step: a tile of `own` blocks extended by H
// halo blocks on each side can be integrated for k <= H steps WITHOUT any exchange -- after s
// steps only the outermost s blocks of each side have been contaminated by the missing
// neighbours, and those are halo copies whose owner tile computes them exactly. One launch
// therefore advances the whole line by k steps with ONE pass over HBM / L2 (all seven per-block
// arrays in, all seven out: 112 B per block per launch = 1.75 B per block-update at k = 64), and
// the inner loop is the on-chip loop of k_resident (all state of a thread's B consecutive blocks
// in registers, one slip per neighbouring thread through shared memory, pcg32 states in shared
// memory), bound by the FP64 issue rate instead of by memory. Two tiles of 256 threads share an
// SM: the load / store phase of one overlaps the steps of the other.
//
// grid = (tiles, R). The new state goes to the other buffer set (neighbouring tiles still read
// the old one as their halo), which also makes a launch revocable: in the stop modes every tile
// logs its per-step partial sums (owned blocks only; per-thread sums are parked in shared memory
// and reduced every FQSB_BK_PARK steps), the tiles' logs are added up in two levels (groups of
// FQSB_BK_GROUP tiles, then the groups; fixed order), and the last tile to finish replays the
// per-step decisions of timeStepsUntilEvent / minimise / minimise_truncate (detail.h:1605-1619,
// 1764-1784, 1858-1886) and either commits the batch (flips the buffer set) or -- the criterion
// fired at step s* < k -- leaves the input set current and asks for a batch of exactly s* steps.
// No host round trip is involved; the host only polls the status every few batches.
//
// FUSE (members of a slab-decomposed line, fixed-step batches; BlockedFuse in fqsb_device.cuh):
// the halo exchange with the neighbouring GPUs rides on the launch -- boundary tiles read the
// member's halo regions from its mailbox and write the outermost owned blocks straight into the
// neighbours' mailboxes over NVLink.
#pragma once

#include "fqsb_kernels.cuh"

namespace fqsb {

#ifndef FQSB_BK_YMID
#define FQSB_BK_YMID 1
#endif

// rare path: a block of the tile left its well. The global index is idx_in + sdidx (the delta
// accumulated during this launch); the input set is never written.
static __device__ __noinline__ int hop_blocked(const Par& P, double un, double* yl, double* yr,
                                               u64* st, const i64* gidx, int d0, int* underflow)
{
    double l = *yl, r = *yr;
    u64 s = *st;
    // (the global index is only read when the block moves left: boundary check of the landscape)
    int moved = well_align_lazy(P, un, l, r, s, [gidx, d0]() { return *gidx + d0; }, underflow);
    *yl = l;
    *yr = r;
    *st = s;
    return moved;
}

// FMA only names the instantiations of the translation units built with FMA contraction
// (-fmad=true, FQSB_FMA_BUILD; fqsb_params.kernel bit 7): same source, contracted by the compiler.
// FUSE: the instantiation whose tiles carry a slab member's halo exchange (BlockedFuse; fixed steps)
template <int POT, int INT, int B, bool UNIT, bool STOP, bool FMA = false, bool FUSE = false>
__global__ void __launch_bounds__(FQSB_BK_T, FQSB_BK_CTAS)
    k_blocked(const __grid_constant__ Par P, const __grid_constant__ State S,
              const __grid_constant__ RunArgs A, const __grid_constant__ BlockedArgs K)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int T = FQSB_BK_T, NW = T / 32, LMAX = B * T;
    __shared__ int s_last;
    const int c = blockIdx.x, r = blockIdx.y;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }
    const int flip = STOP ? ctl.flip : K.flip;
    const int nsteps = STOP ? ctl.batch : K.nsteps;
    if (nsteps <= 0) {
        return;
    }
    const int N = (int)P.N;
    const int own0 = c * K.own;
    const int cnt = N - own0 < K.own ? N - own0 : K.own;
    const int H = K.H;
    const int L = cnt + 2 * H; // local blocks q = 0..L-1 are global blocks own0 - H + q (mod N)
    auto GP = [&](int q) {
        int p = own0 - H + q;
        return p < 0 ? p + N : (p >= N ? p - N : p);
    };

    // On-chip layout: thread t holds the B CONSECUTIVE local blocks q = t*B + j with all their
    // state (u, v, a, wells) in registers, so B - 1 of the 2B stencil neighbours are the thread's
    // own registers; per step a thread only publishes its first and last slip for its two
    // neighbouring threads (double-buffered edge arrays) and reads one slip from each of them.
    double* eL = reinterpret_cast<double*>(smem_raw);          // [2][T + 2] first slip of a thread
    double* eR = eL + 2 * (T + 2);                             // [2][T + 2] last slip of a thread
    u64* sst = reinterpret_cast<u64*>(eR + 2 * (T + 2));       // [LMAX]
    // stop modes: the per-thread sums of a step are PARKED (one 16-byte store) and reduced over
    // the tile once every FQSB_BK_PARK steps, one warp per step, instead of a butterfly per warp
    // and step (whose shuffle / add chain sat on the critical path of every step's barrier)
    double2* park = reinterpret_cast<double2*>(sst + LMAX);    // [2 FQSB_BK_PARK][T]
    double* slog = reinterpret_cast<double*>(park + (STOP ? 2 * FQSB_BK_PARK * T : 0)); // [MAXSTEPS][2]
    int* ilog = reinterpret_cast<int*>(slog + FQSB_BK_MAXSTEPS * 2); // [MAXSTEPS][4] hops, dS, dA
    int* sdidx = ilog + FQSB_BK_MAXSTEPS * 4;                  // [LMAX]

    const i64 base = (i64)r * P.N;
    const double* __restrict__ ui = (flip ? S.u2 : S.u) + base;
    const double* __restrict__ vi = (flip ? S.v2 : S.v) + base;
    const double* __restrict__ ai = (flip ? S.a2 : S.a) + base;
    const double* __restrict__ yli = (flip ? K.yl2 : S.yl) + base;
    const double* __restrict__ yri = (flip ? K.yr2 : S.yr) + base;
    const i64* idxi = (flip ? K.idx2 : S.idx) + base; // (no __restrict__: fused readers patch its halo cells)
    const u64* __restrict__ rngi = (flip ? K.rng2 : S.rng) + base;
    double uf = (flip ? K.uf2 : S.u_frame)[r];

    // ---- fused halo exchange of a slab member (BlockedFuse): roles of this tile, wait for the
    //      neighbours' previous push (also before WRITING into their mailboxes: the slot this
    //      batch fills was last read by their batch before)
    const BlockedFuse& Fz = K.fuse;
    bool fz_reader = false, fz_pusher = false;
    const u64 *fz_in0 = nullptr, *fz_in1 = nullptr; // mail "from prev" -> [0, hc), "from next" -> [N - hc, N)
    if (FUSE && Fz.on) {
        blocked_tile_roles(P.N, Fz.hc, K.own, H, c, &fz_reader, &fz_pusher);
        fz_reader = fz_reader && Fz.pull;
        if (fz_reader || fz_pusher) {
            const u64 e_in = *Fz.epoch + (u64)Fz.batch;
            if (t == 0) {
                const unsigned long long t0 = global_ns();
                while (ld_acquire_sys(Fz.self) < e_in || ld_acquire_sys(Fz.self + 1) < e_in) {
                    if (global_ns() - t0 > Fz.timeout_ns) {
                        Fz.h_status[0] = 1;
                        __threadfence_system();
                        break;
                    }
                    __nanosleep(100);
                }
            }
            __syncthreads();
            fz_in0 = slab_mail(Fz.self, (int)(e_in & 1ULL), 0, Fz.hc);
            fz_in1 = slab_mail(Fz.self, (int)(e_in & 1ULL), 1, Fz.hc);
        }
    }
    // plane q (u, v, a, y_l, y_r, idx, rng) of local block gp: the state arrays, or -- halo
    // regions of a fused reader -- the mailbox
    const u64* const planes[7] = {reinterpret_cast<const u64*>(ui),  reinterpret_cast<const u64*>(vi),
                                  reinterpret_cast<const u64*>(ai),  reinterpret_cast<const u64*>(yli),
                                  reinterpret_cast<const u64*>(yri), reinterpret_cast<const u64*>(idxi),
                                  rngi};
    auto SRC = [&](const int q, const int gp) -> const u64* {
        if (FUSE && fz_reader) {
            if (gp < Fz.hc) {
                return fz_in0 + (i64)q * Fz.hc + gp;
            }
            if (gp >= N - Fz.hc) {
                return fz_in1 + (i64)q * Fz.hc + (gp - (N - Fz.hc));
            }
        }
        return planes[q] + gp;
    };
    auto LDD = [&](const int q, const int gp) {
        return __longlong_as_double((i64)(FUSE ? __ldcg(SRC(q, gp)) : *SRC(q, gp)));
    };

    double u[B], v[B], a[B], yl[B], yr[B];
    // Cuspy: midpoint of the current well beside it (force = ym - u: 1 FP64 instruction instead
    // of 3, same bits; detail.h:164-169)
    constexpr bool YMID = FQSB_BK_YMID && POT == POT_CUSPY && B <= 5; // (B > 5: the registers are worth more)
    double ym[YMID ? B : 1];
    // ownmask: blocks this tile writes back; summask: those of them that enter the sums (a member
    // of a slab-decomposed line leaves out the halo copies of its neighbours' blocks);
    // ghostmask: blocks whose right neighbour is the frozen cell just outside the tile
    unsigned ownmask = 0u, summask = 0u, ghostmask = 0u;
#pragma unroll
    for (int j = 0; j < B; ++j) {
        const int q = t * B + j;
        const int qc = q < L ? q : L - 1;
        const int gp = GP(qc);
        u[j] = LDD(0, gp);
        v[j] = LDD(1, gp);
        a[j] = LDD(2, gp);
        yl[j] = LDD(3, gp);
        yr[j] = LDD(4, gp);
        if (q < L) {
            sst[q] = FUSE ? __ldcg(SRC(6, gp)) : *SRC(6, gp);
            sdidx[q] = 0;
            if (FUSE && fz_reader && (gp < Fz.hc || gp >= N - Fz.hc)) {
                // the well index is only read on the rare path, long after this tile has released
                // the mailbox: park it in the (stale) halo cell of the input set, which nothing
                // else reads in a fused batch
                __stcg(const_cast<i64*>(idxi) + gp, (i64)__ldcg(SRC(5, gp)));
            }
        }
        else { // padding (never stored): a copy of the last block in one unbounded well
            yl[j] = -1e300;
            yr[j] = 1e300;
        }
        if (YMID) {
            ym[j] = 0.5 * (yl[j] + yr[j]);
        }
        if (q + 1 >= L) {
            ghostmask |= 1u << j;
        }
        if (q >= H && q < H + cnt) {
            ownmask |= 1u << j;
            if (gp >= A.own_lo && gp < A.own_hi) {
                summask |= 1u << j;
            }
        }
    }
    // the cells just outside the tile stay frozen at their input value: the error this makes
    // enters at the outermost halo block and moves inwards one block per step
    const double ghost_l = LDD(0, GP(-1)), ghost_r = LDD(0, GP(L));
    if (FUSE && fz_reader) { // this tile is done with the mailbox
        __threadfence();
        __syncthreads();
        if (t == 0) {
            atomicAdd(Fz.count, 1u);
        }
    }

    const double c2 = P.c2; // (0.5*dt)*dt, detail.h:1549
    int underflow = 0;
    int par = 0;

    // ---- positions (detail.h:1549); the edge slips go to the neighbouring threads
    auto phase1 = [&](const int pb) {
#pragma unroll
        for (int j = 0; j < B; ++j) {
            u[j] = u[j] + P.dt * v[j] + c2 * a[j];
        }
        eL[pb * (T + 2) + t] = u[0];
        eR[pb * (T + 2) + t] = u[B - 1];
    };

    // ---- well search (detail.h:144), forces (detail.h:1380-1386), Verlet tail (1552-1565)
    auto phase2 = [&](const int pb, auto accumulate, double& sf, double& sff, int& hops,
                      int& dS, int& dA) {
        // one flag, an OR chain through the compares (re-tested per block on the rare path); the
        // padding blocks beyond the tile sit in an unbounded well and never ask
        bool need = false;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            need |= (u[j] > yr[j]) | !(u[j] > yl[j]);
        }
        if (need) { // rare: some block of this thread left its well
#pragma unroll
            for (int j = 0; j < B; ++j) {
                if ((u[j] > yr[j]) | !(u[j] > yl[j])) {
                    const int q = t * B + j;
                    const int gp = GP(q);
                    int uflag = 0;
                    double l = yl[j], rr = yr[j];
                    const int d0 = sdidx[q];
                    int moved = 0;
                    // inline fast path: one well to the right on a `random` landscape
                    if (P.dist == DIST_RANDOM && u[j] > rr) {
                        const u64 st = sst[q];
                        const double r2 = FQSB_XADD(
                            rr, FQSB_XADD(FQSB_XMUL(pcg_double(st), P.dpar[0]), P.dpar[1]));
                        if (!(u[j] > r2)) {
                            sst[q] = pcg_next(st);
                            l = rr;
                            rr = r2;
                            moved = 1;
                        }
                    }
                    if (moved == 0) {
                        moved = hop_blocked(P, u[j], &l, &rr, sst + q, idxi + gp, d0, &uflag);
                    }
                    yl[j] = l;
                    yr[j] = rr;
                    if (YMID) {
                        ym[j] = 0.5 * (l + rr);
                    }
                    sdidx[q] += moved;
                    if ((ownmask >> j) & 1u) { // halo copies are accounted for by their owner
                        underflow |= uflag;
                    }
                    if ((summask >> j) & 1u) {
                        hops += moved != 0;
                        if (A.track) {
                            track_hop(A, base + gp, idxi[gp] + d0, moved, dS, dA);
                        }
                    }
                }
            }
        }
        // slips of the two neighbouring threads (thread T-1 never reads beyond the arrays: its
        // last block is always the last of the tile or beyond)
        const double from_left = t == 0 ? ghost_l : eR[pb * (T + 2) + t - 1];
        const double from_right = eL[pb * (T + 2) + t + 1];
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int q = t * B + j;
            const double ul = j > 0 ? u[j > 0 ? j - 1 : 0] : from_left;
            const double ur = ((ghostmask >> j) & 1u) ? ghost_r
                                                      : (j < B - 1 ? u[j < B - 1 ? j + 1 : 0] : from_right);
            auto UE = [&](int qq) { return qq < q ? ul : ur; };
            double fi = f_interactions<INT, false, UNIT>(P, UE, nullptr, q, 0, 0, u[j]);
            double fp;
            if (YMID) {
                fp = UNIT ? (ym[j] - u[j]) : (ym[j] - u[j]) * P.mu;
            }
            else {
                fp = f_potential<POT, UNIT>(P, u[j], yl[j], yr[j]);
            }
            double ff = P.k_frame * (uf - u[j]);
            double F = ff + fp + fi;
            double f = verlet_tail<UNIT>(P, F, v[j], a[j]);
            // accumulate: 0 = no sums, 1 = blocks of `summask` only (an unmasked variant for the
            // interior warps saves 16 selects per warp-step and measured no faster)
            if (decltype(accumulate)::value == 1 && ((summask >> j) & 1u)) {
                sf += f * f;
                sff += ff * ff;
            }
        }
    };

    if (!STOP) {
        double sf = 0.0, sff = 0.0;
        int hops = 0, dS = 0, dA = 0;
        for (int it = 0; it < nsteps; ++it) {
            if (A.flow) {
                uf += A.v_frame * P.dt; // detail.h:1642
            }
            phase1(par);
            __syncthreads();
            phase2(par, std::integral_constant<int, 0>{}, sf, sff, hops, dS, dA);
            par ^= 1;
        }
    }
    else {
        for (int i = t; i < FQSB_BK_MAXSTEPS * 4; i += T) {
            ilog[i] = 0;
        }
        // sums of steps s0 .. s0 + cnt - 1 over the tile: warp w takes step s0 + w (threads added
        // in a fixed order: 8 strided terms per lane, then the butterfly)
        auto reduce_parked = [&](const int s0, const int cnt) {
            if (warp < cnt) {
                const int s = s0 + warp;
                const double2* e = park + (s & (2 * FQSB_BK_PARK - 1)) * T;
                double x = 0.0, y = 0.0;
#pragma unroll
                for (int i = 0; i < T / 32; ++i) {
                    const double2 z = e[lane + 32 * i];
                    x += z.x;
                    y += z.y;
                }
                warp_sum2(x, y);
                if (lane == 0) {
                    slog[s * 2] = x;
                    slog[s * 2 + 1] = y;
                }
            }
        };
        static_assert(FQSB_BK_PARK <= NW, "one warp per parked step");
        for (int it = 0; it < nsteps; ++it) {
            double sf = 0.0, sff = 0.0;
            int hops = 0, dS = 0, dA = 0;
            phase1(par);
            __syncthreads();
            if (it > 0 && (it & (FQSB_BK_PARK - 1)) == 0) {
                // (the slots being read are rewritten FQSB_BK_PARK steps from now, i.e. after the
                // next barrier at the earliest)
                reduce_parked(it - FQSB_BK_PARK, FQSB_BK_PARK);
            }
            phase2(par, std::integral_constant<int, 1>{}, sf, sff, hops, dS, dA);
            par ^= 1;
            park[(it & (2 * FQSB_BK_PARK - 1)) * T + t] = make_double2(sf, sff);
            // well changes are rare: integer sums straight into the per-step log
            if (__any_sync(0xffffffffu, hops != 0)) {
                hops = __reduce_add_sync(0xffffffffu, hops);
                if (A.track) {
                    dS = __reduce_add_sync(0xffffffffu, dS);
                    dA = __reduce_add_sync(0xffffffffu, dA);
                }
                if (lane == 0) {
                    atomicAdd(ilog + it * 4, hops);
                    if (A.track) {
                        atomicAdd(ilog + it * 4 + 1, dS);
                        atomicAdd(ilog + it * 4 + 2, dA);
                    }
                }
            }
        }
        __syncthreads();
        {
            const int s0 = ((nsteps - 1) / FQSB_BK_PARK) * FQSB_BK_PARK;
            reduce_parked(s0, nsteps - s0);
        }
    }

    // ---- owned blocks -> the other buffer set
    {
        double* __restrict__ uo = (flip ? S.u : S.u2) + base;
        double* __restrict__ vo = (flip ? S.v : S.v2) + base;
        double* __restrict__ ao = (flip ? S.a : S.a2) + base;
        double* __restrict__ ylo = (flip ? S.yl : K.yl2) + base;
        double* __restrict__ yro = (flip ? S.yr : K.yr2) + base;
        i64* __restrict__ idxo = (flip ? S.idx : K.idx2) + base;
        u64* __restrict__ rngo = (flip ? S.rng : K.rng2) + base;
        bool nan = false;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            if ((ownmask >> j) & 1u) {
                const int q = t * B + j;
                const int gp = own0 - H + q; // owned: no wrap
                const double uu = u[j];
                uo[gp] = uu;
                vo[gp] = v[j];
                ao[gp] = a[j];
                ylo[gp] = yl[j];
                yro[gp] = yr[j];
                rngo[gp] = sst[q];
                const i64 inew = (FUSE ? __ldcg(idxi + gp) : idxi[gp]) + sdidx[q];
                idxo[gp] = inew;
                nan |= uu != uu;
                if (FUSE && fz_pusher) {
                    // the member's outermost owned blocks -> the neighbours' mailboxes (NVLink
                    // peer stores): [hc, 2 hc) becomes prev's "from next", [N - 2 hc, N - hc)
                    // next's "from prev"
                    const int par = (int)((*Fz.epoch + (u64)Fz.batch + 1ULL) & 1ULL);
                    u64* dst = nullptr;
                    if (gp >= Fz.hc && gp < 2 * Fz.hc) {
                        dst = slab_mail(Fz.prev, par, 1, Fz.hc) + (gp - Fz.hc);
                    }
                    else if (gp >= N - 2 * Fz.hc && gp < N - Fz.hc) {
                        dst = slab_mail(Fz.next, par, 0, Fz.hc) + (gp - (N - 2 * Fz.hc));
                    }
                    if (dst) {
                        dst[0] = (u64)__double_as_longlong(uu);
                        dst[Fz.hc] = (u64)__double_as_longlong(v[j]);
                        dst[2 * Fz.hc] = (u64)__double_as_longlong(a[j]);
                        dst[3 * Fz.hc] = (u64)__double_as_longlong(yl[j]);
                        dst[4 * Fz.hc] = (u64)__double_as_longlong(yr[j]);
                        dst[5 * Fz.hc] = (u64)inew;
                        dst[6 * Fz.hc] = sst[q];
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
        if (c == 0 && t == 0) {
            (flip ? S.u_frame : K.uf2)[r] = uf;
        }
    }
    if (!STOP) {
        if (FUSE && fz_pusher) {
            // the last pusher to finish publishes the epoch -- once every reader of THIS launch is
            // done with the mailbox, so that a neighbour that sees the epoch may refill the slot
            // this launch read from (it does so two batches on)
            __threadfence_system();
            __syncthreads();
            if (t == 0 && atomicAdd(Fz.count + 1, 1u) == (unsigned)Fz.n_pushers - 1u) {
                const unsigned long long t0 = global_ns();
                while (*reinterpret_cast<volatile unsigned int*>(Fz.count) < (unsigned)Fz.n_readers) {
                    if (global_ns() - t0 > Fz.timeout_ns) {
                        Fz.h_status[0] = 1;
                        break;
                    }
                    __nanosleep(100);
                }
                Fz.count[0] = 0u;
                Fz.count[1] = 0u;
                __threadfence_system();
                const u64 e_out = *Fz.epoch + (u64)Fz.batch + 1ULL;
                st_release_sys(Fz.prev + 1, e_out); // prev's flag "from next"
                st_release_sys(Fz.next + 0, e_out); // next's flag "from prev"
            }
        }
        return;
    }

    // ---- this tile's per-step sums (warps added in order) -> global log, tile-major
    //      [r][tile][step][FQSB_NLOG]. The totals per step are formed in two levels, both in a
    //      fixed order (deterministic: a redone batch decides exactly as the revoked one): the
    //      last tile of a GROUP of FQSB_BK_GROUP consecutive tiles to finish adds the group up,
    //      the last group to finish adds the groups up. (One level -- the last tile reading
    //      ntiles x nsteps entries alone -- was a serial tail of ~150 us per launch on a line
    //      of 2^20 blocks: 1171 tiles x 64 steps x 40 B through one SM.)
    constexpr int G = FQSB_BK_GROUP, LOGSZ = FQSB_BK_MAXSTEPS * FQSB_NLOG;
    const int ngroups = (K.ntiles + G - 1) / G;
    const int grp = c / G;
    const int gsize = K.ntiles - grp * G < G ? K.ntiles - grp * G : G;
    __syncthreads();
    {
        double* e = K.log + ((size_t)r * K.ntiles + c) * LOGSZ;
        for (int s = t; s < nsteps; s += T) {
            __stcg(e + s * FQSB_NLOG, slog[s * 2]);
            __stcg(e + s * FQSB_NLOG + 1, slog[s * 2 + 1]);
            __stcg(e + s * FQSB_NLOG + 2, (double)ilog[s * 4]);
            __stcg(e + s * FQSB_NLOG + 3, (double)ilog[s * 4 + 1]);
            __stcg(e + s * FQSB_NLOG + 4, (double)ilog[s * 4 + 2]);
        }
    }
    unsigned int* gcount = K.gcount + (size_t)r * ngroups + grp;
    __threadfence();
    __syncthreads();
    if (t == 0) {
        unsigned int ticket = atomicAdd(gcount, 1u);
        s_last = ticket == (unsigned)gsize - 1u;
    }
    __syncthreads();
    if (!s_last) {
        return;
    }
    // ---- the last tile of its group: group totals (tiles added in order)
    __threadfence();
    {
        const double* src = K.log + ((size_t)r * K.ntiles + (size_t)grp * G) * LOGSZ;
        double* dst = K.glog + ((size_t)r * ngroups + grp) * LOGSZ;
        for (int i = t; i < nsteps * FQSB_NLOG; i += T) {
            double x = 0.0;
#pragma unroll 8
            for (int cc = 0; cc < gsize; ++cc) {
                x += __ldcg(src + (size_t)cc * LOGSZ + i);
            }
            __stcg(dst + i, x);
        }
        if (t == 0) {
            *gcount = 0u;
        }
    }
    __threadfence();
    __syncthreads();
    if (t == 0) {
        unsigned int ticket = atomicAdd(&ctl.count, 1u);
        s_last = ticket == (unsigned)ngroups - 1u;
    }
    __syncthreads();
    if (!s_last) {
        return;
    }
    // ---- the last group of the realisation: totals per step (groups added in order), then the
    //      sequential replay of the decisions
    __threadfence();
    double* tot = reinterpret_cast<double*>(park); // [MAXSTEPS][FQSB_NLOG] (the parked sums have been consumed)
    {
        const double* src = K.glog + (size_t)r * ngroups * LOGSZ;
        for (int i = t; i < nsteps * FQSB_NLOG; i += T) {
            double x = 0.0;
#pragma unroll 8
            for (int gg = 0; gg < ngroups; ++gg) {
                x += __ldcg(src + (size_t)gg * LOGSZ + i);
            }
            tot[i] = x;
        }
    }
    __syncthreads();
    if (A.mode == MODE_LOG) {
        // member of a slab-decomposed line (fqsb_slab.inl): hand the member's per-step sums over;
        // the decision is taken on the sums of ALL members (k_slab_import), which then commits
        // the batch or asks for a shorter one
        for (int i = t; i < nsteps * FQSB_NLOG; i += T) {
            A.log[i] = tot[i];
        }
        if (t == 0) {
            ctl.count = 0u;
        }
        return;
    }
    if (warp != 0) {
        return;
    }
    Prog g;
    prog_load(g, ctl);
    RingEntry ring = ring_load(ctl, A, lane);
    int status = ST_RUNNING;
    double sf = 0.0, sff = 0.0;
    int s = 0;
    for (; s < nsteps; ++s) {
        sf = tot[s * FQSB_NLOG];
        sff = tot[s * FQSB_NLOG + 1];
        g.inc++; // detail.h:1541
        status = step_decide(A, g, ring, lane, sf, sff, (int)tot[s * FQSB_NLOG + 2],
                             (int)tot[s * FQSB_NLOG + 3], (int)tot[s * FQSB_NLOG + 4]);
        if (status != ST_RUNNING) {
            break;
        }
    }
    if (status == ST_RUNNING || s == nsteps - 1) {
        // commit: the output set becomes current
        ring_store(ctl, A, lane, ring);
        if (lane == 0) {
            prog_store(g, ctl);
            ctl.residual = residual_from_sums(sf, sff);
            ctl.flip = flip ^ 1;
            const i64 left = A.max_steps - g.steps;
            ctl.batch = (int)(left < K.ksteps ? left : K.ksteps);
            ctl.count = 0u;
            ctl.status = status;
        }
    }
    else if (lane == 0) {
        // the criterion fired inside the batch: keep the input set, redo exactly s + 1 steps
        // (the replay of that shorter batch stops at its last step and commits)
        ctl.batch = s + 1;
        ctl.count = 0u;
    }
}

#ifdef FQSB_BLOCKED_HELPERS
// start of a stop-mode call: size of the first batch
__global__ void k_blocked_begin(const Par P, const State S, int ksteps, i64 max_steps)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < P.R) {
        S.ctl[r].batch = (int)(max_steps < ksteps ? max_steps : ksteps);
    }
}

// end of a fixed-step call (no per-step bookkeeping on the device)
__global__ void k_blocked_fixed_done(const Par P, const State S, i64 nsteps, int flip)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < P.R) {
        Ctl& c = S.ctl[r];
        c.inc += nsteps; // detail.h:1541
        c.steps = nsteps;
        c.flip = flip;
        c.status = ST_EXHAUSTED;
    }
}

// bring the current set back into the primary arrays; quench() on convergence
// (detail.h:1527-1532, 1781)
__global__ void k_blocked_settle(const Par P, const State S, const BlockedArgs K)
{
    const int r = blockIdx.y;
    Ctl& ctl = S.ctl[r];
    const int flip = ctl.flip;
    const bool quench = ctl.status == ST_CONVERGED;
    if (!flip && !quench) {
        return;
    }
    const i64 base = (i64)r * P.N;
    for (i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x; p < P.N;
         p += (i64)gridDim.x * blockDim.x) {
        const i64 g = base + p;
        if (flip) {
            S.u[g] = S.u2[g];
            S.yl[g] = K.yl2[g];
            S.yr[g] = K.yr2[g];
            S.idx[g] = K.idx2[g];
            S.rng[g] = K.rng2[g];
        }
        if (quench) {
            S.v[g] = 0.0;
            S.a[g] = 0.0;
        }
        else if (flip) {
            S.v[g] = S.v2[g];
            S.a[g] = S.a2[g];
        }
    }
    if (flip && blockIdx.x == 0 && threadIdx.x == 0) {
        S.u_frame[r] = K.uf2[r];
    }
}

#endif // FQSB_BLOCKED_HELPERS

} // namespace fqsb
