// fqsb_blocked.cu -- instantiation unit of the temporally blocked kernel (fqsb_blocked.cuh).
// Compiled once per potential x interaction combination (-DFQSB_COMBO=k, the numbering of
// fqsb_resident.cu; only the 1-D nearest-neighbour combinations 0, 1, 2, 6, 7, 8 exist) so the
// objects build in parallel; -DFQSB_COMBO=100 builds the dispatcher and the tile planner.
#include "fqsb_host.h"
#ifndef FQSB_COMBO
#error "compile with -DFQSB_COMBO=<0|1|2|6|7|8|100>"
#endif
#if FQSB_COMBO == 100
#define FQSB_BLOCKED_HELPERS
#endif
#include "fqsb_blocked.cuh"


namespace fqsb {

#if FQSB_COMBO != 100

#if FQSB_COMBO == 0
#define C_POT POT_CUSPY
#define C_INT INT_LAPLACE1D
#elif FQSB_COMBO == 1
#define C_POT POT_CUSPY
#define C_INT INT_QUARTIC1D
#elif FQSB_COMBO == 2
#define C_POT POT_CUSPY
#define C_INT INT_QUARTICGRADIENT1D
#elif FQSB_COMBO == 6
#define C_POT POT_SEMISMOOTH
#define C_INT INT_LAPLACE1D
#elif FQSB_COMBO == 7
#define C_POT POT_SMOOTH
#define C_INT INT_LAPLACE1D
#elif FQSB_COMBO == 8
#define C_POT POT_CUSPY
#define C_INT INT_NONE
#endif

#define FQSB_CAT2(a, b) a##b
#define FQSB_CAT(a, b) FQSB_CAT2(a, b)

// FQSB_FMA_BUILD: the same kernels compiled with -fmad=true (opt-in contracted arithmetic,
// fqsb_params.kernel bit 7; Cuspy combinations 0, 1, 2) behind launch_blocked_fma_<k>
#ifdef FQSB_FMA_BUILD
#define C_FMA true
#define FQSB_BK_LAUNCH_NAME(k) FQSB_CAT(launch_blocked_fma_, k)
#else
#define C_FMA false
#define FQSB_BK_LAUNCH_NAME(k) FQSB_CAT(launch_blocked_, k)
#endif

template <class Kern>
static cudaError_t launch(Kern kernel, const BlockedPlan& plan, const Par& P, const State& S,
                          const RunArgs& A, const BlockedArgs& K, cudaStream_t stream)
{
    const size_t smem = blocked_smem(plan.B, A.mode != MODE_FIXED);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
        return e;
    }
    dim3 grid((unsigned)plan.ntiles, (unsigned)P.R);
    kernel<<<grid, FQSB_BK_T, smem, stream>>>(P, S, A, K);
    return cudaGetLastError();
}

cudaError_t FQSB_BK_LAUNCH_NAME(FQSB_COMBO)(const BlockedPlan& plan, const Par& P,
                                            const State& S, const RunArgs& A,
                                            const BlockedArgs& K, cudaStream_t stream)
{
    const bool unit = unit_parameters(P);
    const bool stop = A.mode != MODE_FIXED;
#define FQSB_BK_CFG(b) \
    if (plan.B == b) { \
        if (!stop && K.fuse.on) { /* a slab member's fixed-step batch with the exchange on board */ \
            return unit ? launch(k_blocked<C_POT, C_INT, b, true, false, C_FMA, true>, plan, P, S, A, K, stream) \
                        : launch(k_blocked<C_POT, C_INT, b, false, false, C_FMA, true>, plan, P, S, A, K, stream); \
        } \
        if (unit) { \
            return stop ? launch(k_blocked<C_POT, C_INT, b, true, true, C_FMA>, plan, P, S, A, K, stream) \
                        : launch(k_blocked<C_POT, C_INT, b, true, false, C_FMA>, plan, P, S, A, K, stream); \
        } \
        return stop ? launch(k_blocked<C_POT, C_INT, b, false, true, C_FMA>, plan, P, S, A, K, stream) \
                    : launch(k_blocked<C_POT, C_INT, b, false, false, C_FMA>, plan, P, S, A, K, stream); \
    }
    FQSB_BK_CFG(2)
    FQSB_BK_CFG(3)
    FQSB_BK_CFG(4)
    FQSB_BK_CFG(5)
    FQSB_BK_CFG(6)
    FQSB_BK_CFG(7)
    FQSB_BK_CFG(8)
    return cudaErrorInvalidConfiguration;
}

#else // dispatcher + planner

#define FQSB_DECL(k) \
    cudaError_t launch_blocked_##k(const BlockedPlan&, const Par&, const State&, const RunArgs&, \
                                   const BlockedArgs&, cudaStream_t);
FQSB_DECL(0) FQSB_DECL(1) FQSB_DECL(2) FQSB_DECL(6) FQSB_DECL(7) FQSB_DECL(8)
FQSB_DECL(fma_0) FQSB_DECL(fma_1) FQSB_DECL(fma_2)

static int blocked_combo(const Par& P)
{
    if (P.rank != 1) {
        return -1;
    }
    if (P.pot == POT_CUSPY) {
        switch (P.inter) {
        case INT_LAPLACE1D: return 0;
        case INT_QUARTIC1D: return 1;
        case INT_QUARTICGRADIENT1D: return 2;
        case INT_NONE: return 8;
        }
    }
    if (P.pot == POT_SEMISMOOTH && P.inter == INT_LAPLACE1D) {
        return 6;
    }
    if (P.pot == POT_SMOOTH && P.inter == INT_LAPLACE1D) {
        return 7;
    }
    return -1;
}

bool blocked_supported(const Par& P) { return blocked_combo(P) >= 0 && P.N >= 2; }

// Tile geometry. A CTA of FQSB_BK_T threads x B blocks per thread holds own + 2 H <= T B local
// blocks and costs ~B time units per step whatever its fill. The tiles resident on one SM share
// its FP64 issue slots (two tiles per SM only overlap each other's load / store phases and
// latencies), so the grid of ntiles x R CTAs costs ceil(ntiles R / SMs) x B per step: pick the
// (B, ntiles) with the cheapest, ties broken towards fewer tiles (less halo). [Counting
// FQSB_BK_CTAS x SMs slots instead gave the members of an 8-GPU slab 205 tiles of B = 3 -- two
// tiles on 57 SMs, one on the others -- where 147 tiles of B = 4 do one tile per SM.]
// `own_hint` > 0 fixes the tile size (tests), `ksteps_hint` > 0 the steps per launch.
BlockedPlan blocked_plan(const Par& P, int ksteps_hint, int own_hint)
{
    BlockedPlan best;
    memset(&best, 0, sizeof best);
    const i64 N = P.N;
    int ksteps = ksteps_hint > 0 ? ksteps_hint : FQSB_BK_MAXSTEPS;
    if (ksteps > FQSB_BK_MAXSTEPS) {
        ksteps = FQSB_BK_MAXSTEPS;
    }
    int H = P.inter == INT_NONE ? 0 : ksteps;
    if (H > (int)((N - 1) / 2)) { // a halo never wraps onto its own tile's far side twice
        H = (int)((N - 1) / 2);
        if (H < 1) {
            H = 1;
        }
        ksteps = H;
    }
    int sms = 148;
    {
        // (cudaDeviceGetAttribute, not cudaGetDeviceProperties: the latter costs milliseconds per
        // call and the planner runs once per dynamics call)
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
            sms = n;
        }
    }
    if (own_hint > 0) {
        i64 own = own_hint < N ? own_hint : N;
        int B = (int)((own + 2 * H + FQSB_BK_T - 1) / FQSB_BK_T);
        if (B < 2) {
            B = 2;
        }
        if (B <= 8) {
            best.B = B;
            best.own = (int)own;
            best.H = H;
            best.ksteps = ksteps;
            best.ntiles = (int)((N + own - 1) / own);
            return best;
        }
    }
    i64 best_cost = -1;
    for (int B = 2; B <= 8; ++B) {
        const i64 cap = (i64)B * FQSB_BK_T - 2 * H;
        if (cap < 1) {
            continue;
        }
        const i64 nt_min = (N + cap - 1) / cap;
        for (i64 nt = nt_min; nt < nt_min + sms; ++nt) {
            const i64 waves = (nt * P.R + sms - 1) / sms;
            const i64 cost = waves * B;
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best.B = B;
                best.ntiles = (int)nt;
            }
            if (P.R >= sms) {
                break; // many realisations: the waves are smooth, keep the fewest tiles
            }
        }
    }
    best.own = (int)((N + best.ntiles - 1) / best.ntiles);
    best.ntiles = (int)((N + best.own - 1) / best.own);
    best.H = H;
    best.ksteps = ksteps;
    return best;
}

cudaError_t launch_blocked(const BlockedPlan& plan, const Par& P, const State& S,
                           const RunArgs& A, const BlockedArgs& K, cudaStream_t stream)
{
    const int combo = blocked_combo(P);
    if (P.fma && combo >= 0 && combo <= 2) { // opt-in contracted arithmetic (Cuspy lines)
        switch (combo) {
        case 0: return launch_blocked_fma_0(plan, P, S, A, K, stream);
        case 1: return launch_blocked_fma_1(plan, P, S, A, K, stream);
        case 2: return launch_blocked_fma_2(plan, P, S, A, K, stream);
        }
    }
    switch (combo) {
    case 0: return launch_blocked_0(plan, P, S, A, K, stream);
    case 1: return launch_blocked_1(plan, P, S, A, K, stream);
    case 2: return launch_blocked_2(plan, P, S, A, K, stream);
    case 6: return launch_blocked_6(plan, P, S, A, K, stream);
    case 7: return launch_blocked_7(plan, P, S, A, K, stream);
    case 8: return launch_blocked_8(plan, P, S, A, K, stream);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_blocked_begin(const Par& P, const State& S, int ksteps, i64 max_steps,
                                 cudaStream_t stream)
{
    k_blocked_begin<<<(unsigned)((P.R + 127) / 128), 128, 0, stream>>>(P, S, ksteps, max_steps);
    return cudaGetLastError();
}

cudaError_t launch_blocked_fixed_done(const Par& P, const State& S, i64 nsteps, int flip,
                                      cudaStream_t stream)
{
    k_blocked_fixed_done<<<(unsigned)((P.R + 127) / 128), 128, 0, stream>>>(P, S, nsteps, flip);
    return cudaGetLastError();
}

cudaError_t launch_blocked_settle(const Par& P, const State& S, const BlockedArgs& K,
                                  cudaStream_t stream)
{
    i64 tiles = (P.N + 2047) / 2048;
    if (tiles > 1024) {
        tiles = 1024;
    }
    dim3 grid((unsigned)tiles, (unsigned)P.R);
    k_blocked_settle<<<grid, 256, 0, stream>>>(P, S, K);
    return cudaGetLastError();
}

#endif

} // namespace fqsb
