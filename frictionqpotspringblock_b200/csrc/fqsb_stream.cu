// fqsb_stream.cu -- instantiation unit of the streaming kernels + kernel dispatch tables.
#include <cmath>
#include <cstdlib>

#include "fqsb_host.h"
#include "fqsb_kernels.cuh"

// fqsb_slowdist.cu: the generic kernels compiled with the `gamma` / `normal` distributions
cudaError_t launch_stream_step_slowdist(const void* P, const void* S, const void* A,
                                        cudaStream_t stream, int flip, int finalise);
cudaError_t launch_stream_sweep_slowdist(const void* P, const void* S, const void* A,
                                         cudaStream_t stream, int flip, int first, int sweep_arg);

namespace fqsb {

static bool slow_dist(const Par& P) { return P.dist == DIST_GAMMA || P.dist == DIST_NORMAL; }

#define FQSB_DECL(k) \
    cudaError_t launch_resident_##k(const ResidentCfg&, const Par&, const State&, \
                                    const RunArgs&, cudaStream_t);
FQSB_DECL(0) FQSB_DECL(1) FQSB_DECL(2) FQSB_DECL(3) FQSB_DECL(4)
FQSB_DECL(5) FQSB_DECL(6) FQSB_DECL(7) FQSB_DECL(8)
FQSB_DECL(fma_0) FQSB_DECL(fma_1) FQSB_DECL(fma_2)

// the systems of Line1d.h:112-677 and Line2d.h:77-162 (+ interaction-free, Particles.h)
static int combo_of(int pot, int inter)
{
    if (pot == POT_CUSPY) {
        switch (inter) {
        case INT_LAPLACE1D: return 0;
        case INT_QUARTIC1D: return 1;
        case INT_QUARTICGRADIENT1D: return 2;
        case INT_LONGRANGE1D: return 3;
        case INT_LAPLACE2D: return 4;
        case INT_QUARTICGRADIENT2D: return 5;
        case INT_NONE: return 8;
        }
    }
    if (pot == POT_SEMISMOOTH && inter == INT_LAPLACE1D) {
        return 6;
    }
    if (pot == POT_SMOOTH && inter == INT_LAPLACE1D) {
        return 7;
    }
    return -1;
}

bool combination_supported(int pot, int inter) { return combo_of(pot, inter) >= 0; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: set it once on every
// device a kernel is launched on (a process may drive several GPUs, e.g. the members of a slab)
template <class K>
static cudaError_t ensure_dynamic_smem(K kernel, size_t smem, bool (&done)[64])
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        return e;
    }
    if (dev >= 0 && dev < 64 && done[dev]) {
        return cudaSuccess;
    }
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && dev >= 0 && dev < 64) {
        done[dev] = true;
    }
    return e;
}

cudaError_t launch_resident(const ResidentCfg& c, const Par& P, const State& S,
                            const RunArgs& A, cudaStream_t stream)
{
    if (P.fma) { // opt-in contracted arithmetic: Cuspy lines with nearest-neighbour stencils
        switch (combo_of(P.pot, P.inter)) {
        case 0: return launch_resident_fma_0(c, P, S, A, stream);
        case 1: return launch_resident_fma_1(c, P, S, A, stream);
        case 2: return launch_resident_fma_2(c, P, S, A, stream);
        }
    }
    switch (combo_of(P.pot, P.inter)) {
    case 0: return launch_resident_0(c, P, S, A, stream);
    case 1: return launch_resident_1(c, P, S, A, stream);
    case 2: return launch_resident_2(c, P, S, A, stream);
    case 3: return launch_resident_3(c, P, S, A, stream);
    case 4: return launch_resident_4(c, P, S, A, stream);
    case 5: return launch_resident_5(c, P, S, A, stream);
    case 6: return launch_resident_6(c, P, S, A, stream);
    case 7: return launch_resident_7(c, P, S, A, stream);
    case 8: return launch_resident_8(c, P, S, A, stream);
    }
    return cudaErrorInvalidValue;
}

// tiles per realisation of the step kernel that launch_stream_step() picks
// (thermal systems -- External = RandomNormalForcing -- take the generic kernel)
static bool use_tiled_1d(const Par& P)
{
    return !P.thermal && !slow_dist(P) && P.inter <= INT_QUARTICGRADIENT1D && P.N % 2 == 0;
}

static bool use_tiled_2d(const Par& P)
{
    return !P.thermal && !slow_dist(P) && P.rank == 2 && P.cols % 2 == 0;
}

int stream_step_tiles(const Par& P, int generic_tiles)
{
    if (use_tiled_1d(P)) {
        return (int)((P.N + FQSB_ST_TILE - 1) / FQSB_ST_TILE);
    }
    if (use_tiled_2d(P)) {
        const int ty = P.s2_ty > 0 ? P.s2_ty : FQSB_S2_TY;
        return ((P.cols + FQSB_S2_TX - 1) / FQSB_S2_TX) * ((P.rows + ty - 1) / ty);
    }
    return generic_tiles;
}

// CTAs per realisation of the tiled 2-D no-passing sweep
static int stream_sweep_tiles(const Par& P)
{
    const int ty = P.s2_ty_np > 0 ? P.s2_ty_np : FQSB_S2_TY;
    return ((P.cols + FQSB_S2_TX - 1) / FQSB_S2_TX) * ((P.rows + ty - 1) / ty);
}

// CTAs per launch (x dimension) of the kernel launch_stream_step / launch_stream_sweep pick
int stream_launch_tiles(const Par& P, int generic_tiles, bool sweep)
{
    if (sweep) {
        return use_tiled_2d(P) ? stream_sweep_tiles(P) : generic_tiles;
    }
    return stream_step_tiles(P, generic_tiles);
}

// Rows per CTA of the row-marching 2-D kernels: the grid (strips x bands x realisations) should
// fill whole waves of the `resident` CTAs the device holds at once, at the price of 2 halo rows
// per band (`halo` = their relative cost per row of a band). At most `max_tiles` CTAs per
// realisation (the size of the per-CTA partial-sum buffer).
int plan_band_rows(const Par& P, int resident, double halo, int max_tiles)
{
    if (P.rank != 2) {
        return 0;
    }
    const int strips = (P.cols + FQSB_S2_TX - 1) / FQSB_S2_TX;
    // fewest rows per band that keep strips x bands within max_tiles (very large interfaces)
    const int max_bands = max_tiles / strips > 0 ? max_tiles / strips : 1;
    const int ty_min = (P.rows + max_bands - 1) / max_bands;
    const int ty_lo = ty_min > 8 ? ty_min : 8;
    const int ty_hi = 2 * ty_lo > 128 ? 2 * ty_lo : 128;
    int best = ty_lo;
    double best_cost = 1e300;
    for (int ty = ty_lo; ty <= ty_hi; ++ty) {
        const int bands = (P.rows + ty - 1) / ty;
        if (strips * bands > max_tiles) {
            continue;
        }
        const double n = (double)strips * bands * (double)P.R;
        const double waves = n / resident;
        const double cost = std::ceil(waves) / waves * (1.0 + halo / ty);
        if (cost < best_cost - 1e-12) {
            best_cost = cost;
            best = ty;
        }
    }
    return best;
}

const char* stream_step_name(const Par& P)
{
    return use_tiled_1d(P) ? "stream_1d" : (use_tiled_2d(P) ? "stream_2d" : "stream");
}

cudaError_t launch_stream_step(const Par& P, const State& S, const RunArgs& A,
                               cudaStream_t stream, int flip, int finalise)
{
    if (use_tiled_1d(P)) {
        dim3 grid((unsigned)stream_step_tiles(P, 0), (unsigned)P.R);
        const bool unit = unit_parameters(P);
#define FQSB_TILED(pot, inter) \
    if (unit) \
        k_stream_1d<pot, inter, true><<<grid, FQSB_ST_THREADS, 0, stream>>>(P, S, A, flip, finalise); \
    else \
        k_stream_1d<pot, inter, false><<<grid, FQSB_ST_THREADS, 0, stream>>>(P, S, A, flip, finalise); \
    break;
        switch (combo_of(P.pot, P.inter)) {
        case 0: FQSB_TILED(POT_CUSPY, INT_LAPLACE1D)
        case 1: FQSB_TILED(POT_CUSPY, INT_QUARTIC1D)
        case 2: FQSB_TILED(POT_CUSPY, INT_QUARTICGRADIENT1D)
        case 6: FQSB_TILED(POT_SEMISMOOTH, INT_LAPLACE1D)
        case 7: FQSB_TILED(POT_SMOOTH, INT_LAPLACE1D)
        case 8: FQSB_TILED(POT_CUSPY, INT_NONE)
        default: return cudaErrorInvalidValue;
        }
        return cudaGetLastError();
    }
    if (use_tiled_2d(P)) {
        dim3 grid((unsigned)stream_step_tiles(P, 0), (unsigned)P.R);
        const bool unit = unit_parameters(P);
        static const bool bulk = [] { // 0: register-staged look-ahead instead of the TMA engine
            const char* e = std::getenv("FQSB_S2_BULK");
            return e ? std::atoi(e) != 0 : true;
        }();
        if (bulk) {
            constexpr int NS = 4;
            const size_t smem = stream_2d_bulk_smem(NS);
#define FQSB_2D_BULK(inter, unit_) \
    { \
        static bool done[64] = {}; \
        const cudaError_t attr = ensure_dynamic_smem(k_stream_2d_bulk<inter, unit_, NS>, smem, done); \
        if (attr != cudaSuccess) { \
            return attr; \
        } \
        k_stream_2d_bulk<inter, unit_, NS><<<grid, FQSB_S2_THREADS, smem, stream>>>(P, S, A, flip, \
                                                                                  finalise); \
        return cudaGetLastError(); \
    }
            if (P.inter == INT_LAPLACE2D) {
                if (unit)
                    FQSB_2D_BULK(INT_LAPLACE2D, true)
                else
                    FQSB_2D_BULK(INT_LAPLACE2D, false)
            }
            else {
                if (unit)
                    FQSB_2D_BULK(INT_QUARTICGRADIENT2D, true)
                else
                    FQSB_2D_BULK(INT_QUARTICGRADIENT2D, false)
            }
#undef FQSB_2D_BULK
        }
        if (P.inter == INT_LAPLACE2D) {
            if (unit)
                k_stream_2d<INT_LAPLACE2D, true><<<grid, FQSB_S2_THREADS, 0, stream>>>(P, S, A, flip, finalise);
            else
                k_stream_2d<INT_LAPLACE2D, false><<<grid, FQSB_S2_THREADS, 0, stream>>>(P, S, A, flip, finalise);
        }
        else {
            if (unit)
                k_stream_2d<INT_QUARTICGRADIENT2D, true><<<grid, FQSB_S2_THREADS, 0, stream>>>(P, S, A, flip, finalise);
            else
                k_stream_2d<INT_QUARTICGRADIENT2D, false><<<grid, FQSB_S2_THREADS, 0, stream>>>(P, S, A, flip, finalise);
        }
        return cudaGetLastError();
    }
    if (slow_dist(P)) {
        return launch_stream_step_slowdist(&P, &S, &A, stream, flip, finalise);
    }
    dim3 grid((unsigned)S.tiles, (unsigned)P.R);
    if (P.thermal) { // Line1d.h:261-330, 486-556; Particles.h System_Cuspy_RandomForcing
        switch (combo_of(P.pot, P.inter)) {
        case 0:
            k_stream_step<POT_CUSPY, INT_LAPLACE1D, true><<<grid, 256, 0, stream>>>(P, S, A, flip, finalise);
            break;
        case 1:
            k_stream_step<POT_CUSPY, INT_QUARTIC1D, true><<<grid, 256, 0, stream>>>(P, S, A, flip, finalise);
            break;
        case 8:
            k_stream_step<POT_CUSPY, INT_NONE, true><<<grid, 256, 0, stream>>>(P, S, A, flip, finalise);
            break;
        default: return cudaErrorInvalidValue;
        }
        return cudaGetLastError();
    }
#define FQSB_STREAM(pot, inter) \
    k_stream_step<pot, inter><<<grid, 256, 0, stream>>>(P, S, A, flip, finalise); \
    break;
    switch (combo_of(P.pot, P.inter)) {
    case 0: FQSB_STREAM(POT_CUSPY, INT_LAPLACE1D)
    case 1: FQSB_STREAM(POT_CUSPY, INT_QUARTIC1D)
    case 2: FQSB_STREAM(POT_CUSPY, INT_QUARTICGRADIENT1D)
    case 3: FQSB_STREAM(POT_CUSPY, INT_LONGRANGE1D)
    case 4: FQSB_STREAM(POT_CUSPY, INT_LAPLACE2D)
    case 5: FQSB_STREAM(POT_CUSPY, INT_QUARTICGRADIENT2D)
    case 6: FQSB_STREAM(POT_SEMISMOOTH, INT_LAPLACE1D)
    case 7: FQSB_STREAM(POT_SMOOTH, INT_LAPLACE1D)
    case 8: FQSB_STREAM(POT_CUSPY, INT_NONE)
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_stream_sweep(const Par& P, const State& S, const RunArgs& A,
                                cudaStream_t stream, int flip, int first, int do_sweep)
{
    if (use_tiled_2d(P)) {
        dim3 grid2((unsigned)stream_sweep_tiles(P), (unsigned)P.R);
        static const bool bulk = [] { // 0: register-staged look-ahead instead of the TMA engine
            const char* e = std::getenv("FQSB_S2_NP_BULK");
            return e ? std::atoi(e) != 0 : true;
        }();
        if (bulk) {
            constexpr int NS = FQSB_S2_BULK_STAGES;
            static bool done[64] = {};
            const cudaError_t attr = ensure_dynamic_smem(k_stream_np_2d_bulk<NS, 2>,
                                                         stream_np_2d_bulk_smem(NS), done);
            if (attr != cudaSuccess) {
                return attr;
            }
            k_stream_np_2d_bulk<NS, 2><<<grid2, FQSB_S2_THREADS, stream_np_2d_bulk_smem(NS), stream>>>(
                P, S, A, flip, first, do_sweep);
            return cudaGetLastError();
        }
        k_stream_np_2d<FQSB_S2_NP_CTAS><<<grid2, FQSB_S2_THREADS, 0, stream>>>(P, S, A, flip, first,
                                                                             do_sweep);
        return cudaGetLastError();
    }
    if (slow_dist(P)) {
        return launch_stream_sweep_slowdist(&P, &S, &A, stream, flip, first, do_sweep);
    }
    dim3 grid((unsigned)S.tiles, (unsigned)P.R);
    if (P.inter == INT_LAPLACE2D) {
        k_stream_np<INT_LAPLACE2D><<<grid, 256, 0, stream>>>(P, S, A, flip, first, do_sweep);
    }
    else {
        k_stream_np<INT_LAPLACE1D><<<grid, 256, 0, stream>>>(P, S, A, flip, first, do_sweep);
    }
    return cudaGetLastError();
}

} // namespace fqsb
