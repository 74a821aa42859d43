// fqsb_slowdist.cu -- the generic streaming kernels compiled WITH the `gamma` / `normal` yield
// distributions (FQSB_SLOW_DISTS, see fqsb_device.cuh). Systems with one of these distributions
// take these kernels for every stepping call (fqsb_stream.cu routes them here); all other
// translation units stay free of the special-function code. Everything is compiled into its own
// namespace so that the template kernels do not collide with their fast twins (same names,
// different bodies).
#define FQSB_SLOW_DISTS
#define fqsb fqsb_slowdist
#include "fqsb_kernels.cuh"
#undef fqsb

namespace S = fqsb_slowdist;

// Par / State / RunArgs are the same plain structs in both namespaces (same header)
cudaError_t launch_stream_step_slowdist(const void* Pv, const void* Sv, const void* Av,
                                        cudaStream_t stream, int flip, int finalise)
{
    const S::Par& P = *static_cast<const S::Par*>(Pv);
    const S::State& St = *static_cast<const S::State*>(Sv);
    const S::RunArgs& A = *static_cast<const S::RunArgs*>(Av);
    dim3 grid((unsigned)St.tiles, (unsigned)P.R);
#define FQSB_SLOW(pot, inter) \
    if (P.thermal) \
        S::k_stream_step<pot, inter, true><<<grid, 256, 0, stream>>>(P, St, A, flip, finalise); \
    else \
        S::k_stream_step<pot, inter, false><<<grid, 256, 0, stream>>>(P, St, A, flip, finalise); \
    return cudaGetLastError();
    if (P.pot == S::POT_CUSPY) {
        switch (P.inter) {
        case S::INT_NONE: FQSB_SLOW(S::POT_CUSPY, S::INT_NONE)
        case S::INT_LAPLACE1D: FQSB_SLOW(S::POT_CUSPY, S::INT_LAPLACE1D)
        case S::INT_QUARTIC1D: FQSB_SLOW(S::POT_CUSPY, S::INT_QUARTIC1D)
        case S::INT_QUARTICGRADIENT1D: FQSB_SLOW(S::POT_CUSPY, S::INT_QUARTICGRADIENT1D)
        case S::INT_LONGRANGE1D: FQSB_SLOW(S::POT_CUSPY, S::INT_LONGRANGE1D)
        case S::INT_LAPLACE2D: FQSB_SLOW(S::POT_CUSPY, S::INT_LAPLACE2D)
        case S::INT_QUARTICGRADIENT2D: FQSB_SLOW(S::POT_CUSPY, S::INT_QUARTICGRADIENT2D)
        }
    }
    if (P.pot == S::POT_SEMISMOOTH && P.inter == S::INT_LAPLACE1D) {
        FQSB_SLOW(S::POT_SEMISMOOTH, S::INT_LAPLACE1D)
    }
    if (P.pot == S::POT_SMOOTH && P.inter == S::INT_LAPLACE1D) {
        FQSB_SLOW(S::POT_SMOOTH, S::INT_LAPLACE1D)
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_stream_sweep_slowdist(const void* Pv, const void* Sv, const void* Av,
                                         cudaStream_t stream, int flip, int first, int sweep_arg)
{
    const S::Par& P = *static_cast<const S::Par*>(Pv);
    const S::State& St = *static_cast<const S::State*>(Sv);
    const S::RunArgs& A = *static_cast<const S::RunArgs*>(Av);
    dim3 grid((unsigned)St.tiles, (unsigned)P.R);
    if (P.inter == S::INT_LAPLACE2D) {
        S::k_stream_np<S::INT_LAPLACE2D><<<grid, 256, 0, stream>>>(P, St, A, flip, first, sweep_arg);
    }
    else {
        S::k_stream_np<S::INT_LAPLACE1D><<<grid, 256, 0, stream>>>(P, St, A, flip, first, sweep_arg);
    }
    return cudaGetLastError();
}
