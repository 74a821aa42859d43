// fqsb_resident.cu -- instantiation unit of the resident (one CTA = one realisation) kernels.
// Compiled once per potential x interaction combination (-DFQSB_COMBO=k) so the objects build
// in parallel; combination 9 holds the no-passing kernels.
#include "fqsb_host.h"
#include "fqsb_kernels.cuh"

#ifndef FQSB_COMBO
#error "compile with -DFQSB_COMBO=<0..9>"
#endif

#define FQSB_CAT2(a, b) a##b
#define FQSB_CAT(a, b) FQSB_CAT2(a, b)

// FQSB_FMA_BUILD: the same kernels compiled with -fmad=true (the opt-in "contracted arithmetic" of
// fqsb_params.kernel bit 7): distinct instantiations (template flag FMA) behind launch_resident_fma_<k>
#ifdef FQSB_FMA_BUILD
#define C_FMA true
#define FQSB_LAUNCH_NAME(k) FQSB_CAT(launch_resident_fma_, k)
#else
#define C_FMA false
#define FQSB_LAUNCH_NAME(k) FQSB_CAT(launch_resident_, k)
#endif

namespace fqsb {

#if FQSB_COMBO == 0
#define C_POT POT_CUSPY
#define C_INT INT_LAPLACE1D
#elif FQSB_COMBO == 1
#define C_POT POT_CUSPY
#define C_INT INT_QUARTIC1D
#elif FQSB_COMBO == 2
#define C_POT POT_CUSPY
#define C_INT INT_QUARTICGRADIENT1D
#elif FQSB_COMBO == 3
#define C_POT POT_CUSPY
#define C_INT INT_LONGRANGE1D
#elif FQSB_COMBO == 4
#define C_POT POT_CUSPY
#define C_INT INT_LAPLACE2D
#elif FQSB_COMBO == 5
#define C_POT POT_CUSPY
#define C_INT INT_QUARTICGRADIENT2D
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

template <class K>
static cudaError_t launch(K kernel, int T, size_t smem, const Par& P, const State& S,
                          const RunArgs& A, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
        return e;
    }
    kernel<<<(unsigned)P.R, T, smem, stream>>>(P, S, A);
    return cudaGetLastError();
}

#if FQSB_COMBO < 9

cudaError_t FQSB_LAUNCH_NAME(FQSB_COMBO)(const ResidentCfg& c, const Par& P, const State& S,
                                         const RunArgs& A, cudaStream_t stream)
{
    const bool stop = A.mode != MODE_FIXED;
    const size_t smem = resident_smem(P, c, stop);
    const bool full = (i64)c.B * c.T == P.N;
    const bool unit = unit_parameters(P);
    const bool flow = A.flow != 0; // driven: blocks change wells all the time (inline hop path)
#define FQSB_TRY_MODE(b, t, full_, unit_) \
    if (stop) \
        return launch(k_resident<C_POT, C_INT, b, t, false, full_, unit_, true, true, C_FMA>, t, smem, P, S, A, stream); \
    if (flow) \
        return launch(k_resident<C_POT, C_INT, b, t, false, full_, unit_, false, true, C_FMA>, t, smem, P, S, A, stream); \
    return launch(k_resident<C_POT, C_INT, b, t, false, full_, unit_, false, false, C_FMA>, t, smem, P, S, A, stream);
#define FQSB_TRY_CFG(b, t) \
    if (c.B == b && c.T == t) { \
        if (full && unit) { \
            FQSB_TRY_MODE(b, t, true, true) \
        } \
        if (full) { \
            FQSB_TRY_MODE(b, t, true, false) \
        } \
        if (unit) { \
            FQSB_TRY_MODE(b, t, false, true) \
        } \
        FQSB_TRY_MODE(b, t, false, false) \
    }
    FQSB_TRY_CFG(1, 256)
    FQSB_TRY_CFG(2, 512)
    FQSB_TRY_CFG(4, 512)
    FQSB_TRY_CFG(8, 512)
    return cudaErrorInvalidConfiguration;
}

#else // no-passing

template <int INT>
static cudaError_t launch_np(const ResidentCfg& c, const Par& P, const State& S,
                             const RunArgs& A, cudaStream_t stream)
{
    const size_t smem = resident_np_smem(P, c);
    if (c.B == 1 && c.T == 256) {
        return launch(k_resident_nopassing<INT, 1, 256>, 256, smem, P, S, A, stream);
    }
    if (c.B == 2 && c.T == 512) {
        return launch(k_resident_nopassing<INT, 2, 512>, 512, smem, P, S, A, stream);
    }
    if (c.B == 4 && c.T == 512) {
        return launch(k_resident_nopassing<INT, 4, 512>, 512, smem, P, S, A, stream);
    }
    if (c.B == 8 && c.T == 512) {
        return launch(k_resident_nopassing<INT, 8, 512>, 512, smem, P, S, A, stream);
    }
    return cudaErrorInvalidConfiguration;
}

cudaError_t launch_resident_nopassing(const ResidentCfg& c, const Par& P, const State& S,
                                      const RunArgs& A, cudaStream_t stream)
{
    if (P.inter == INT_LAPLACE2D) {
        return launch_np<INT_LAPLACE2D>(c, P, S, A, stream);
    }
    return launch_np<INT_LAPLACE1D>(c, P, S, A, stream);
}

#endif

} // namespace fqsb
