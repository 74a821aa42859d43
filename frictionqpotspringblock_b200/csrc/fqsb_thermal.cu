// fqsb_thermal.cu -- instantiation unit of the resident kernel of the thermal systems (K2t).
#include "fqsb_host.h"
#include "fqsb_thermal.cuh"

namespace fqsb {

size_t resident_thermal_smem(const Par& P, const ResidentCfg& c)
{
    // us[2][N+2], fth[N], fnew[2][N], jump tables, nrel[N], dincs[N], bw[2][K], dl[N] (u16)
    const size_t n = (size_t)P.N, k = (size_t)c.B * (size_t)(c.T / 32);
    return (2 * (n + 2) + 3 * n + 2 * FQSB_TH_JUMPS) * 8 + (2 * n + 2 * k) * sizeof(int) + n * 2;
}

template <int INT, int B, int T, bool FULL, bool UNIT>
static cudaError_t launch_k(size_t smem, const Par& P, const State& S, const RunArgs& A,
                            const Thermal& TH, cudaStream_t stream)
{
    auto kernel = k_resident_thermal<INT, B, T, FULL, UNIT>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
        return e;
    }
    kernel<<<(unsigned)P.R, T + 32, smem, stream>>>(P, S, A, TH); // + the producer warp
    return cudaGetLastError();
}

template <int INT, int B, int T>
static cudaError_t launch_one(size_t smem, const Par& P, const State& S, const RunArgs& A,
                              const Thermal& TH, cudaStream_t stream)
{
    const bool full = (i64)B * T == P.N;
    const bool unit = unit_parameters(P);
    if (full) {
        return unit ? launch_k<INT, B, T, true, true>(smem, P, S, A, TH, stream)
                    : launch_k<INT, B, T, true, false>(smem, P, S, A, TH, stream);
    }
    return unit ? launch_k<INT, B, T, false, true>(smem, P, S, A, TH, stream)
                : launch_k<INT, B, T, false, false>(smem, P, S, A, TH, stream);
}

template <int INT>
static cudaError_t launch_cfg(const ResidentCfg& c, const Par& P, const State& S,
                              const RunArgs& A, const Thermal& TH, cudaStream_t stream)
{
    const size_t smem = resident_thermal_smem(P, c);
    if (c.B == 1 && c.T == 256) {
        return launch_one<INT, 1, 256>(smem, P, S, A, TH, stream);
    }
    if (c.B == 2 && c.T == 512) {
        return launch_one<INT, 2, 512>(smem, P, S, A, TH, stream);
    }
    if (c.B == 4 && c.T == 512) {
        return launch_one<INT, 4, 512>(smem, P, S, A, TH, stream);
    }
    if (c.B == 8 && c.T == 512) {
        return launch_one<INT, 8, 512>(smem, P, S, A, TH, stream);
    }
    return cudaErrorInvalidConfiguration;
}

// Cuspy x {Laplace1d, Quartic1d, no interactions}: Line1d.h:261-330, 486-556, Particles.h
cudaError_t launch_resident_thermal(const ResidentCfg& c, const Par& P, const State& S,
                                    const RunArgs& A, const Thermal& TH, cudaStream_t stream)
{
    if (P.pot != POT_CUSPY) {
        return cudaErrorInvalidValue;
    }
    switch (P.inter) {
    case INT_LAPLACE1D: return launch_cfg<INT_LAPLACE1D>(c, P, S, A, TH, stream);
    case INT_QUARTIC1D: return launch_cfg<INT_QUARTIC1D>(c, P, S, A, TH, stream);
    case INT_NONE: return launch_cfg<INT_NONE>(c, P, S, A, TH, stream);
    }
    return cudaErrorInvalidValue;
}

} // namespace fqsb
