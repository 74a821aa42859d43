// fqsb_host.h -- host-side declarations shared by the C-ABI translation unit and the kernel
// instantiation units.
#pragma once

#include "fqsb_device.cuh"

namespace fqsb {

// blocks-per-thread x threads configurations of the resident kernels
struct ResidentCfg {
    int B, T;
};

// picks the smallest configuration that holds N blocks; B == 0 if N does not fit on chip
inline ResidentCfg resident_cfg(i64 N)
{
    if (N <= 256) {
        return {1, 256};
    }
    if (N <= 1024) {
        return {1, 1024};
    }
    if (N <= 2048) {
        return {2, 1024};
    }
    if (N <= 4096) {
        return {8, 512};
    }
    if (N <= 8192) {
        return {8, 1024};
    }
    return {0, 0};
}

inline size_t resident_smem(const Par& P, const ResidentCfg& c)
{
    size_t n = (size_t)P.N;
    size_t words = 3 * n + (P.inter == INT_LONGRANGE1D ? n : 0) + 2 * (size_t)(c.T / 32);
    return words * 8 + (size_t)(c.T / 32) * 4 * sizeof(int);
}

// defined in fqsb_resident.cu (one object per potential x interaction combination)
cudaError_t launch_resident(const ResidentCfg& cfg, const Par& P, const State& S,
                            const RunArgs& A, cudaStream_t stream);
cudaError_t launch_resident_nopassing(const ResidentCfg& cfg, const Par& P, const State& S,
                                      const RunArgs& A, cudaStream_t stream);
// defined in fqsb_stream.cu
cudaError_t launch_stream_step(const Par& P, const State& S, const RunArgs& A,
                               cudaStream_t stream);
cudaError_t launch_stream_sweep(const Par& P, const State& S, const RunArgs& A,
                                cudaStream_t stream);
bool combination_supported(int pot, int inter);

} // namespace fqsb
