// fqsb_host.h -- host-side declarations shared by the C-ABI translation unit and the kernel
// instantiation units.
#pragma once

#include "fqsb_device.cuh"

namespace fqsb {

// blocks-per-thread x threads (x wells-in-shared-memory) configurations of the resident kernels
struct ResidentCfg {
    int B, T;
    bool ysmem;
};

static const ResidentCfg kResidentCfgs[] = {
    {1, 256, false},  // 0: N <= 256
    {2, 512, false},  // 1: N <= 1024 (1.43 us/step at N = 1000, vs 2.07 with 1 x 1024)
    {4, 512, false},  // 2: N <= 2048
    {8, 512, false},  // 3: N <= 4096
};
static const int kNumResidentCfgs = 4;

// largest line the resident kernels hold on chip
static const i64 kResidentMaxN = 4096;

// default configuration for N blocks; `variant` > 0 forces kResidentCfgs[variant - 1]
// (B == 0 if the line does not fit on chip)
inline ResidentCfg resident_cfg(i64 N, int variant = 0)
{
    if (variant > 0 && variant <= kNumResidentCfgs) {
        ResidentCfg c = kResidentCfgs[variant - 1];
        if ((i64)c.B * c.T >= N) {
            return c;
        }
    }
    if (N <= 256) {
        return kResidentCfgs[0];
    }
    if (N <= 1024) {
        return kResidentCfgs[1];
    }
    if (N <= 2048) {
        return kResidentCfgs[2];
    }
    if (N <= kResidentMaxN) {
        return kResidentCfgs[3];
    }
    return {0, 0, false};
}

// mu, m and k1 exactly 1: the kernels skip those multiplications (bit-identical results)
inline bool unit_parameters(const Par& P) { return P.mu == 1.0 && P.m == 1.0 && P.k1 == 1.0; }

// dynamic shared memory of k_resident (must mirror the carve-up in the kernel)
inline size_t resident_smem(const Par& P, const ResidentCfg& c, bool stop = true)
{
    const size_t n = (size_t)P.N;
    const size_t ghosts = P.inter < INT_LAPLACE2D ? 2 : 0;
    // slip buffers: nearest-neighbour lines use the transposed slots j*T + t (B*T of them)
    const bool nn1d = P.inter < INT_LAPLACE2D && P.inter != INT_LONGRANGE1D;
    size_t words = 2 * (nn1d ? (size_t)c.B * c.T : n + ghosts) + n + (c.ysmem ? 2 * n : 0) +
                   (P.inter == INT_LONGRANGE1D ? n : 0) + 4 * (size_t)(c.T / 32);
    // + the parked partial sums of the stop modes ([FQSB_SKIP_K - 1][T] pairs + their warp sums)
    return words * 8 + (size_t)(c.T / 32) * 8 * sizeof(int) + n * sizeof(int) +
           (stop ? (size_t)(FQSB_SKIP_K - 1) * (c.T * 16 + (c.T / 32) * 16) : 0);
}

// dynamic shared memory of k_resident_nopassing: us[2][N], sst[N], red[NW][2]
inline size_t resident_np_smem(const Par& P, const ResidentCfg& c)
{
    return (3 * (size_t)P.N + 2 * (size_t)(c.T / 32)) * 8;
}

// defined in fqsb_resident.cu (one object per potential x interaction combination)
cudaError_t launch_resident(const ResidentCfg& cfg, const Par& P, const State& S,
                            const RunArgs& A, cudaStream_t stream);
cudaError_t launch_resident_nopassing(const ResidentCfg& cfg, const Par& P, const State& S,
                                      const RunArgs& A, cudaStream_t stream);
// defined in fqsb_blocked.cu (one object per potential x interaction combination): K2b, the
// temporally blocked kernel for 1-D lines beyond one CTA
struct BlockedPlan {
    int B;      // blocks per thread (tile capacity B * 512 local blocks)
    int own;    // owned blocks per tile
    int H;      // halo blocks on each side
    int ksteps; // steps per launch
    int ntiles;
};
bool blocked_supported(const Par& P);
BlockedPlan blocked_plan(const Par& P, int ksteps_hint, int own_hint);
// dynamic shared memory of k_blocked (must mirror the carve-up in the kernel)
inline size_t blocked_smem(int B, bool stop = true)
{
    const size_t lmax = (size_t)B * FQSB_BK_T;
    // edge slips [2][2][T + 2], pcg32 states, parked sums (stop modes), per-step logs, well-move
    // counters
    return 4 * ((size_t)FQSB_BK_T + 2) * 8 + lmax * 8 +
           (stop ? (size_t)2 * FQSB_BK_PARK * FQSB_BK_T * 16 : 0) +
           FQSB_BK_MAXSTEPS * (2 * 8 + 4 * 4) + lmax * 4;
}
cudaError_t launch_blocked(const BlockedPlan& plan, const Par& P, const State& S,
                           const RunArgs& A, const BlockedArgs& K, cudaStream_t stream);
cudaError_t launch_blocked_begin(const Par& P, const State& S, int ksteps, i64 max_steps,
                                 cudaStream_t stream);
cudaError_t launch_blocked_fixed_done(const Par& P, const State& S, i64 nsteps, int flip,
                                      cudaStream_t stream);
cudaError_t launch_blocked_settle(const Par& P, const State& S, const BlockedArgs& K,
                                  cudaStream_t stream);
// defined in fqsb_thermal.cu: K2t, the resident kernel of the thermal systems (fixed-step calls)
size_t resident_thermal_smem(const Par& P, const ResidentCfg& c);
cudaError_t launch_resident_thermal(const ResidentCfg& cfg, const Par& P, const State& S,
                                    const RunArgs& A, const Thermal& TH, cudaStream_t stream);
// defined in fqsb_stream.cu
// `flip`: parity of the launch within the call (which buffer set is the input);
// `finalise`: run the per-step stop decision (stop modes and flowSteps)
cudaError_t launch_stream_step(const Par& P, const State& S, const RunArgs& A,
                               cudaStream_t stream, int flip, int finalise);
int stream_step_tiles(const Par& P, int generic_tiles);
int stream_launch_tiles(const Par& P, int generic_tiles, bool sweep);
int plan_band_rows(const Par& P, int resident, double halo, int max_tiles);
const char* stream_step_name(const Par& P);
// no-passing: launch l decides sweep l-1 and performs sweep l (`first`: nothing to decide yet)
cudaError_t launch_stream_sweep(const Par& P, const State& S, const RunArgs& A,
                                cudaStream_t stream, int flip, int first, int do_sweep);
bool combination_supported(int pot, int inter);
// defined in fqsb_longrange.cu (K7: DMMA Toeplitz GEMM)
size_t lr_gemm_smem(i64 N);
cudaError_t launch_lr_gemm(const Par& P, const double* tab, const double* W, double* Y,
                           cudaStream_t stream);
cudaError_t launch_lr_step(const Par& P, const State& S, const RunArgs& A, const double* tab,
                           double rowsum, double* W, double* Y, cudaStream_t stream, int finalise);

} // namespace fqsb
