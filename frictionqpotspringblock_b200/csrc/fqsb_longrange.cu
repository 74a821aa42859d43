// fqsb_longrange.cu -- K7: the LongRange 1/r^(1+alpha) interaction of an ensemble as a batched
// circulant mat-vec on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64).
//
// Reference: LongRange1d::force (detail.h:849-867), an O(N^2) scalar loop
//     f_p = sum_{q != p} (u_q - u_p) * pref[min(|q-p|, N-|q-p|)],   pref[d] = k / d^(alpha+1).
// For R realisations this is   F[N x R] = C[N x N] . W[N x R] - rowsum * W   with the circulant
// C[p][q] = tab[(p - q) mod N], tab[d] = pref[min(d, N - d)], tab[0] = 0, and W = u - c_r
// (c_r = u_frame of the realisation: the result is invariant under a uniform shift and the shift
// removes the cancellation between the two terms when u ~ 1e4).
// C is never formed: the A fragments of the MMA are read from the 1-D table in shared memory.
// tcgen05 has no FP64 kind, so on sm_100a the FP64 tensor path is mma.sync.
//
// One time step of a LongRange system that does not fit the resident kernel is three launches:
//   k_lr_positions : u += dt*v + 0.5*dt^2*a (detail.h:1549), W = u - c_r
//   k_lr_gemm      : Y = C . W                          (this file, tensor cores)
//   k_lr_finish    : well search, f_int = Y - rowsum*W, forces, Verlet tail, stop decision
#include "fqsb_host.h"
#include "fqsb_kernels.cuh"

namespace fqsb {

#define LR_BM 128 // output rows (blocks p) per CTA
#define LR_BN 64  // realisations per CTA
#define LR_BK 32  // q per pipeline stage
#define LR_LD (LR_BK + 4) // smem row stride of the W tile: conflict-free B fragments

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int bytes = valid ? 16 : 0; // zero-fill when out of range
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}

// Y[r][p] = sum_q tab[(p - q) mod N] * W[r][q]      (W, Y row-major [R][N])
__global__ void __launch_bounds__(256)
    k_lr_gemm(const double* __restrict__ tab, const double* __restrict__ W, double* __restrict__ Y,
              const int N, const int R)
{
    extern __shared__ __align__(16) unsigned char lr_smem[];
    double* stab = reinterpret_cast<double*>(lr_smem);     // [N]
    double* sW = stab + ((N + 1) & ~1);                    // [2][LR_BN][LR_LD]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int p0 = blockIdx.x * LR_BM, r0 = blockIdx.y * LR_BN;
    const int wm = (warp & 3) * 32; // warp tile: 32 rows x 32 realisations
    const int wn = (warp >> 2) * 32;

    for (int k = t; k < N; k += 256) {
        stab[k] = tab[k];
    }

    auto load_stage = [&](int buf, int q0) {
        // LR_BN rows x LR_BK doubles = 64 x 16 chunks of 16 B; 4 chunks per thread
        double* dst = sW + (size_t)buf * LR_BN * LR_LD;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int chunk = t + c * 256;
            const int row = chunk >> 4, col = (chunk & 15) * 2;
            const int r = r0 + row, q = q0 + col;
            const bool ok = r < R && q < N; // N is even on this path, so q + 1 < N as well
            cp_async16(dst + row * LR_LD + col, W + (size_t)(ok ? r : 0) * N + (ok ? q : 0), ok);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };

    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            acc[mi][ni][0] = 0.0;
            acc[mi][ni][1] = 0.0;
        }
    }

    const int nk = (N + LR_BK - 1) / LR_BK;
    load_stage(0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) {
            load_stage((kt + 1) & 1, (kt + 1) * LR_BK);
            asm volatile("cp.async.wait_group 1;\n" ::);
        }
        else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
        const double* ws = sW + (size_t)(kt & 1) * LR_BN * LR_LD;
        const int q0 = kt * LR_BK;
#pragma unroll
        for (int kk = 0; kk < LR_BK / 4; ++kk) {
            double a[4], b[4];
            // A[p][q] = tab[(p - q) mod N], p = p0 + wm + mi*8 + g, q = q0 + kk*4 + t4
            int d = (p0 + wm + g) - (q0 + kk * 4 + t4);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                int dd = d + mi * 8;
                dd = dd < 0 ? dd + N : (dd >= N ? dd - N : dd);
                dd = dd < 0 ? dd + N : dd; // p0 + ... may exceed N on a ragged last tile
                a[mi] = stab[dd];
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                b[ni] = ws[(wn + ni * 8 + g) * LR_LD + kk * 4 + t4];
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
                }
            }
        }
        __syncthreads();
    }

    // D fragment: rows g, columns 2*t4, 2*t4+1
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int p = p0 + wm + mi * 8 + g;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int r = r0 + wn + ni * 8 + 2 * t4;
            if (p < N) {
                if (r < R) {
                    Y[(size_t)r * N + p] = acc[mi][ni][0];
                }
                if (r + 1 < R) {
                    Y[(size_t)(r + 1) * N + p] = acc[mi][ni][1];
                }
            }
        }
    }
}

// u += dt*v + 0.5*dt^2*a in place (no neighbour reads in this path), W = u - u_frame
__global__ void __launch_bounds__(256)
    k_lr_positions(const __grid_constant__ Par P, const __grid_constant__ State S,
                   const __grid_constant__ RunArgs A, double* __restrict__ W)
{
    const int r = blockIdx.y;
    const Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }
    const i64 base = (i64)r * P.N;
    const double c2 = 0.5 * P.dt * P.dt;
    double uf = S.u_frame[r];
    if (A.flow) {
        uf += A.v_frame * P.dt;
    }
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < (int)P.N; p += gridDim.x * blockDim.x) {
        const i64 gp = base + p;
        double un = S.u[gp] + P.dt * S.v[gp] + c2 * S.a[gp];
        S.u[gp] = un;
        W[gp] = un - uf;
    }
}

__global__ void __launch_bounds__(256)
    k_lr_finish(const __grid_constant__ Par P, const __grid_constant__ State S,
                const __grid_constant__ RunArgs A, const double* __restrict__ W,
                const double* __restrict__ Y, const double rowsum, const int finalise)
{
    __shared__ double scratch[32 * 5];
    __shared__ int iscratch[32 * 4];
    __shared__ int s_last;
    const int r = blockIdx.y;
    Ctl& ctl = S.ctl[r];
    if (ctl.status != ST_RUNNING) {
        return;
    }
    const i64 base = (i64)r * P.N;
    double uf = S.u_frame[r];
    if (A.flow) {
        uf += A.v_frame * P.dt;
    }
    double acc[2] = {0.0, 0.0};
    int hops = 0, dS = 0, dA = 0, underflow = 0;
    bool nan = false;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < (int)P.N; p += gridDim.x * blockDim.x) {
        const i64 gp = base + p;
        const double un = S.u[gp];
        double yl = S.yl[gp], yr = S.yr[gp];
        if (un > yr || !(un > yl)) {
            u64 st = S.rng[gp];
            i64 i_before = S.idx[gp];
            int moved = well_align(P, un, yl, yr, st, i_before, &underflow);
            S.rng[gp] = st;
            S.idx[gp] = i_before + moved;
            S.yl[gp] = yl;
            S.yr[gp] = yr;
            hops += moved != 0;
            track_hop(A, gp, i_before, moved, dS, dA);
        }
        double fi = Y[gp] - rowsum * W[gp];
        double fp = f_potential_rt(P, un, yl, yr);
        double ff = P.k_frame * (uf - un);
        double F = ff + fp + fi;
        double v = S.v[gp], a = S.a[gp];
        double f = verlet_tail(P, F, v, a);
        S.v[gp] = v;
        S.a[gp] = a;
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
    if (finalise) {
        stream_finalise(P, S, A, r, /*flip=*/1, uf, acc, hops, dS, dA, scratch, iscratch, &s_last);
    }
}

size_t lr_gemm_smem(i64 N)
{
    return (size_t)(((N + 1) & ~(i64)1) + 2 * LR_BN * LR_LD) * sizeof(double);
}

// forces only (getters / residual): Y = C . (u - u_frame)
cudaError_t launch_lr_gemm(const Par& P, const double* tab, const double* W, double* Y,
                           cudaStream_t stream)
{
    const size_t smem = lr_gemm_smem(P.N);
    cudaError_t e = cudaFuncSetAttribute(k_lr_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
        return e;
    }
    dim3 grid((unsigned)((P.N + LR_BM - 1) / LR_BM), (unsigned)((P.R + LR_BN - 1) / LR_BN));
    k_lr_gemm<<<grid, 256, smem, stream>>>(tab, W, Y, (int)P.N, (int)P.R);
    return cudaGetLastError();
}

cudaError_t launch_lr_step(const Par& P, const State& S, const RunArgs& A, const double* tab,
                           double rowsum, double* W, double* Y, cudaStream_t stream, int finalise)
{
    dim3 grid((unsigned)S.tiles, (unsigned)P.R);
    k_lr_positions<<<grid, 256, 0, stream>>>(P, S, A, W);
    cudaError_t e = launch_lr_gemm(P, tab, W, Y, stream);
    if (e != cudaSuccess) {
        return e;
    }
    k_lr_finish<<<grid, 256, 0, stream>>>(P, S, A, W, Y, rowsum, finalise);
    return cudaGetLastError();
}

} // namespace fqsb
