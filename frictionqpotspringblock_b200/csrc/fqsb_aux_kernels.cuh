// fqsb_aux_kernels.cuh -- non-template helper kernels (construction, well alignment K4, force
// materialisation K8, reductions K3/K5, chunk views, streaming settle). Included by fqsb_api.cu
// only.
#pragma once

#include "fqsb_kernels.cuh"

namespace fqsb {

// after a streaming call: realisations whose current state sits in the second buffer set are
// copied back (and quenched when they converged, detail.h:1527-1532)
__global__ void k_stream_settle(const Par P, const State S, int with_va)
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
        if (flip) {
            S.u[base + p] = S.u2[base + p];
        }
        if (quench) {
            S.v[base + p] = 0.0;
            S.a[base + p] = 0.0;
        }
        else if (flip && with_va) {
            S.v[base + p] = S.v2[base + p];
            S.a[base + p] = S.a2[base + p];
        }
    }
}

// fixed-step streaming calls keep no per-step bookkeeping on the device: settle it at the end
__global__ void k_stream_fixed_done(const Par P, const State S, i64 nsteps, int inplace)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < P.R) {
        Ctl& c = S.ctl[r];
        c.inc += nsteps; // detail.h:1541
        c.steps = nsteps;
        c.flip = inplace ? 0 : (int)(nsteps & 1);
        c.status = ST_EXHAUSTED;
    }
}

// LongRange GEMM path: W = u - u_frame, and f_interactions = Y - rowsum * W
__global__ void k_lr_shift(const Par P, const State S, double* W)
{
    const i64 n = P.N * P.R;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < n; g += (i64)gridDim.x * blockDim.x) {
        W[g] = S.u[g] - S.u_frame[g / P.N];
    }
}

__global__ void k_stream_settle_flags(const Par P, const State S)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < P.R) {
        S.ctl[r].flip = 0;
    }
}

// =============================================================================================
// External = RandomNormalForcing: updated_inc() (detail.h:1369-1375 -> 931-943). Block p is
// redrawn when inc >= next[p]; the k-th due block (in block order) takes the k-th next draw of
// the realisation's single pcg32 stream. One CTA per realisation: ordered compaction (ballot +
// scan of the warp counts) gives every due block its rank, an O(log rank) LCG jump gives it its
// generator state; the stream then advances by the number of draws. `inc_add` = 1 when called
// ahead of a time step (m_inc++ comes first, detail.h:1541-1544), 0 for refresh()/set_inc();
// `only_running`: skip realisations whose stepping call has already ended.
// normal(mu, sigma) = mu + sigma*sqrt(2)*erf_inv(2r - 1), r = next_double() (prrng; the device
// uses erf_inv_dev, a few 1e-16 relative from boost's long-double evaluation).
// =============================================================================================
__global__ void __launch_bounds__(1024) k_thermal_draw(const Par P, const State S, const Thermal T,
                                                       int inc_add, int only_running)
{
    __shared__ int wcount[32];
    __shared__ int wexcl[33];
    const int r = blockIdx.x;
    const Ctl& ctl = S.ctl[r];
    if (only_running && ctl.status != ST_RUNNING) {
        return;
    }
    const i64 inc = ctl.inc + inc_add;
    const i64 base = (i64)r * P.N;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const u64 st0 = T.state[r];
    u64 carry = 0;
    for (i64 p0 = 0; p0 < P.N; p0 += blockDim.x) {
        const i64 p = p0 + t;
        const bool due = p < P.N && inc >= T.next[base + p];
        const unsigned bal = __ballot_sync(0xffffffffu, due);
        if (lane == 0) {
            wcount[warp] = __popc(bal);
        }
        __syncthreads();
        if (warp == 0) {
            int c = lane < (int)(blockDim.x >> 5) ? wcount[lane] : 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) {
                    incl += y;
                }
            }
            wexcl[lane] = incl - c;
            if (lane == 31) {
                wexcl[32] = incl;
            }
        }
        __syncthreads();
        if (p < P.N) {
            double val = T.f_ext[base + p];
            if (due) {
                const u64 rank = carry + (u64)wexcl[warp] + (u64)__popc(bal & ((1u << lane) - 1u));
                const u64 st = pcg_advance_inc(st0, rank, T.inc_rng);
                const double z = pcg_double(st);
                val = T.mean + T.sigma_sqrt2 * erf_inv_dev(2.0 * z - 1.0);
                T.f_ext[base + p] = val;
                T.next[base + p] += T.dinc[base + p];
            }
            T.f_sys[base + p] = val; // std::copy(m_f_thermal..., f), detail.h:942
        }
        carry += (u64)wexcl[32];
        __syncthreads();
    }
    if (t == 0 && carry) {
        T.state[r] = pcg_advance_inc(st0, carry, T.inc_rng);
    }
}

// =============================================================================================
// construction, alignment, forces, reductions, chunk views
// =============================================================================================

// Line1d.h:148-157 / Line2d.h:28-34: initstate = seed + flat index, y = cumsum + offset; then
// detail.h:1127-1138: u = v = a = 0 and the first align.
__global__ void k_init(const Par P, const State S)
{
    const i64 n = P.N * P.R;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < n; g += (i64)gridDim.x * blockDim.x) {
        const i64 r = g / P.N, p = g - r * P.N;
        u64 gp = (u64)p;
        if (P.seed_period) { // slab: local block p is global block (seed_first + p) mod period
            gp = (P.seed_first + (u64)p) % P.seed_period;
        }
        u64 st = pcg_seed(P.seed + (u64)r * P.seed_stride + gp);
        double d0 = spacing_peek(P, st);
        if (P.consumes) {
            st = pcg_next(st);
        }
        // virtual well i = -1: (y[-1], y[0]] = (offset, offset + d_0]
        double yl = P.offset, yr = P.offset + d0;
        int underflow = 0;
        int moved = well_align(P, 0.0, yl, yr, st, -1, &underflow);
        if (-1 + moved < 0) {
            S.err[0] = 1;
        }
        S.u[g] = 0.0;
        S.v[g] = 0.0;
        S.a[g] = 0.0;
        S.yl[g] = yl;
        S.yr[g] = yr;
        S.idx[g] = -1 + moved;
        S.rng[g] = st;
    }
}

// updated_u()'s m_chunk->align(u) for every block (detail.h:144); optional uniform shift
// u += du[r] first (advanceUniformly, detail.h:2041)
__global__ void k_align(const Par P, const State S, const double* du)
{
    const i64 n = P.N * P.R;
    int underflow = 0;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < n; g += (i64)gridDim.x * blockDim.x) {
        double u = S.u[g];
        if (du) {
            u += du[g / P.N];
            S.u[g] = u;
        }
        double yl = S.yl[g], yr = S.yr[g];
        if (u > yr || !(u > yl)) {
            u64 st = S.rng[g];
            i64 i0 = S.idx[g];
            int moved = well_align(P, u, yl, yr, st, i0, &underflow);
            S.rng[g] = st;
            S.idx[g] = i0 + moved;
            S.yl[g] = yl;
            S.yr[g] = yr;
        }
    }
    if (underflow) {
        S.err[0] = 1;
    }
}

// `system.chunk.align(u)` (python-prrng pcg32_tensor_cumsum::align): the wells follow the given
// positions; the system's slips are left alone
__global__ void k_align_to(const Par P, const State S, const double* u)
{
    const i64 n = P.N * P.R;
    int underflow = 0;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < n; g += (i64)gridDim.x * blockDim.x) {
        const double ug = u[g];
        double yl = S.yl[g], yr = S.yr[g];
        if (ug > yr || !(ug > yl)) {
            u64 st = S.rng[g];
            i64 i0 = S.idx[g];
            int moved = well_align(P, ug, yl, yr, st, i0, &underflow);
            S.rng[g] = st;
            S.idx[g] = i0 + moved;
            S.yl[g] = yl;
            S.yr[g] = yr;
        }
    }
    if (underflow) {
        S.err[0] = 1;
    }
}

// K8: materialise force arrays on demand (getters, detail.h:1429-1468).
// mask bits: 1 potential, 2 interactions, 4 frame, 8 damping; f is always re-summed
// (detail.h:1324) from the stored components.
struct ForceArrays {
    double *f, *f_pot, *f_int, *f_frame, *f_damp;
    const double *lr_w, *lr_y; // mask bit 16: f_int = lr_y - lr_rowsum * lr_w (K7 path)
    double lr_rowsum;
};

__global__ void k_forces(const Par P, const State S, const ForceArrays F, int mask)
{
    const i64 n = P.N * P.R;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < n; g += (i64)gridDim.x * blockDim.x) {
        const i64 r = g / P.N;
        const int p = (int)(g - r * P.N);
        const double* ur = S.u + r * P.N;
        const double uc = ur[p];
        if (mask & 1) {
            F.f_pot[g] = f_potential_rt(P, uc, S.yl[g], S.yr[g]);
        }
        if (mask & 16) {
            F.f_int[g] = F.lr_y[g] - F.lr_rowsum * F.lr_w[g];
        }
        else if (mask & 2) {
            int i = 0, j = 0;
            if (P.rank == 2) {
                i = p / P.cols;
                j = p - i * P.cols;
            }
            auto U = [&](int q) { return ur[q]; };
            F.f_int[g] = f_interactions_rt(P, U, S.pref, p, i, j, uc);
        }
        if (mask & 4) {
            F.f_frame[g] = P.k_frame * (S.u_frame[r] - uc);
        }
        if (mask & 8) {
            F.f_damp[g] = -P.eta * S.v[g];
        }
        if (S.f_thermal) { // detail.h:1326-1329
            F.f[g] = F.f_frame[g] + F.f_pot[g] + F.f_int[g] + F.f_damp[g] + S.f_thermal[g];
        }
        else {
            F.f[g] = F.f_frame[g] + F.f_pot[g] + F.f_int[g] + F.f_damp[g];
        }
    }
}

// per-realisation reductions, two stages with fixed order.
// what: 0 residual sums from stored arrays {sum f^2, sum f_frame^2}
//       1 residual sums derived on the fly from the state
//       2 {sum v^2, sum f_frame} (temperature, mean frame force)
//       3 {min displacement, off-branch flag} (maxUniformDisplacement, detail.h:176-187,282-334)
//       4 {sum (i - i_n), #(i != i_n)} and part[2] = sum |i - i_n|
__global__ void __launch_bounds__(256) k_reduce(const Par P, const State S, const ForceArrays F,
                                                int what, int direction, const i64* i_n,
                                                double* part /* [R][tiles][4] */, int own_lo,
                                                int own_hi)
{
    __shared__ double scratch[32 * 3];
    const int r = blockIdx.y;
    const i64 base = (i64)r * P.N;
    const int N = (int)P.N;
    double acc[3] = {0.0, 0.0, 0.0};
    double mn = 1.7976931348623157e308;
    const double uf = S.u_frame[r];
    const double* ur = S.u + base;
    for (int p = own_lo + blockIdx.x * blockDim.x + threadIdx.x; p < own_hi;
         p += gridDim.x * blockDim.x) {
        const i64 g = base + p;
        if (what == 0) {
            acc[0] += F.f[g] * F.f[g];
            acc[1] += F.f_frame[g] * F.f_frame[g];
        }
        else if (what == 1) {
            const double uc = ur[p];
            int i = 0, j = 0;
            if (P.rank == 2) {
                i = p / P.cols;
                j = p - i * P.cols;
            }
            auto U = [&](int q) { return ur[q]; };
            double fi = f_interactions_rt(P, U, S.pref, p, i, j, uc);
            double fp = f_potential_rt(P, uc, S.yl[g], S.yr[g]);
            double ff = P.k_frame * (uf - uc);
            double f = ff + fp + fi + (-P.eta * S.v[g]);
            if (S.f_thermal) {
                f += S.f_thermal[g];
            }
            acc[0] += f * f;
            acc[1] += ff * ff;
        }
        else if (what == 2) {
            acc[0] += S.v[g] * S.v[g];
            acc[1] += P.k_frame * (uf - ur[p]);
        }
        else if (what == 3) {
            const double uc = ur[p], yl = S.yl[g], yr = S.yr[g];
            if (P.pot == POT_CUSPY) {
                mn = fmin(mn, direction > 0 ? yr - uc : uc - yl);
            }
            else {
                double xi = 0.5 * (yl + yr);
                double u_r = (P.mu * xi + P.kappa * yr) / (P.mu + P.kappa);
                double u_l = (P.mu * xi + P.kappa * yl) / (P.mu + P.kappa);
                if (uc < u_l || !(uc <= u_r)) {
                    acc[1] += 1.0;
                }
                else {
                    mn = fmin(mn, direction > 0 ? u_r - uc : uc - u_l);
                }
            }
        }
        else {
            i64 d = S.idx[g] - i_n[g];
            acc[0] += (double)d;
            acc[1] += d != 0 ? 1.0 : 0.0;
            acc[2] += (double)(d < 0 ? -d : d);
        }
    }
    block_sum<3>(acc, scratch);
    mn = warp_min(mn);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        scratch[threadIdx.x >> 5] = mn;
    }
    __syncthreads();
    mn = warp_min((threadIdx.x & 31) < (blockDim.x >> 5) ? scratch[threadIdx.x & 31]
                                                         : 1.7976931348623157e308);
    if (threadIdx.x == 0) {
        double* o = part + ((size_t)r * gridDim.x + blockIdx.x) * 4;
        o[0] = acc[0];
        o[1] = acc[1];
        o[2] = acc[2];
        o[3] = mn;
    }
}

__global__ void k_reduce_final(const double* part, int tiles, double* out /* [R][4] */)
{
    const int r = blockIdx.x;
    const int lane = threadIdx.x;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, mn = 1.7976931348623157e308;
    for (int c = lane; c < tiles; c += 32) {
        const double* o = part + ((size_t)r * tiles + c) * 4;
        a0 += o[0];
        a1 += o[1];
        a2 += o[2];
        mn = fmin(mn, o[3]);
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    mn = warp_min(mn);
    if (lane == 0) {
        out[4 * r] = a0;
        out[4 * r + 1] = a1;
        out[4 * r + 2] = a2;
        out[4 * r + 3] = mn;
    }
}

// slab halo exchange: the full state of `count` consecutive blocks as 7 planes of 8-byte words
// (u, v, a, y_l, y_r, idx, rng)
__global__ void k_export_cells(const State S, i64 first, i64 count, u64* buf)
{
    for (i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x; k < count; k += (i64)gridDim.x * blockDim.x) {
        const i64 g = first + k;
        buf[k] = (u64)__double_as_longlong(S.u[g]);
        buf[count + k] = (u64)__double_as_longlong(S.v[g]);
        buf[2 * count + k] = (u64)__double_as_longlong(S.a[g]);
        buf[3 * count + k] = (u64)__double_as_longlong(S.yl[g]);
        buf[4 * count + k] = (u64)__double_as_longlong(S.yr[g]);
        buf[5 * count + k] = (u64)S.idx[g];
        buf[6 * count + k] = S.rng[g];
    }
}

__global__ void k_import_cells(const State S, i64 first, i64 count, const u64* buf)
{
    for (i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x; k < count; k += (i64)gridDim.x * blockDim.x) {
        const i64 g = first + k;
        S.u[g] = __longlong_as_double((i64)buf[k]);
        S.v[g] = __longlong_as_double((i64)buf[count + k]);
        S.a[g] = __longlong_as_double((i64)buf[2 * count + k]);
        S.yl[g] = __longlong_as_double((i64)buf[3 * count + k]);
        S.yr[g] = __longlong_as_double((i64)buf[4 * count + k]);
        S.idx[g] = (i64)buf[5 * count + k];
        S.rng[g] = buf[6 * count + k];
    }
}

// quench(): v = a = 0 (detail.h:1527-1532)
__global__ void k_zero_va(const Par P, const State S)
{
    const i64 n = P.N * P.R;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < n; g += (i64)gridDim.x * blockDim.x) {
        S.v[g] = 0.0;
        S.a[g] = 0.0;
    }
}

// `system.chunk.data`: y[p, first[p] .. first[p]+ny) regenerated from the current well
__global__ void k_chunk_data(const Par P, const State S, const i64* first, int ny, double* out)
{
    const i64 n = P.N * P.R;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < n; g += (i64)gridDim.x * blockDim.x) {
        double yl = S.yl[g], yr = S.yr[g];
        u64 st = S.rng[g];
        i64 i = S.idx[g];
        const i64 f0 = first[g];
        double* o = out + g * (i64)ny;
        // walk the window (yl = y[i], yr = y[i+1]) so that i == f0
        while (i > f0) {
            u64 sb = st;
            if (P.consumes) {
                st = pcg_prev(st);
                sb = pcg_prev(st);
            }
            yr = yl;
            yl = yl - spacing_peek(P, sb);
            --i;
        }
        while (i < f0) {
            double d = spacing_peek(P, st);
            if (P.consumes) {
                st = pcg_next(st);
            }
            yl = yr;
            yr = yr + d;
            ++i;
        }
        for (int k = 0; k < ny; ++k) {
            o[k] = yl;
            double d = spacing_peek(P, st);
            if (P.consumes) {
                st = pcg_next(st);
            }
            yl = yr;
            yr = yr + d;
        }
    }
}

// `system.chunk.restore(state, value, index)`: y[index] = value, generator as state_at(index)
__global__ void k_chunk_restore(const Par P, const State S, const u64* state, const double* value,
                                const i64* index)
{
    const i64 n = P.N * P.R;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < n; g += (i64)gridDim.x * blockDim.x) {
        u64 st = state[g];
        double d = spacing_peek(P, st); // d_index
        S.yr[g] = value[g];
        S.yl[g] = value[g] - d;
        S.idx[g] = index[g] - 1;
        S.rng[g] = P.consumes ? pcg_next(st) : st;
    }
}

} // namespace fqsb
