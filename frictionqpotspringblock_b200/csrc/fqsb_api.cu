// fqsb_api.cu -- implementation of the C ABI declared in include/fqsb.h.
// Host-side orchestration only: every number is produced by the kernels in fqsb_kernels.cuh.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <cstdlib>
#include <string>
#include <vector>

// the helper kernels of this unit (k_init, k_align, k_chunk_data, ...) draw from every distribution
#define FQSB_SLOW_DISTS
#include "../../include/fqsb.h"
#include "fqsb_host.h"
#include "fqsb_aux_kernels.cuh"

using namespace fqsb;

static thread_local std::string g_err;

static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

static int cuda_fail(cudaError_t e, const char* what)
{
    return fail(FQSB_ECUDA, std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
}

#define CU(x) \
    do { \
        cudaError_t e_ = (x); \
        if (e_ != cudaSuccess) { \
            return cuda_fail(e_, #x); \
        } \
    } while (0)

#define TRY(x) \
    do { \
        int rc_ = (x); \
        if (rc_ != FQSB_OK) { \
            return rc_; \
        } \
    } while (0)

// the reference's assertion text (config.h:19-24)
#define ASSERT_MSG(expr) (std::string("fqsb: assertion failed (") + expr + ") \n\t")

struct fqsb_slab_state;
static void slab_free(struct fqsb_system* s);

struct fqsb_system {
    fqsb_params par;
    Par P;
    State S;
    ForceArrays F;
    cudaStream_t stream;
    bool own_stream;
    int device;
    i64 N, R, n;
    bool forces_valid;  // F holds the forces of the current state
    bool forces_frozen; // F is kept as is (stale after trigger(), detail.h:1975-1976)
    double* d_red;      // [R][tiles][8] partial sums
    double* d_out;      // [R][4]
    double* d_du;       // [R]
    i64* d_in;          // [R*N] i_n scratch
    void* d_scratch;    // generic staging
    size_t scratch_bytes;
    double* h_out;      // pinned [R][4]
    Ctl* h_ctl;         // pinned [R]
    int* h_err;         // pinned [2]
    double* d_pref;
    // K7 (LongRange as a DMMA Toeplitz GEMM) for lines beyond the resident kernel
    bool lr_gemm;
    double *d_lr_tab, *d_lr_w, *d_lr_y;
    double lr_rowsum;
    // slab decomposition support: owned range, state snapshot, log buffer
    i64 own_lo, own_hi;
    double *snap_d[5];
    i64* snap_idx;
    u64* snap_rng;
    double* snap_uf;
    Ctl* snap_ctl;
    bool snap_valid;
    double* d_log;
    size_t log_cap;
    // K2b (temporally blocked tiles of a long 1-D line): second set of the well arrays, log
    BlockedArgs bk;
    bool bk_ready;
    int bk_log_tiles;
    // External = RandomNormalForcing (detail.h:881-1000): enabled by fqsb_enable_random_forcing
    bool thermal;
    Thermal th;
    // member of a slab-decomposed system (fqsb_slab.inl)
    fqsb_slab_state* slab;
    i64* d_mark;     // [R*N] indices kept by fqsb_mark_indices (nullptr until then)
    // internal streams of fqsb_run_from_host (chunks of realisations in flight side by side)
    cudaStream_t pipe_stream[4];
    cudaEvent_t pipe_ev;
    i64 launches, steps;
    const char* last_kernel;
    cudaEvent_t ev0, ev1;    // bracket the stepping-kernel launches of the last dynamics call
    double kernel_ms;        // device time between them, summed over the call's launches
    i64 kernel_launches;     // stepping-kernel launches of the last dynamics call
    std::vector<void*> allocs;
};

static unsigned grid_for(i64 n)
{
    i64 g = (n + 255) / 256;
    const i64 cap = 148 * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <class T>
static int dev_alloc(fqsb_system* s, T** p, size_t count)
{
    void* q = nullptr;
    CU(cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 1));
    s->allocs.push_back(q);
    *p = (T*)q;
    return FQSB_OK;
}

static int scratch(fqsb_system* s, size_t bytes)
{
    if (bytes > s->scratch_bytes) {
        if (s->d_scratch) {
            CU(cudaStreamSynchronize(s->stream));
            CU(cudaFree(s->d_scratch));
            s->d_scratch = nullptr;
            s->scratch_bytes = 0;
        }
        CU(cudaMalloc(&s->d_scratch, bytes));
        s->scratch_bytes = bytes;
    }
    return FQSB_OK;
}

static int enter(fqsb_system* s)
{
    if (!s) {
        return fail(FQSB_EASSERT, "null handle");
    }
    CU(cudaSetDevice(s->device));
    return FQSB_OK;
}

static int check_flags(fqsb_system* s)
{
    CU(cudaMemcpyAsync(s->h_err, s->S.err, 2 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (s->h_err[0] || s->h_err[1]) {
        int e0 = s->h_err[0], e1 = s->h_err[1];
        CU(cudaMemsetAsync(s->S.err, 0, 2 * sizeof(int), s->stream));
        if (e1) {
            return fail(FQSB_ENAN, "NaN entries found"); // detail.h:1568
        }
        if (e0) {
            return fail(FQSB_EASSERT,
                        "yield landscape exhausted below its first entry (lower the offset)");
        }
    }
    return FQSB_OK;
}

static void invalidate_forces(fqsb_system* s)
{
    s->forces_valid = false;
    s->forces_frozen = false;
}

static int ensure_forces(fqsb_system* s)
{
    if (!s->F.f) {
        TRY(dev_alloc(s, &s->F.f, (size_t)s->n));
        TRY(dev_alloc(s, &s->F.f_pot, (size_t)s->n));
        TRY(dev_alloc(s, &s->F.f_int, (size_t)s->n));
        TRY(dev_alloc(s, &s->F.f_frame, (size_t)s->n));
        TRY(dev_alloc(s, &s->F.f_damp, (size_t)s->n));
        s->forces_valid = false;
        s->forces_frozen = false;
    }
    if (!s->forces_valid && !s->forces_frozen) {
        int mask = 15;
        if (s->lr_gemm) { // f_interactions through the tensor-core GEMM instead of O(N^2) loads
            k_lr_shift<<<grid_for(s->n), 256, 0, s->stream>>>(s->P, s->S, s->d_lr_w);
            cudaError_t e = launch_lr_gemm(s->P, s->d_lr_tab, s->d_lr_w, s->d_lr_y, s->stream);
            if (e != cudaSuccess) {
                return cuda_fail(e, "k_lr_gemm");
            }
            s->F.lr_w = s->d_lr_w;
            s->F.lr_y = s->d_lr_y;
            s->F.lr_rowsum = s->lr_rowsum;
            s->launches += 2;
            mask = 13 | 16;
        }
        k_forces<<<grid_for(s->n), 256, 0, s->stream>>>(s->P, s->S, s->F, mask);
        CU(cudaGetLastError());
        s->launches++;
        s->forces_valid = true;
    }
    return FQSB_OK;
}

// recompute one component inside a frozen (possibly stale) force set
static int frozen_update(fqsb_system* s, int mask)
{
    k_forces<<<grid_for(s->n), 256, 0, s->stream>>>(s->P, s->S, s->F, mask);
    CU(cudaGetLastError());
    s->launches++;
    return FQSB_OK;
}

static int reduce(fqsb_system* s, int what, int direction, const i64* i_n_dev)
{
    dim3 grid((unsigned)s->S.tiles, (unsigned)s->R);
    k_reduce<<<grid, 256, 0, s->stream>>>(s->P, s->S, s->F, what, direction, i_n_dev, s->d_red,
                                          (int)s->own_lo, (int)s->own_hi);
    k_reduce_final<<<(unsigned)s->R, 32, 0, s->stream>>>(s->d_red, s->S.tiles, s->d_out);
    CU(cudaGetLastError());
    s->launches += 2;
    return FQSB_OK;
}

static int reduce_to_host(fqsb_system* s, int what, int direction, const i64* i_n_dev)
{
    TRY(reduce(s, what, direction, i_n_dev));
    CU(cudaMemcpyAsync(s->h_out, s->d_out, (size_t)s->R * 4 * sizeof(double),
                       cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

static int align(fqsb_system* s, const double* du_dev)
{
    k_align<<<grid_for(s->n), 256, 0, s->stream>>>(s->P, s->S, du_dev);
    CU(cudaGetLastError());
    s->launches++;
    return FQSB_OK;
}

static int pull_ctl(fqsb_system* s)
{
    CU(cudaMemcpyAsync(s->h_ctl, s->S.ctl, (size_t)s->R * sizeof(Ctl), cudaMemcpyDeviceToHost,
                       s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

static int push_ctl(fqsb_system* s)
{
    CU(cudaMemcpyAsync(s->S.ctl, s->h_ctl, (size_t)s->R * sizeof(Ctl), cudaMemcpyHostToDevice,
                       s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

// updated_inc() of a thermal system (detail.h:1369-1375): redraw the due blocks, refresh
// System::m_f_thermal; the stored force arrays follow on the next getter
static int thermal_update(fqsb_system* s, int inc_add, int only_running)
{
    if (!s->thermal) {
        return FQSB_OK;
    }
    const unsigned threads = s->N >= 1024 ? 1024u : (unsigned)(((s->N + 31) / 32) * 32);
    k_thermal_draw<<<(unsigned)s->R, threads, 0, s->stream>>>(s->P, s->S, s->th, inc_add,
                                                              only_running);
    CU(cudaGetLastError());
    s->launches++;
    return FQSB_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" {

const char* fqsb_last_error(void) { return g_err.c_str(); }
int fqsb_abi_version(void) { return FQSB_ABI_VERSION; }
const char* fqsb_version(void) { return "0.1.0"; }

int fqsb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void* fqsb_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void fqsb_host_free(void* p)
{
    if (p) {
        cudaFreeHost(p);
    }
}

int fqsb_create(const fqsb_params* par, fqsb_system** out)
{
    if (!par || !out) {
        return fail(FQSB_EASSERT, "null argument");
    }
    *out = nullptr;
    if (par->rank != 1 && par->rank != 2) {
        return fail(FQSB_EASSERT, ASSERT_MSG("rank == 1 || rank == 2"));
    }
    if (par->shape[0] < 1 || (par->rank == 2 && par->shape[1] < 1)) {
        return fail(FQSB_EASSERT, ASSERT_MSG("shape > 0"));
    }
    switch (par->distribution) {
    case FQSB_DIST_RANDOM:
    case FQSB_DIST_DELTA:
    case FQSB_DIST_EXPONENTIAL:
    case FQSB_DIST_POWER:
    case FQSB_DIST_PARETO:
    case FQSB_DIST_WEIBULL:
    case FQSB_DIST_GAMMA:
    case FQSB_DIST_NORMAL:
        break;
    default:
        return fail(FQSB_EASSERT, "Unknown distribution: " + std::to_string(par->distribution));
    }
    if (!combination_supported(par->potential, par->interactions)) {
        return fail(FQSB_EUNSUPPORTED, "potential x interactions combination not available");
    }
    const bool two_d = par->interactions == FQSB_INT_LAPLACE2D ||
                       par->interactions == FQSB_INT_QUARTICGRADIENT2D;
    if (two_d != (par->rank == 2)) {
        return fail(FQSB_EASSERT, ASSERT_MSG("rank matches interactions"));
    }
    if (par->minimisation == FQSB_MIN_OVERDAMPED &&
        !(par->potential == FQSB_POT_CUSPY && (par->interactions == FQSB_INT_LAPLACE1D ||
                                               par->interactions == FQSB_INT_LAPLACE2D))) {
        return fail(FQSB_EUNSUPPORTED, "Minimisation not implementated"); // detail.h:1692,1695
    }
    int ndev = fqsb_device_count();
    if (ndev < 1) {
        return fail(FQSB_ECUDA, "no CUDA device: this library has no CPU fallback");
    }
    int device = par->device;
    if (device < 0) {
        CU(cudaGetDevice(&device));
    }
    if (device >= ndev) {
        return fail(FQSB_ECUDA, "CUDA device ordinal out of range");
    }
    CU(cudaSetDevice(device));

    fqsb_system* s = new fqsb_system();
    s->par = *par;
    s->device = device;
    s->own_stream = true;
    s->stream = nullptr;
    s->forces_valid = false;
    s->forces_frozen = false;
    s->d_scratch = nullptr;
    s->scratch_bytes = 0;
    s->launches = 0;
    s->steps = 0;
    s->last_kernel = "";
    s->ev0 = s->ev1 = nullptr;
    s->lr_gemm = false;
    s->slab = nullptr;
    s->d_mark = nullptr;
    for (int k = 0; k < 4; ++k) {
        s->pipe_stream[k] = nullptr;
    }
    s->pipe_ev = nullptr;
    s->own_lo = 0;
    s->own_hi = 0;
    for (int k = 0; k < 5; ++k) {
        s->snap_d[k] = nullptr;
    }
    s->snap_idx = nullptr;
    s->snap_rng = nullptr;
    s->snap_uf = nullptr;
    s->snap_ctl = nullptr;
    s->snap_valid = false;
    s->d_log = nullptr;
    s->log_cap = 0;
    memset(&s->bk, 0, sizeof s->bk);
    s->bk_ready = false;
    s->bk_log_tiles = 0;
    s->d_lr_tab = s->d_lr_w = s->d_lr_y = nullptr;
    s->lr_rowsum = 0.0;
    s->kernel_ms = 0.0;
    s->kernel_launches = 0;
    s->thermal = false;
    memset(&s->th, 0, sizeof s->th);
    memset(&s->F, 0, sizeof s->F);
    memset(&s->S, 0, sizeof s->S);

    Par& P = s->P;
    memset(&P, 0, sizeof P);
    P.pot = par->potential;
    P.inter = par->interactions;
    P.rank = par->rank;
    P.dist = par->distribution;
    P.rows = (int)par->shape[0];
    P.cols = par->rank == 2 ? (int)par->shape[1] : 1;
    P.N = (i64)P.rows * P.cols;
    P.R = par->nrealisations > 0 ? par->nrealisations : 1;
    P.consumes = par->distribution != FQSB_DIST_DELTA;
    P.fma = (par->kernel & FQSB_KERNEL_FMA) ? 1 : 0;
    P.m = par->m;
    P.inv_m = 1.0 / par->m; // detail.h:1110
    P.eta = par->eta;
    P.mu = par->mu;
    P.kappa = par->kappa;
    P.k1 = par->k1;
    P.k2 = par->k2;
    P.k_frame = par->k_frame;
    P.dt = par->dt;
    P.c2 = 0.5 * par->dt * par->dt;
    {
        // prrng defaults for omitted parameters
        double def[4] = {1.0, 0.0, 0.0, 0.0};
        if (par->distribution == FQSB_DIST_PARETO || par->distribution == FQSB_DIST_WEIBULL ||
            par->distribution == FQSB_DIST_GAMMA) {
            def[1] = 1.0;
        }
        if (par->distribution == FQSB_DIST_NORMAL) { // normal(mu = 0, sigma = 1), offset 0
            def[0] = 0.0;
            def[1] = 1.0;
        }
        for (int k = 0; k < 4; ++k) {
            P.dpar[k] = k < par->nparameters ? par->parameters[k] : def[k];
        }
    }
    P.offset = par->offset;
    P.seed = par->seed;
    P.seed_stride = par->seed_stride > 0 ? (u64)par->seed_stride : (u64)P.N;
    P.seed_first = par->seed_period > 0 ? (u64)par->seed_first : 0ULL;
    P.seed_period = par->seed_period > 0 ? (u64)par->seed_period : 0ULL;
    s->N = P.N;
    s->own_hi = P.N;
    s->R = P.R;
    s->n = P.N * P.R;
    s->par.nrealisations = P.R;
    s->par.seed_stride = (int64_t)P.seed_stride;
    s->par.device = device;
    if (P.N >= (i64)1 << 31 || P.R >= 65535) {
        delete s;
        return fail(FQSB_EASSERT, ASSERT_MSG("size < 2^31 && nrealisations < 65535"));
    }

    int rc = FQSB_OK;
    auto build = [&]() -> int {
        CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&s->ev0));
        CU(cudaEventCreate(&s->ev1));
        State& S = s->S;
        const size_t n = (size_t)s->n;
        TRY(dev_alloc(s, &S.u, n));
        TRY(dev_alloc(s, &S.v, n));
        TRY(dev_alloc(s, &S.a, n));
        TRY(dev_alloc(s, &S.yl, n));
        TRY(dev_alloc(s, &S.yr, n));
        TRY(dev_alloc(s, &S.idx, n));
        TRY(dev_alloc(s, &S.rng, n));
        TRY(dev_alloc(s, &S.u_frame, (size_t)s->R));
        TRY(dev_alloc(s, &S.ctl, (size_t)s->R));
        TRY(dev_alloc(s, &S.err, 2));
        S.tiles = (int)grid_for(s->N);
        {
            // the tiled 1-D step uses one CTA per FQSB_ST_TILE blocks: the partial-sum buffer
            // must cover it for lines beyond 148 * 16 tiles (N > 4.8e6 with the streaming kernel)
            const i64 tiled = (s->N + FQSB_ST_TILE - 1) / FQSB_ST_TILE;
            if (P.rank == 1 && tiled > S.tiles) {
                S.tiles = (int)tiled;
            }
        }
        if (P.rank == 2) {
            // k_stream_2d: 2 CTAs per SM (126 registers), 2 halo rows of u,v,a = 0.75 rows of
            // traffic per band; k_stream_np_2d: FQSB_S2_NP_CTAS CTAs per SM, 2 halo rows of u = 0.5 rows
            int sms = 148;
            CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
            P.s2_ty = plan_band_rows(P, 2 * sms, 0.75, S.tiles);
            P.s2_ty_np = plan_band_rows(P, FQSB_S2_NP_CTAS * sms, 0.5, S.tiles);
            // tuning knobs (tools/line2d.py): rows per CTA forced from the environment
            if (const char* e = std::getenv("FQSB_S2_TY")) {
                const int ty = std::atoi(e);
                if (ty >= 1 && ((P.cols + FQSB_S2_TX - 1) / FQSB_S2_TX) * ((P.rows + ty - 1) / ty) <= S.tiles) {
                    P.s2_ty = ty;
                }
            }
            if (const char* e = std::getenv("FQSB_S2_TY_NP")) {
                const int ty = std::atoi(e);
                if (ty >= 1 && ((P.cols + FQSB_S2_TX - 1) / FQSB_S2_TX) * ((P.rows + ty - 1) / ty) <= S.tiles) {
                    P.s2_ty_np = ty;
                }
            }
        }
        TRY(dev_alloc(s, &s->d_red, (size_t)s->R * S.tiles * FQSB_NPART));
        S.part = s->d_red;
        TRY(dev_alloc(s, &s->d_out, (size_t)s->R * 4));
        TRY(dev_alloc(s, &s->d_du, (size_t)s->R));
        s->d_in = nullptr;
        s->d_pref = nullptr;
        CU(cudaHostAlloc((void**)&s->h_out, (size_t)s->R * 4 * sizeof(double), cudaHostAllocDefault));
        CU(cudaHostAlloc((void**)&s->h_ctl, (size_t)s->R * sizeof(Ctl), cudaHostAllocDefault));
        CU(cudaHostAlloc((void**)&s->h_err, 2 * sizeof(int), cudaHostAllocDefault));
        CU(cudaMemsetAsync(S.err, 0, 2 * sizeof(int), s->stream));
        CU(cudaMemsetAsync(S.u_frame, 0, (size_t)s->R * sizeof(double), s->stream));
        CU(cudaMemsetAsync(S.ctl, 0, (size_t)s->R * sizeof(Ctl), s->stream));
        if (P.inter == INT_LONGRANGE1D) { // detail.h:829-844 (host pow, as the reference)
            std::vector<double> pref((size_t)P.N, 0.0);
            for (i64 d = 1; d < P.N; ++d) {
                pref[(size_t)d] = P.k1 / std::pow((double)d, P.k2 + 1.0);
            }
            TRY(dev_alloc(s, &s->d_pref, (size_t)P.N));
            CU(cudaMemcpyAsync(s->d_pref, pref.data(), (size_t)P.N * sizeof(double),
                               cudaMemcpyHostToDevice, s->stream));
            CU(cudaStreamSynchronize(s->stream));
            S.pref = s->d_pref;
            // circulant table tab[k] = pref[min(k, N-k)] (detail.h:859-862) and its row sum
            // chosen for lines beyond the resident kernel, when streaming is forced, and -- in
            // auto mode -- for ensembles, where the exact-order O(N^2) sum inside the resident
            // kernel (bit-identical to the reference's loop) would leave the tensor cores idle
            const bool stream_path = (par->kernel & 15) == 2 || resident_cfg(P.N).B == 0 ||
                                     ((par->kernel & 15) == 0 && P.R >= 32);
            if (stream_path && P.N % 2 == 0 && P.N >= 256 && lr_gemm_smem(P.N) <= 227 * 1024 &&
                P.dist != DIST_GAMMA && P.dist != DIST_NORMAL) {
                std::vector<double> tab((size_t)P.N, 0.0);
                double rowsum = 0.0;
                for (i64 k = 1; k < P.N; ++k) {
                    i64 d = k < P.N - k ? k : P.N - k;
                    tab[(size_t)k] = pref[(size_t)d];
                    rowsum += tab[(size_t)k];
                }
                TRY(dev_alloc(s, &s->d_lr_tab, (size_t)P.N));
                TRY(dev_alloc(s, &s->d_lr_w, n));
                TRY(dev_alloc(s, &s->d_lr_y, n));
                CU(cudaMemcpyAsync(s->d_lr_tab, tab.data(), (size_t)P.N * sizeof(double),
                                   cudaMemcpyHostToDevice, s->stream));
                CU(cudaStreamSynchronize(s->stream));
                s->lr_rowsum = rowsum;
                s->lr_gemm = true;
            }
        }
        k_init<<<grid_for(s->n), 256, 0, s->stream>>>(P, S);
        CU(cudaGetLastError());
        s->launches++;
        return check_flags(s);
    };
    rc = build();
    if (rc != FQSB_OK) {
        std::string keep = g_err;
        fqsb_destroy(s);
        if (rc == FQSB_EASSERT && keep.find("yield landscape") != std::string::npos) {
            keep = "u = 0 lies below the first yield position: lower the offset";
        }
        g_err = keep;
        return rc;
    }
    *out = s;
    return FQSB_OK;
}

void fqsb_destroy(fqsb_system* s)
{
    if (!s) {
        return;
    }
    cudaSetDevice(s->device);
    if (s->stream) {
        cudaStreamSynchronize(s->stream);
    }
    slab_free(s);
    for (int k = 0; k < 4; ++k) {
        if (s->pipe_stream[k]) {
            cudaStreamDestroy(s->pipe_stream[k]);
        }
    }
    if (s->pipe_ev) {
        cudaEventDestroy(s->pipe_ev);
    }
    for (void* p : s->allocs) {
        cudaFree(p);
    }
    if (s->d_scratch) {
        cudaFree(s->d_scratch);
    }
    if (s->h_out) {
        cudaFreeHost(s->h_out);
    }
    if (s->h_ctl) {
        cudaFreeHost(s->h_ctl);
    }
    if (s->h_err) {
        cudaFreeHost(s->h_err);
    }
    if (s->own_stream && s->stream) {
        cudaStreamDestroy(s->stream);
    }
    if (s->ev0) {
        cudaEventDestroy(s->ev0);
    }
    if (s->ev1) {
        cudaEventDestroy(s->ev1);
    }
    cudaGetLastError();
    delete s;
}

int fqsb_get_params(const fqsb_system* s, fqsb_params* out)
{
    if (!s || !out) {
        return fail(FQSB_EASSERT, "null argument");
    }
    *out = s->par;
    return FQSB_OK;
}

int64_t fqsb_size(const fqsb_system* s) { return s ? s->N : 0; }
int64_t fqsb_nrealisations(const fqsb_system* s) { return s ? s->R : 0; }

int fqsb_set_stream(fqsb_system* s, void* cuda_stream)
{
    TRY(enter(s));
    CU(cudaStreamSynchronize(s->stream));
    if (s->own_stream && s->stream) {
        CU(cudaStreamDestroy(s->stream));
    }
    s->stream = (cudaStream_t)cuda_stream;
    s->own_stream = false;
    return FQSB_OK;
}

void* fqsb_get_stream(const fqsb_system* s) { return s ? (void*)s->stream : nullptr; }
int64_t fqsb_launch_count(const fqsb_system* s) { return s ? s->launches : 0; }
int64_t fqsb_step_count(const fqsb_system* s) { return s ? s->steps : 0; }
const char* fqsb_last_kernel(const fqsb_system* s) { return s ? s->last_kernel : ""; }
double fqsb_last_kernel_seconds(const fqsb_system* s) { return s ? 1e-3 * s->kernel_ms : 0.0; }
int64_t fqsb_last_kernel_launches(const fqsb_system* s) { return s ? s->kernel_launches : 0; }

// ---- state in ---------------------------------------------------------------------------------
static int upload(fqsb_system* s, double* dst, const double* src, int64_t n)
{
    if (n != s->n) {
        return fail(FQSB_EASSERT, ASSERT_MSG("xt::has_shape(arg, m_u.shape())")); // detail.h:1278
    }
    CU(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    return FQSB_OK;
}

int fqsb_set_u(fqsb_system* s, const double* u, int64_t n)
{
    TRY(enter(s));
    TRY(upload(s, s->S.u, u, n));
    TRY(align(s, nullptr)); // updated_u(), detail.h:1280
    invalidate_forces(s);
    return check_flags(s);
}

int fqsb_set_v(fqsb_system* s, const double* v, int64_t n)
{
    TRY(enter(s));
    TRY(upload(s, s->S.v, v, n));
    if (s->forces_frozen) {
        TRY(frozen_update(s, 8)); // updated_v(), detail.h:1294
    }
    else {
        s->forces_valid = false;
    }
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_set_a(fqsb_system* s, const double* a, int64_t n)
{
    TRY(enter(s));
    TRY(upload(s, s->S.a, a, n));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_set_u_frame(fqsb_system* s, const double* u_frame)
{
    TRY(enter(s));
    CU(cudaMemcpyAsync(s->S.u_frame, u_frame, (size_t)s->R * sizeof(double),
                       cudaMemcpyHostToDevice, s->stream));
    if (s->forces_frozen) {
        TRY(frozen_update(s, 4)); // computeForceFrame + computeForce, detail.h:1256-1257
    }
    else {
        s->forces_valid = false;
    }
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_set_inc(fqsb_system* s, const int64_t* inc)
{
    TRY(enter(s));
    TRY(pull_ctl(s));
    for (i64 r = 0; r < s->R; ++r) { // detail.h:1241-1247
        s->h_ctl[r].inc = inc[r];
        s->h_ctl[r].qs_first = inc[r];
        s->h_ctl[r].qs_last = inc[r];
    }
    TRY(push_ctl(s));
    if (s->thermal) { // updated_inc(), detail.h:1246
        TRY(thermal_update(s, 0, 0));
        if (!s->forces_frozen) {
            s->forces_valid = false;
        }
        CU(cudaStreamSynchronize(s->stream));
    }
    return FQSB_OK;
}

int fqsb_set_t(fqsb_system* s, const double* t)
{
    TRY(enter(s));
    TRY(pull_ctl(s));
    for (i64 r = 0; r < s->R; ++r) { // detail.h:1231-1235 (quirk Q5: qs_* untouched)
        i64 inc = (i64)std::round(t[r] / s->P.dt);
        double tt = (double)inc * s->P.dt;
        if (!(std::fabs(tt - t[r]) <= 1e-8 + 1e-5 * std::fabs(t[r]))) {
            return fail(FQSB_EASSERT, ASSERT_MSG("xt::allclose(this->t(), arg)"));
        }
        s->h_ctl[r].inc = inc;
    }
    return push_ctl(s);
}

int fqsb_refresh(fqsb_system* s)
{
    TRY(enter(s));
    TRY(align(s, nullptr));
    TRY(thermal_update(s, 0, 0)); // updated_inc(), detail.h:1314
    invalidate_forces(s);
    return check_flags(s);
}

int fqsb_quench(fqsb_system* s)
{
    TRY(enter(s));
    k_zero_va<<<grid_for(s->n), 256, 0, s->stream>>>(s->P, s->S);
    CU(cudaGetLastError());
    s->launches++;
    if (s->forces_frozen) {
        TRY(frozen_update(s, 8));
    }
    else {
        s->forces_valid = false;
    }
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

// ---- state out --------------------------------------------------------------------------------
static int array_ptr(fqsb_system* s, int which, const double** p)
{
    switch (which) {
    case FQSB_U:
        *p = s->S.u;
        return FQSB_OK;
    case FQSB_V:
        *p = s->S.v;
        return FQSB_OK;
    case FQSB_A:
        *p = s->S.a;
        return FQSB_OK;
    default:
        break;
    }
    if (which < FQSB_F || which > FQSB_F_DAMPING) {
        return fail(FQSB_EASSERT, "unknown array id");
    }
    TRY(ensure_forces(s));
    const double* f[] = {s->F.f, s->F.f_pot, s->F.f_frame, s->F.f_int, s->F.f_damp};
    *p = f[which - FQSB_F];
    return FQSB_OK;
}

int fqsb_get(fqsb_system* s, int which, double* out, int64_t n)
{
    TRY(enter(s));
    if (n != s->n) {
        return fail(FQSB_EASSERT, ASSERT_MSG("size of the output buffer"));
    }
    const double* src = nullptr;
    TRY(array_ptr(s, which, &src));
    CU(cudaMemcpyAsync(out, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_get_device(fqsb_system* s, int which, const double** out_device)
{
    TRY(enter(s));
    TRY(array_ptr(s, which, out_device));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_get_u_frame(fqsb_system* s, double* out)
{
    TRY(enter(s));
    CU(cudaMemcpyAsync(out, s->S.u_frame, (size_t)s->R * sizeof(double), cudaMemcpyDeviceToHost,
                       s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_get_inc(fqsb_system* s, int64_t* out)
{
    TRY(enter(s));
    TRY(pull_ctl(s));
    for (i64 r = 0; r < s->R; ++r) {
        out[r] = s->h_ctl[r].inc;
    }
    return FQSB_OK;
}

int fqsb_get_t(fqsb_system* s, double* out)
{
    TRY(enter(s));
    TRY(pull_ctl(s));
    for (i64 r = 0; r < s->R; ++r) {
        out[r] = (double)s->h_ctl[r].inc * s->P.dt; // detail.h:1479
    }
    return FQSB_OK;
}

int fqsb_qs_activity(fqsb_system* s, int64_t* first, int64_t* last)
{
    TRY(enter(s));
    TRY(pull_ctl(s));
    for (i64 r = 0; r < s->R; ++r) {
        if (first) {
            first[r] = s->h_ctl[r].qs_first;
        }
        if (last) {
            last[r] = s->h_ctl[r].qs_last;
        }
    }
    return FQSB_OK;
}

int fqsb_residual(fqsb_system* s, double* out)
{
    TRY(enter(s));
    // frozen forces are reduced as stored; otherwise they are derived on the fly
    if (s->lr_gemm && !s->forces_frozen) {
        TRY(ensure_forces(s));
    }
    TRY(reduce_to_host(s, (s->forces_frozen || s->lr_gemm) ? 0 : 1, 1, nullptr));
    for (i64 r = 0; r < s->R; ++r) { // detail.h:1512-1520
        double r_fres = std::sqrt(s->h_out[4 * r]);
        double r_fext = std::sqrt(s->h_out[4 * r + 1]);
        out[r] = r_fext != 0.0 ? r_fres / r_fext : r_fres;
    }
    return FQSB_OK;
}

int fqsb_temperature(fqsb_system* s, double* out)
{
    TRY(enter(s));
    TRY(reduce_to_host(s, 2, 1, nullptr));
    for (i64 r = 0; r < s->R; ++r) { // detail.h:1502
        out[r] = 0.5 * s->P.m * s->h_out[4 * r] / (double)s->N;
    }
    return FQSB_OK;
}

int fqsb_mean_f_frame(fqsb_system* s, double* out)
{
    TRY(enter(s));
    TRY(reduce_to_host(s, 2, 1, nullptr));
    for (i64 r = 0; r < s->R; ++r) {
        out[r] = s->h_out[4 * r + 1] / (double)s->N;
    }
    return FQSB_OK;
}

// ---- dynamics ---------------------------------------------------------------------------------
__global__ void k_ctl_begin(const Par P, const State S, int track_user, int overdamped,
                            const double* red_out)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.R) {
        return;
    }
    Ctl& c = S.ctl[r];
    c.status = ST_RUNNING;
    c.init = 1;
    c.steps = 0;
    c.S = track_user ? (i64)red_out[4 * r + 2] : 0;
    c.A = track_user ? (i64)red_out[4 * r + 1] : 0;
    c.s_n = 0;
    c.count = 0u;
    c.flip = 0;
    c.residual = 0.0;
    for (int k = 0; k < FQSB_RING; ++k) {
        c.ring[k] = __longlong_as_double(0x7ff0000000000000LL); // +inf (App. A.4)
        c.ring_den[k] = 1.0;
    }
    if (overdamped) { // detail.h:1704-1705
        c.qs_first = c.inc;
        c.qs_last = c.inc;
    }
}

// `gamma` / `normal` landscapes only exist in the generic streaming kernels (fqsb_slowdist.cu)
static bool slow_distribution(const fqsb_system* s)
{
    return s->P.dist == DIST_GAMMA || s->P.dist == DIST_NORMAL;
}

static bool use_resident(const fqsb_system* s, ResidentCfg* cfg, int mode)
{
    *cfg = resident_cfg(s->N, (s->par.kernel >> 4) & 15);
    if (slow_distribution(s)) {
        return false;
    }
    if (mode == MODE_LOG || s->own_lo != 0 || s->own_hi != s->N) {
        return false; // slab batches run on the streaming kernels
    }
    if (s->lr_gemm || s->thermal) {
        return false; // LongRange through the tensor-core GEMM (K7); thermal: streaming kernel
    }
    if ((s->par.kernel & 15) == 2 || cfg->B == 0) {
        return false;
    }
    return resident_smem(s->P, *cfg) <= 227 * 1024;
}

// K2t: fixed-step calls (timeSteps / flowSteps) of a thermal system whose line fits one CTA
static bool use_resident_thermal(const fqsb_system* s, ResidentCfg* cfg, int mode)
{
    if (!s->thermal || mode != MODE_FIXED || (s->par.kernel & 15) == 2 || s->own_lo != 0 ||
        s->own_hi != s->N || slow_distribution(s)) {
        return false;
    }
    *cfg = resident_cfg(s->N, (s->par.kernel >> 4) & 15);
    return cfg->B != 0 && resident_thermal_smem(s->P, *cfg) <= 227 * 1024;
}

// K2b: 1-D nearest-neighbour lines beyond one CTA take the temporally blocked kernel
// (par.kernel: low 4 bits 0 = auto, 3 = force; bits 8..15 steps per launch, bits 16..31 owned
// blocks per tile -- both 0 = planner's choice)
static bool use_blocked(const fqsb_system* s, int mode, bool overdamped)
{
    const int sel = s->par.kernel & 15;
    if (overdamped || mode == MODE_LOG || s->own_lo != 0 || s->own_hi != s->N || s->lr_gemm ||
        s->thermal || slow_distribution(s)) {
        return false;
    }
    if (!blocked_supported(s->P)) {
        return false;
    }
    return sel == 3 || (sel == 0 && s->N > kResidentMaxN);
}

static int ensure_stream_buffers(fqsb_system* s);
static int run_host_ring(fqsb_system* s, RunArgs A, bool overdamped, bool track_user);

// host-only view of the tile planner (tests, tools)
extern "C" int fqsb_plan_blocked(int64_t n_blocks, int64_t n_realisations, int has_interactions,
                                 int steps_per_launch, int own_hint, int64_t halo_cells,
                                 int64_t* out)
{
    if (!out || n_blocks < 2 || n_realisations < 1) {
        return fail(FQSB_EASSERT, "fqsb_plan_blocked: n_blocks >= 2, n_realisations >= 1, out != NULL");
    }
    Par P;
    memset(&P, 0, sizeof P);
    P.rank = 1;
    P.N = n_blocks;
    P.R = n_realisations;
    P.pot = POT_CUSPY;
    P.inter = has_interactions ? INT_LAPLACE1D : INT_NONE;
    const BlockedPlan pl = blocked_plan(P, steps_per_launch, own_hint);
    if (pl.B < 2 || pl.B > 8) {
        return fail(FQSB_EUNSUPPORTED, "no tile geometry for the blocked kernel");
    }
    int sms = 148, dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
        sms = n;
    }
    cudaGetLastError(); // (no device: the planner falls back to 148 SMs, nothing to report)
    out[0] = pl.B;
    out[1] = pl.own;
    out[2] = pl.H;
    out[3] = pl.ksteps;
    out[4] = pl.ntiles;
    out[5] = ((int64_t)pl.ntiles * n_realisations + sms - 1) / sms;
    out[6] = out[7] = 0;
    if (halo_cells > 0) {
        for (int c = 0; c < pl.ntiles; ++c) {
            bool rd, pu;
            blocked_tile_roles(n_blocks, halo_cells, pl.own, pl.H, c, &rd, &pu);
            out[6] += rd;
            out[7] += pu;
        }
    }
    return FQSB_OK;
}

static int ensure_blocked_buffers(fqsb_system* s, const BlockedPlan& plan)
{
    TRY(ensure_stream_buffers(s));
    if (!s->bk_ready) {
        TRY(dev_alloc(s, &s->bk.yl2, (size_t)s->n));
        TRY(dev_alloc(s, &s->bk.yr2, (size_t)s->n));
        TRY(dev_alloc(s, &s->bk.idx2, (size_t)s->n));
        TRY(dev_alloc(s, &s->bk.rng2, (size_t)s->n));
        TRY(dev_alloc(s, &s->bk.uf2, (size_t)s->R));
        s->bk_ready = true;
    }
    if (plan.ntiles > s->bk_log_tiles) { // (the geometry depends on the batch length)
        const size_t ngroups = ((size_t)plan.ntiles + FQSB_BK_GROUP - 1) / FQSB_BK_GROUP;
        CU(cudaStreamSynchronize(s->stream)); // (no launch may still count in the old buffer)
        TRY(dev_alloc(s, &s->bk.log,
                      (size_t)s->R * FQSB_BK_MAXSTEPS * (size_t)plan.ntiles * FQSB_NLOG));
        TRY(dev_alloc(s, &s->bk.glog, (size_t)s->R * FQSB_BK_MAXSTEPS * ngroups * FQSB_NLOG));
        TRY(dev_alloc(s, &s->bk.gcount, (size_t)s->R * ngroups));
        CU(cudaMemsetAsync(s->bk.gcount, 0, (size_t)s->R * ngroups * sizeof(unsigned int),
                           s->stream));
        s->bk_log_tiles = plan.ntiles;
    }
    s->bk.own = plan.own;
    s->bk.H = plan.H;
    s->bk.ksteps = plan.ksteps;
    s->bk.ntiles = plan.ntiles;
    return FQSB_OK;
}

// one dynamics call on the blocked kernel: batches of up to plan.ksteps steps per launch
static int run_blocked(fqsb_system* s, RunArgs A)
{
    // steps per launch: long batches amortise the pass over memory (measured on one line of
    // 2^20 blocks: 4.76 / 4.44 / 3.66 us per step at 16 / 32 / 64 steps per launch); a stop
    // inside a batch costs a replay of the batch up to that step, so timeStepsUntilEvent --
    // which stops within a few steps while an avalanche runs -- takes short ones
    int ksteps = (s->par.kernel >> 8) & 255;
    if (ksteps == 0) {
        ksteps = A.mode == MODE_UNTIL_EVENT ? 8 : FQSB_BK_MAXSTEPS;
    }
    const BlockedPlan plan = blocked_plan(s->P, ksteps, (s->par.kernel >> 16) & 0xffff);
    if (plan.B < 2 || plan.B > 8 || blocked_smem(plan.B) > 227 * 1024) {
        return fail(FQSB_EUNSUPPORTED, "no tile geometry for the blocked kernel");
    }
    TRY(ensure_blocked_buffers(s, plan));
    s->last_kernel = "blocked_1d";
    BlockedArgs K = s->bk;
    if (A.mode == MODE_FIXED) {
        CU(cudaEventRecord(s->ev0, s->stream));
        i64 done = 0;
        int flip = 0;
        while (done < A.max_steps) {
            const i64 left = A.max_steps - done;
            K.nsteps = (int)(left < plan.ksteps ? left : plan.ksteps);
            K.flip = flip;
            cudaError_t e = launch_blocked(plan, s->P, s->S, A, K, s->stream);
            if (e != cudaSuccess) {
                return cuda_fail(e, "blocked kernel launch");
            }
            done += K.nsteps;
            flip ^= 1;
            s->launches++;
            s->kernel_launches++;
        }
        CU(cudaEventRecord(s->ev1, s->stream));
        CU(launch_blocked_fixed_done(s->P, s->S, A.max_steps, flip, s->stream));
        s->launches++;
    }
    else {
        CU(launch_blocked_begin(s->P, s->S, plan.ksteps, A.max_steps, s->stream));
        s->launches++;
        i64 nb = 4;
        for (;;) {
            CU(cudaEventRecord(s->ev0, s->stream));
            for (i64 b = 0; b < nb; ++b) {
                cudaError_t e = launch_blocked(plan, s->P, s->S, A, K, s->stream);
                if (e != cudaSuccess) {
                    return cuda_fail(e, "blocked kernel launch");
                }
            }
            CU(cudaEventRecord(s->ev1, s->stream));
            s->launches += nb;
            s->kernel_launches += nb;
            TRY(pull_ctl(s));
            {
                float ms = 0.f;
                CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
                s->kernel_ms += ms;
            }
            bool running = false;
            for (i64 r = 0; r < s->R; ++r) {
                running |= s->h_ctl[r].status == ST_RUNNING;
            }
            if (!running) {
                break;
            }
            if (nb < 64) {
                nb *= 2;
            }
        }
    }
    CU(launch_blocked_settle(s->P, s->S, s->bk, s->stream));
    {
        const unsigned rg = (unsigned)((s->R + 127) / 128);
        k_stream_settle_flags<<<rg, 128, 0, s->stream>>>(s->P, s->S);
        CU(cudaGetLastError());
    }
    s->launches += 2;
    if (A.mode == MODE_FIXED) {
        TRY(pull_ctl(s));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        s->kernel_ms += ms;
    }
    return FQSB_OK;
}

static int ensure_stream_buffers(fqsb_system* s)
{
    if (s->lr_gemm) {
        return FQSB_OK; // the LongRange GEMM path updates in place (no neighbour reads)
    }
    if (!s->S.u2) {
        TRY(dev_alloc(s, &s->S.u2, (size_t)s->n));
        if (s->par.minimisation != FQSB_MIN_OVERDAMPED) {
            TRY(dev_alloc(s, &s->S.v2, (size_t)s->n));
            TRY(dev_alloc(s, &s->S.a2, (size_t)s->n));
        }
    }
    return FQSB_OK;
}

// Runs one dynamics call to completion. On return h_ctl holds the final control blocks.
static int run(fqsb_system* s, RunArgs A, bool overdamped, bool track_user)
{
    if (A.mode != MODE_FIXED && A.mode != MODE_LOG && A.niter_tol < 1) {
        return fail(FQSB_EASSERT, ASSERT_MSG("niter_tol >= 1"));
    }
    if (A.mode != MODE_FIXED && A.mode != MODE_LOG && A.niter_tol > FQSB_RING) {
        // the device keeps the StopList in the 32 lanes of a warp; longer lists (the reference
        // takes any size, detail.h:1676-1689) are replayed on the host over logged batches
        if (s->R != 1) {
            return fail(FQSB_EUNSUPPORTED,
                        "niter_tol > 32 is available for single systems (nrealisations == 1)");
        }
        return run_host_ring(s, A, overdamped, track_user);
    }
    A.own_lo = (int)s->own_lo;
    A.own_hi = (int)s->own_hi;
    const unsigned rg = (unsigned)((s->R + 127) / 128);
    k_ctl_begin<<<rg, 128, 0, s->stream>>>(s->P, s->S, track_user ? 1 : 0, overdamped ? 1 : 0,
                                           s->d_out);
    CU(cudaGetLastError());
    s->launches++;
    s->kernel_ms = 0.0;
    s->kernel_launches = 0;
    invalidate_forces(s);
    if (A.max_steps <= 0) {
        TRY(pull_ctl(s));
        for (i64 r = 0; r < s->R; ++r) {
            s->h_ctl[r].status = ST_EXHAUSTED;
        }
        return FQSB_OK;
    }

    ResidentCfg cfg;
    const int sel = s->par.kernel & 15;
    if (sel != 1 && sel != 2 && use_blocked(s, A.mode, overdamped)) {
        TRY(run_blocked(s, A));
    }
    else if (use_resident(s, &cfg, A.mode) || use_resident_thermal(s, &cfg, A.mode)) {
        const bool thermal = s->thermal;
        s->last_kernel = thermal ? "resident_thermal"
                                 : (overdamped ? "resident_nopassing" : "resident");
        const i64 chunk = (i64)1 << 20;
        for (;;) {
            A.launch_steps = A.max_steps < chunk ? A.max_steps : chunk;
            CU(cudaEventRecord(s->ev0, s->stream));
            cudaError_t e =
                thermal ? launch_resident_thermal(cfg, s->P, s->S, A, s->th, s->stream)
                        : (overdamped ? launch_resident_nopassing(cfg, s->P, s->S, A, s->stream)
                                      : launch_resident(cfg, s->P, s->S, A, s->stream));
            if (e != cudaSuccess) {
                return cuda_fail(e, "resident kernel launch");
            }
            CU(cudaEventRecord(s->ev1, s->stream));
            s->launches++;
            s->kernel_launches++;
            TRY(pull_ctl(s));
            {
                float ms = 0.f;
                CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
                s->kernel_ms += ms;
            }
            bool running = false;
            for (i64 r = 0; r < s->R; ++r) {
                running |= s->h_ctl[r].status == ST_RUNNING;
            }
            if (!running) {
                break;
            }
        }
    }
    else {
        if ((s->par.kernel & 15) == 1 && !slow_distribution(s)) {
            return fail(FQSB_EUNSUPPORTED, "system too large for the resident kernel");
        }
        TRY(ensure_stream_buffers(s));
        s->last_kernel = overdamped ? "stream_nopassing"
                                    : (s->lr_gemm ? "stream_longrange_dmma" : stream_step_name(s->P));
        // fixed-step calls without a moving frame need no per-step decision on the device
        // (thermal systems need Ctl::inc up to date ahead of every step)
        const int finalise = (A.mode != MODE_FIXED || A.flow || s->thermal) ? 1 : 0;
        // upper bound on launches still useful (no-passing: launch l decides sweep l-1)
        i64 remaining = overdamped ? A.max_steps + 1 : A.max_steps;
        i64 launched = 0;
        i64 batch = 16;
        for (;;) {
            i64 nb = A.mode == MODE_FIXED ? (remaining < 2048 ? remaining : 2048)
                                          : (remaining < batch ? remaining : batch);
            CU(cudaEventRecord(s->ev0, s->stream));
            for (i64 b = 0; b < nb; ++b) {
                if (s->thermal) { // m_inc++; updated_inc(): detail.h:1541-1544
                    TRY(thermal_update(s, 1, 1));
                }
                cudaError_t e =
                    overdamped ? launch_stream_sweep(s->P, s->S, A, s->stream,
                                                     (int)((launched + b) & 1), launched + b == 0,
                                                     launched + b < A.max_steps)
                    : s->lr_gemm
                        ? launch_lr_step(s->P, s->S, A, s->d_lr_tab, s->lr_rowsum, s->d_lr_w,
                                         s->d_lr_y, s->stream, finalise)
                        : launch_stream_step(s->P, s->S, A, s->stream, (int)((launched + b) & 1),
                                             finalise);
                if (e != cudaSuccess) {
                    return cuda_fail(e, "stream kernel launch");
                }
            }
            CU(cudaEventRecord(s->ev1, s->stream));
            launched += nb;
            s->launches += s->lr_gemm ? 3 * nb : nb;
            s->kernel_launches += s->lr_gemm ? 3 * nb : nb;
            remaining -= nb;
            if (!finalise && !overdamped && remaining <= 0) {
                k_stream_fixed_done<<<rg, 128, 0, s->stream>>>(s->P, s->S, A.max_steps,
                                                               s->lr_gemm ? 1 : 0);
                CU(cudaGetLastError());
                s->launches++;
            }
            TRY(pull_ctl(s));
            {
                float ms = 0.f;
                CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
                s->kernel_ms += ms;
            }
            bool running = false;
            for (i64 r = 0; r < s->R; ++r) {
                running |= s->h_ctl[r].status == ST_RUNNING;
            }
            if (!running || remaining <= 0) {
                break;
            }
            if (batch < 256) {
                batch *= 2;
            }
        }
        dim3 grid((unsigned)s->S.tiles, (unsigned)s->R);
        k_stream_settle<<<grid, 256, 0, s->stream>>>(s->P, s->S, overdamped ? 0 : 1);
        k_stream_settle_flags<<<rg, 128, 0, s->stream>>>(s->P, s->S);
        CU(cudaGetLastError());
        s->launches += 2;
        if (overdamped) {
            // wells moved by the discarded look-ahead sweep follow the kept configuration again
            TRY(align(s, nullptr));
        }
    }
    for (i64 r = 0; r < s->R; ++r) {
        s->steps += s->h_ctl[r].steps;
    }
    TRY(check_flags(s));
    for (i64 r = 0; r < s->R; ++r) {
        if (s->h_ctl[r].status == ST_NAN) {
            return fail(FQSB_ENAN, "NaN entries found");
        }
    }
    return FQSB_OK;
}

static void fill_ret(const fqsb_system* s, int64_t* ret)
{
    if (!ret) {
        return;
    }
    for (i64 r = 0; r < s->R; ++r) {
        const Ctl& c = s->h_ctl[r];
        switch (c.status) {
        case ST_CONVERGED:
            ret[r] = 0;
            break;
        case ST_EVENT:
        case ST_TRUNCATED:
            ret[r] = c.steps;
            break;
        default:
            ret[r] = c.steps + 1; // detail.h:1621,1791 (quirk Q4)
            break;
        }
    }
}

static int require_dynamic(const fqsb_system* s)
{
    if (s->par.minimisation == FQSB_MIN_OVERDAMPED) {
        // Line1d.h:231-234: the no-passing system hides the dynamics
        return fail(FQSB_EUNSUPPORTED, "no dynamics available for an overdamped system");
    }
    return FQSB_OK;
}

static RunArgs make_args(int mode, i64 max_steps)
{
    RunArgs A;
    memset(&A, 0, sizeof A);
    A.mode = mode;
    A.max_steps = max_steps;
    A.niter_tol = 1;
    return A;
}

int fqsb_time_steps(fqsb_system* s, int64_t n)
{
    TRY(enter(s));
    TRY(require_dynamic(s));
    if (n < 0) {
        return fail(FQSB_EASSERT, ASSERT_MSG("n + 1 < std::numeric_limits<long>::max()"));
    }
    return run(s, make_args(MODE_FIXED, n), false, false);
}

int fqsb_flow_steps(fqsb_system* s, int64_t n, double v_frame)
{
    TRY(enter(s));
    TRY(require_dynamic(s));
    if (n < 0) {
        return fail(FQSB_EASSERT, ASSERT_MSG("n + 1 < std::numeric_limits<long>::max()"));
    }
    RunArgs A = make_args(MODE_FIXED, n);
    A.flow = 1;
    A.v_frame = v_frame;
    return run(s, A, false, false);
}

int fqsb_time_steps_until_event(fqsb_system* s, double tol, int64_t niter_tol, int64_t max_iter,
                                int64_t* ret)
{
    TRY(enter(s));
    TRY(require_dynamic(s));
    if (!(tol < 1.0)) {
        return fail(FQSB_EASSERT, ASSERT_MSG("tol < 1.0")); // detail.h:1597
    }
    RunArgs A = make_args(MODE_UNTIL_EVENT, max_iter);
    A.tol = tol;
    A.tol2 = tol * tol;
    A.niter_tol = (int)niter_tol;
    TRY(run(s, A, false, false));
    fill_ret(s, ret);
    return FQSB_OK;
}

static int snapshot_index(fqsb_system* s)
{
    if (!s->d_in) {
        TRY(dev_alloc(s, &s->d_in, (size_t)s->n));
    }
    CU(cudaMemcpyAsync(s->d_in, s->S.idx, (size_t)s->n * sizeof(i64), cudaMemcpyDeviceToDevice,
                       s->stream));
    return FQSB_OK;
}

static int no_convergence(const fqsb_system* s, int max_iter_is_error)
{
    if (!max_iter_is_error) {
        return FQSB_OK;
    }
    for (i64 r = 0; r < s->R; ++r) {
        if (s->h_ctl[r].status == ST_EXHAUSTED) {
            return fail(FQSB_ENOCONV, "No convergence found"); // detail.h:1788,1889
        }
    }
    return FQSB_OK;
}

int fqsb_minimise(fqsb_system* s, double tol, int64_t niter_tol, int64_t max_iter,
                  int time_activity, int max_iter_is_error, int64_t* ret)
{
    TRY(enter(s));
    if (!(tol < 1.0)) {
        return fail(FQSB_EASSERT, ASSERT_MSG("tol < 1.0")); // detail.h:1684
    }
    if (s->par.minimisation == FQSB_MIN_NONE) {
        return fail(FQSB_EUNSUPPORTED, "Minimisation not implementated"); // detail.h:1691-1693
    }
    const bool overdamped = s->par.minimisation == FQSB_MIN_OVERDAMPED;
    if (overdamped && time_activity) {
        return fail(FQSB_EASSERT, ASSERT_MSG("!time_activity")); // detail.h:1696
    }
    RunArgs A = make_args(MODE_MINIMISE, max_iter);
    A.tol = tol;
    A.tol2 = tol * tol;
    A.niter_tol = (int)niter_tol;
    if (time_activity) {
        TRY(snapshot_index(s)); // detail.h:1760-1762
        A.track = 1;
        A.i_n = s->d_in;
    }
    TRY(run(s, A, overdamped, false));
    fill_ret(s, ret);
    return no_convergence(s, max_iter_is_error);
}

int fqsb_minimise_truncate(fqsb_system* s, const int64_t* i_n, int64_t A_truncate,
                           int64_t S_truncate, double tol, int64_t niter_tol, int64_t max_iter,
                           int time_activity, int max_iter_is_error, int64_t* ret)
{
    TRY(enter(s));
    TRY(require_dynamic(s));
    if (!(tol < 1.0)) {
        return fail(FQSB_EASSERT, ASSERT_MSG("tol < 1.0")); // detail.h:1844
    }
    if (!time_activity) {
        return fail(FQSB_EASSERT, ASSERT_MSG("time_activity")); // detail.h:1847
    }
    if (!s->d_in) {
        TRY(dev_alloc(s, &s->d_in, (size_t)s->n));
    }
    CU(cudaMemcpyAsync(s->d_in, i_n, (size_t)s->n * sizeof(i64), cudaMemcpyHostToDevice,
                       s->stream));
    TRY(reduce(s, 4, 1, s->d_in)); // initial S, A against the caller's i_n
    RunArgs A = make_args(MODE_TRUNCATE, max_iter);
    A.tol = tol;
    A.tol2 = tol * tol;
    A.niter_tol = (int)niter_tol;
    A.track = 1;
    A.i_n = s->d_in;
    A.A_truncate = A_truncate;
    A.S_truncate = S_truncate;
    TRY(run(s, A, false, true));
    fill_ret(s, ret);
    return no_convergence(s, max_iter_is_error);
}

// ---- fused "state in -> timeSteps -> state out" with the copies hidden behind the kernels ---------
// system.u = u; system.v = v; system.a = a; system.timeSteps(nsteps); out = system.u, ... as ONE
// call. The realisations of an ensemble are independent, so the handle cuts them into chunks and
// runs chunk c on internal stream c mod 4: its host-to-device copies, updated_u() (k_align), the
// resident kernel over its realisations and the device-to-host copies of its results queue up
// behind each other while the copy engines already move the neighbouring chunks. Same arithmetic,
// same results as the separate calls (fqsb_set_u / _v / _a, fqsb_time_steps, fqsb_get,
// fqsb_mean_f_frame), which remain the fallback for systems the resident kernel does not take.
// u, v, a: [R*size] host (pinned for real overlap) or NULL (keep the current array);
// out_u, out_v, out_a: [R*size] or NULL; out_mean_f_frame: [R] or NULL.
int fqsb_run_from_host(fqsb_system* s, const double* u, const double* v, const double* a, int64_t n,
                       int64_t nsteps, double* out_u, double* out_v, double* out_a,
                       double* out_mean_f_frame)
{
    TRY(enter(s));
    TRY(require_dynamic(s));
    if (n != s->n) {
        return fail(FQSB_EASSERT, ASSERT_MSG("xt::has_shape(arg, m_u.shape())")); // detail.h:1278
    }
    if (nsteps < 0) {
        return fail(FQSB_EASSERT, ASSERT_MSG("n + 1 < std::numeric_limits<long>::max()"));
    }
    ResidentCfg cfg;
    const bool pipelined = use_resident(s, &cfg, MODE_FIXED) && s->R >= 8 && nsteps > 0 &&
                           nsteps <= ((i64)1 << 20) && !s->forces_frozen;
    if (!pipelined) { // the same sequence through the public calls
        if (u) {
            TRY(fqsb_set_u(s, u, n));
        }
        if (v) {
            TRY(fqsb_set_v(s, v, n));
        }
        if (a) {
            TRY(fqsb_set_a(s, a, n));
        }
        TRY(fqsb_time_steps(s, nsteps));
        if (out_u) {
            TRY(fqsb_get(s, FQSB_U, out_u, n));
        }
        if (out_v) {
            TRY(fqsb_get(s, FQSB_V, out_v, n));
        }
        if (out_a) {
            TRY(fqsb_get(s, FQSB_A, out_a, n));
        }
        if (out_mean_f_frame) {
            TRY(fqsb_mean_f_frame(s, out_mean_f_frame));
        }
        return FQSB_OK;
    }
    constexpr int NS = 4;
    if (!s->pipe_ev) {
        for (int k = 0; k < NS; ++k) {
            CU(cudaStreamCreateWithFlags(&s->pipe_stream[k], cudaStreamNonBlocking));
        }
        CU(cudaEventCreateWithFlags(&s->pipe_ev, cudaEventDisableTiming));
    }
    // chunk: a whole number of waves of one-CTA-per-SM realisations, ~32 MB per array
    int sms = 148;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
    i64 per = ((i64)32 << 20) / (s->N * 8);
    per = per < sms ? sms : (per / sms) * sms;
    if (const char* e = std::getenv("FQSB_PIPE_CHUNK")) {
        const i64 x = std::atoll(e);
        per = x > 0 ? x : per;
    }
    const i64 nchunks = (s->R + per - 1) / per;
    CU(cudaEventRecord(s->pipe_ev, s->stream)); // earlier work of the handle comes first
    for (int k = 0; k < NS; ++k) {
        CU(cudaStreamWaitEvent(s->pipe_stream[k], s->pipe_ev, 0));
    }
    RunArgs A = make_args(MODE_FIXED, nsteps);
    A.launch_steps = nsteps;
    A.own_lo = 0;
    A.own_hi = (int)s->N;
    s->last_kernel = "resident";
    invalidate_forces(s);
    for (i64 c = 0; c < nchunks; ++c) {
        const i64 r0 = c * per;
        const i64 cnt = s->R - r0 < per ? s->R - r0 : per;
        const size_t off = (size_t)(r0 * s->N), cn = (size_t)(cnt * s->N);
        cudaStream_t st = s->pipe_stream[c % NS];
        Par Pc = s->P;
        Pc.R = cnt;
        State Sc = s->S;
        Sc.u += off;
        Sc.v += off;
        Sc.a += off;
        Sc.yl += off;
        Sc.yr += off;
        Sc.idx += off;
        Sc.rng += off;
        Sc.u_frame += r0;
        Sc.ctl += r0;
        if (u) {
            CU(cudaMemcpyAsync(Sc.u, u + off, cn * 8, cudaMemcpyHostToDevice, st));
        }
        if (v) {
            CU(cudaMemcpyAsync(Sc.v, v + off, cn * 8, cudaMemcpyHostToDevice, st));
        }
        if (a) {
            CU(cudaMemcpyAsync(Sc.a, a + off, cn * 8, cudaMemcpyHostToDevice, st));
        }
        if (u) {
            k_align<<<grid_for((i64)cn), 256, 0, st>>>(Pc, Sc, nullptr); // updated_u(), detail.h:1280
        }
        k_ctl_begin<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(Pc, Sc, 0, 0, s->d_out);
        cudaError_t e = launch_resident(cfg, Pc, Sc, A, st);
        if (e != cudaSuccess) {
            return cuda_fail(e, "resident kernel launch");
        }
        s->launches += u ? 3 : 2;
        s->kernel_launches++;
        if (out_mean_f_frame) {
            dim3 grid((unsigned)s->S.tiles, (unsigned)cnt);
            double* part = s->d_red + (size_t)r0 * s->S.tiles * 4;
            k_reduce<<<grid, 256, 0, st>>>(Pc, Sc, s->F, 2, 1, nullptr, part, 0, (int)s->N);
            k_reduce_final<<<(unsigned)cnt, 32, 0, st>>>(part, s->S.tiles, s->d_out + 4 * r0);
            CU(cudaMemcpyAsync(s->h_out + 4 * r0, s->d_out + 4 * r0, (size_t)cnt * 4 * sizeof(double),
                               cudaMemcpyDeviceToHost, st));
            s->launches += 2;
        }
        if (out_u) {
            CU(cudaMemcpyAsync(out_u + off, Sc.u, cn * 8, cudaMemcpyDeviceToHost, st));
        }
        if (out_v) {
            CU(cudaMemcpyAsync(out_v + off, Sc.v, cn * 8, cudaMemcpyDeviceToHost, st));
        }
        if (out_a) {
            CU(cudaMemcpyAsync(out_a + off, Sc.a, cn * 8, cudaMemcpyDeviceToHost, st));
        }
        CU(cudaGetLastError());
    }
    for (int k = 0; k < NS; ++k) {
        CU(cudaStreamSynchronize(s->pipe_stream[k]));
    }
    if (out_mean_f_frame) {
        for (i64 r = 0; r < s->R; ++r) {
            out_mean_f_frame[r] = s->h_out[4 * r + 1] / (double)s->N;
        }
    }
    s->steps += s->R * nsteps;
    s->kernel_ms = 0.0;
    return check_flags(s);
}

// ---- event-driven protocol --------------------------------------------------------------------
int fqsb_max_uniform_displacement(fqsb_system* s, int direction, double* out)
{
    TRY(enter(s));
    if (direction != 1 && direction != -1) {
        return fail(FQSB_EASSERT, ASSERT_MSG("direction == 1 || direction == -1"));
    }
    if (s->par.potential == FQSB_POT_SMOOTH) {
        return fail(FQSB_EUNSUPPORTED, "Operation not possible."); // detail.h:420
    }
    TRY(align(s, nullptr)); // m_chunk->align(u), detail.h:180,286
    TRY(reduce_to_host(s, 3, direction, nullptr));
    for (i64 r = 0; r < s->R; ++r) {
        out[r] = s->h_out[4 * r + 1] > 0.0 ? 0.0 : s->h_out[4 * r + 3];
    }
    return check_flags(s);
}

// eventDrivenStep (detail.h:1933-1960) + advanceUniformly(du, false) (detail.h:2027-2050) for
// every realisation: du[r] particle shift, u_frame[r] += du * (k_frame + mu) / k_frame
__global__ void k_event_prepare(const Par P, const State S, double eps, int kick, int direction,
                                const double* red_out, double* du, double* ret)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.R) {
        return;
    }
    double dup;
    if (!kick) {
        double d = red_out[4 * r + 1] > 0.0 ? 0.0 : red_out[4 * r + 3];
        if (d < 0.5 * eps) {
            du[r] = 0.0;
            ret[r] = 0.0;
            return;
        }
        dup = direction > 0 ? d - 0.5 * eps : 0.5 * eps - d;
    }
    else {
        dup = direction > 0 ? eps : -eps;
    }
    double duf = dup * (P.k_frame + P.mu) / P.k_frame;
    du[r] = dup;
    S.u_frame[r] += duf;
    ret[r] = duf;
}

static int advance(fqsb_system* s)
{
    // u += du[r]; updated_u()
    TRY(align(s, s->d_du));
    invalidate_forces(s);
    return FQSB_OK;
}

int fqsb_event_driven_step(fqsb_system* s, double eps, int kick, int direction, double* du_frame)
{
    TRY(enter(s));
    if (direction != 1 && direction != -1) {
        return fail(FQSB_EASSERT, ASSERT_MSG("direction == 1 || direction == -1"));
    }
    if (!kick) {
        if (s->par.potential == FQSB_POT_SMOOTH) {
            return fail(FQSB_EUNSUPPORTED, "Operation not possible.");
        }
        TRY(align(s, nullptr));
        TRY(reduce(s, 3, direction, nullptr));
    }
    const unsigned rg = (unsigned)((s->R + 127) / 128);
    TRY(scratch(s, (size_t)s->R * sizeof(double)));
    double* d_ret = (double*)s->d_scratch;
    k_event_prepare<<<rg, 128, 0, s->stream>>>(s->P, s->S, eps, kick, direction, s->d_out,
                                               s->d_du, d_ret);
    CU(cudaGetLastError());
    s->launches++;
    TRY(advance(s));
    CU(cudaMemcpyAsync(s->h_out, d_ret, (size_t)s->R * sizeof(double), cudaMemcpyDeviceToHost,
                       s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (du_frame) {
        memcpy(du_frame, s->h_out, (size_t)s->R * sizeof(double));
    }
    return check_flags(s);
}

int fqsb_trigger(fqsb_system* s, int64_t r, int64_t p, double eps, int direction)
{
    TRY(enter(s));
    if (r < 0 || r >= s->R || p < 0 || p >= s->N) {
        return fail(FQSB_EASSERT, ASSERT_MSG("(size_type)p < m_N")); // detail.h:1974
    }
    // quirk Q1 (detail.h:1975-1976): u changes, forces and (for non-Cuspy) the well do not
    TRY(ensure_forces(s));
    if (s->par.potential == FQSB_POT_CUSPY) {
        TRY(align(s, nullptr)); // Cuspy::trigger aligns first, detail.h:197
    }
    const i64 g = r * s->N + p;
    double y = 0.0;
    CU(cudaMemcpyAsync(&y, (direction > 0 ? s->S.yr : s->S.yl) + g, sizeof(double),
                       cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    double u = direction > 0 ? y + 0.5 * eps : y - 0.5 * eps; // detail.h:199-204
    CU(cudaMemcpyAsync(s->S.u + g, &u, sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->forces_frozen = true;
    return check_flags(s);
}

__global__ void k_fixed_force_prepare(const Par P, const State S, const double* target,
                                      const double* red_out, double* du)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.R) {
        return;
    }
    double mean = red_out[4 * r + 1] / (double)P.N;
    double dup = (target[r] - mean) / P.mu; // detail.h:1991 (quirk Q10: divides by mu)
    du[r] = dup;
    S.u_frame[r] += dup * (P.k_frame + P.mu) / P.k_frame;
}

int fqsb_advance_to_fixed_force(fqsb_system* s, const double* f_frame, int allow_plastic)
{
    TRY(enter(s));
    TRY(snapshot_index(s)); // detail.h:1990
    TRY(reduce(s, 2, 1, nullptr));
    TRY(scratch(s, (size_t)s->R * sizeof(double)));
    CU(cudaMemcpyAsync(s->d_scratch, f_frame, (size_t)s->R * sizeof(double),
                       cudaMemcpyHostToDevice, s->stream));
    const unsigned rg = (unsigned)((s->R + 127) / 128);
    k_fixed_force_prepare<<<rg, 128, 0, s->stream>>>(s->P, s->S, (const double*)s->d_scratch,
                                                     s->d_out, s->d_du);
    CU(cudaGetLastError());
    s->launches++;
    TRY(advance(s));
    TRY(reduce_to_host(s, 4, 1, s->d_in));
    TRY(check_flags(s));
    if (!allow_plastic) {
        for (i64 r = 0; r < s->R; ++r) {
            if (s->h_out[4 * r + 1] != 0.0) {
                return fail(FQSB_EASSERT,
                            ASSERT_MSG("allow_plastic || xt::all(xt::equal(m_chunk->index_at_align(), i_n))"));
            }
        }
    }
    return FQSB_OK;
}

// ---- yield landscape ----------------------------------------------------------------------------
int fqsb_chunk_index_at_align(fqsb_system* s, int64_t* out, int64_t n)
{
    TRY(enter(s));
    if (n != s->n) {
        return fail(FQSB_EASSERT, ASSERT_MSG("size of the output buffer"));
    }
    CU(cudaMemcpyAsync(out, s->S.idx, (size_t)n * sizeof(i64), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_chunk_left_of_align(fqsb_system* s, double* out, int64_t n)
{
    TRY(enter(s));
    if (n != s->n) {
        return fail(FQSB_EASSERT, ASSERT_MSG("size of the output buffer"));
    }
    CU(cudaMemcpyAsync(out, s->S.yl, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_chunk_right_of_align(fqsb_system* s, double* out, int64_t n)
{
    TRY(enter(s));
    if (n != s->n) {
        return fail(FQSB_EASSERT, ASSERT_MSG("size of the output buffer"));
    }
    CU(cudaMemcpyAsync(out, s->S.yr, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_chunk_align(fqsb_system* s, const double* u, int64_t n)
{
    TRY(enter(s));
    if (n != s->n || !u) {
        return fail(FQSB_EASSERT, ASSERT_MSG("xt::has_shape(u, m_u.shape())"));
    }
    TRY(scratch(s, (size_t)n * sizeof(double)));
    CU(cudaMemcpyAsync(s->d_scratch, u, (size_t)n * sizeof(double), cudaMemcpyHostToDevice,
                       s->stream));
    k_align_to<<<grid_for(s->n), 256, 0, s->stream>>>(s->P, s->S, (const double*)s->d_scratch);
    CU(cudaGetLastError());
    s->launches++;
    invalidate_forces(s);
    return check_flags(s);
}

int fqsb_chunk_data(fqsb_system* s, const int64_t* first, int64_t nyield, double* out)
{
    TRY(enter(s));
    if (nyield < 1 || nyield > (1 << 24)) {
        return fail(FQSB_EASSERT, ASSERT_MSG("0 < nyield"));
    }
    for (i64 g = 0; g < s->n; ++g) {
        if (first[g] < 0) {
            return fail(FQSB_EASSERT, ASSERT_MSG("first >= 0"));
        }
    }
    const size_t nb_first = (size_t)s->n * sizeof(i64);
    const size_t nb_out = (size_t)s->n * (size_t)nyield * sizeof(double);
    TRY(scratch(s, nb_first + nb_out));
    i64* d_first = (i64*)s->d_scratch;
    double* d_o = (double*)((char*)s->d_scratch + nb_first);
    CU(cudaMemcpyAsync(d_first, first, nb_first, cudaMemcpyHostToDevice, s->stream));
    k_chunk_data<<<grid_for(s->n), 256, 0, s->stream>>>(s->P, s->S, d_first, (int)nyield, d_o);
    CU(cudaGetLastError());
    s->launches++;
    CU(cudaMemcpyAsync(out, d_o, nb_out, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_chunk_state_at(fqsb_system* s, const int64_t* index, uint64_t* state, int64_t n)
{
    TRY(enter(s));
    if (n != s->n) {
        return fail(FQSB_EASSERT, ASSERT_MSG("size of the output buffer"));
    }
    std::vector<u64> rng((size_t)n);
    std::vector<i64> idx((size_t)n);
    CU(cudaMemcpyAsync(rng.data(), s->S.rng, (size_t)n * sizeof(u64), cudaMemcpyDeviceToHost,
                       s->stream));
    CU(cudaMemcpyAsync(idx.data(), s->S.idx, (size_t)n * sizeof(i64), cudaMemcpyDeviceToHost,
                       s->stream));
    CU(cudaStreamSynchronize(s->stream));
    for (i64 g = 0; g < n; ++g) {
        // the stored state's next draw is d_{i+2}
        state[g] = s->P.consumes ? pcg_advance(rng[(size_t)g], index[g] - (idx[(size_t)g] + 2))
                                 : rng[(size_t)g];
    }
    return FQSB_OK;
}

int fqsb_chunk_restore(fqsb_system* s, const uint64_t* state, const double* value,
                       const int64_t* index, int64_t n)
{
    TRY(enter(s));
    if (n != s->n) {
        return fail(FQSB_EASSERT, ASSERT_MSG("size of the input buffers"));
    }
    for (i64 g = 0; g < n; ++g) {
        if (index[g] < 0) {
            return fail(FQSB_EASSERT, ASSERT_MSG("index >= 0"));
        }
    }
    const size_t nb = (size_t)n * 8;
    TRY(scratch(s, 3 * nb));
    char* base = (char*)s->d_scratch;
    CU(cudaMemcpyAsync(base, state, nb, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(base + nb, value, nb, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(base + 2 * nb, index, nb, cudaMemcpyHostToDevice, s->stream));
    k_chunk_restore<<<grid_for(s->n), 256, 0, s->stream>>>(
        s->P, s->S, (const u64*)base, (const double*)(base + nb), (const i64*)(base + 2 * nb));
    CU(cudaGetLastError());
    s->launches++;
    TRY(align(s, nullptr));
    invalidate_forces(s);
    return check_flags(s);
}

static int avalanche_out(fqsb_system* s, const i64* i_n_dev, int64_t* out_S, int64_t* out_A)
{
    TRY(reduce_to_host(s, 4, 1, i_n_dev));
    for (i64 r = 0; r < s->R; ++r) {
        if (out_S) {
            out_S[r] = (int64_t)s->h_out[4 * r];
        }
        if (out_A) {
            out_A[r] = (int64_t)s->h_out[4 * r + 1];
        }
    }
    return FQSB_OK;
}

int fqsb_avalanche(fqsb_system* s, const int64_t* i_n, int64_t n, int64_t* out_S, int64_t* out_A)
{
    TRY(enter(s));
    if (n != s->n || !i_n) {
        return fail(FQSB_EASSERT, ASSERT_MSG("xt::has_shape(i_n, m_u.shape())"));
    }
    if (!s->d_in) {
        TRY(dev_alloc(s, &s->d_in, (size_t)s->n));
    }
    CU(cudaMemcpyAsync(s->d_in, i_n, (size_t)s->n * sizeof(i64), cudaMemcpyHostToDevice,
                       s->stream));
    return avalanche_out(s, s->d_in, out_S, out_A);
}

static int mark_index(fqsb_system* s)
{
    if (!s->d_mark) {
        TRY(dev_alloc(s, &s->d_mark, (size_t)s->n));
    }
    CU(cudaMemcpyAsync(s->d_mark, s->S.idx, (size_t)s->n * sizeof(i64), cudaMemcpyDeviceToDevice,
                       s->stream));
    return FQSB_OK;
}

// N1: the examples' `i_n = system.chunk.index_at_align` kept on the device
int fqsb_mark_indices(fqsb_system* s)
{
    TRY(enter(s));
    return mark_index(s);
}

int fqsb_avalanche_since_mark(fqsb_system* s, int64_t* out_S, int64_t* out_A)
{
    TRY(enter(s));
    if (!s->d_mark) {
        return fail(FQSB_EASSERT, "no marked indices (call fqsb_mark_indices first)");
    }
    return avalanche_out(s, s->d_mark, out_S, out_A);
}

int fqsb_event_record(fqsb_system* s, int64_t* S_abs, int64_t* A, int64_t* first, int64_t* last)
{
    TRY(enter(s));
    TRY(pull_ctl(s));
    for (i64 r = 0; r < s->R; ++r) {
        const Ctl& c = s->h_ctl[r];
        if (S_abs) {
            S_abs[r] = c.S;
        }
        if (A) {
            A[r] = c.A;
        }
        if (first) {
            first[r] = c.qs_first;
        }
        if (last) {
            last[r] = c.qs_last;
        }
    }
    return FQSB_OK;
}

// ---- slab decomposition primitives (used by frictionqpotspringblock_b200/slab.py) ---------------
int fqsb_set_owned_range(fqsb_system* s, int64_t lo, int64_t hi)
{
    TRY(enter(s));
    if (lo < 0 || hi > s->N || lo >= hi) {
        return fail(FQSB_EASSERT, ASSERT_MSG("0 <= lo < hi <= size"));
    }
    s->own_lo = lo;
    s->own_hi = hi;
    return FQSB_OK;
}

int fqsb_logged_steps(fqsb_system* s, int64_t k, double* log)
{
    TRY(enter(s));
    if (k < 1) {
        return fail(FQSB_EASSERT, ASSERT_MSG("k >= 1"));
    }
    const size_t need = (size_t)s->R * (size_t)k * FQSB_NLOG;
    if (need > s->log_cap) {
        TRY(dev_alloc(s, &s->d_log, need));
        s->log_cap = need;
    }
    CU(cudaMemsetAsync(s->d_log, 0, need * sizeof(double), s->stream));
    RunArgs A = make_args(MODE_LOG, k);
    A.log = s->d_log;
    TRY(run(s, A, s->par.minimisation == FQSB_MIN_OVERDAMPED, false));
    CU(cudaMemcpyAsync(log, s->d_log, need * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_snapshot(fqsb_system* s)
{
    TRY(enter(s));
    const size_t n = (size_t)s->n;
    if (!s->snap_idx) {
        for (int k = 0; k < 5; ++k) {
            TRY(dev_alloc(s, &s->snap_d[k], n));
        }
        TRY(dev_alloc(s, &s->snap_idx, n));
        TRY(dev_alloc(s, &s->snap_rng, n));
        TRY(dev_alloc(s, &s->snap_uf, (size_t)s->R));
        TRY(dev_alloc(s, &s->snap_ctl, (size_t)s->R));
    }
    const double* src[5] = {s->S.u, s->S.v, s->S.a, s->S.yl, s->S.yr};
    for (int k = 0; k < 5; ++k) {
        CU(cudaMemcpyAsync(s->snap_d[k], src[k], n * 8, cudaMemcpyDeviceToDevice, s->stream));
    }
    CU(cudaMemcpyAsync(s->snap_idx, s->S.idx, n * 8, cudaMemcpyDeviceToDevice, s->stream));
    CU(cudaMemcpyAsync(s->snap_rng, s->S.rng, n * 8, cudaMemcpyDeviceToDevice, s->stream));
    CU(cudaMemcpyAsync(s->snap_uf, s->S.u_frame, (size_t)s->R * 8, cudaMemcpyDeviceToDevice,
                       s->stream));
    CU(cudaMemcpyAsync(s->snap_ctl, s->S.ctl, (size_t)s->R * sizeof(Ctl),
                       cudaMemcpyDeviceToDevice, s->stream));
    s->snap_valid = true;
    return FQSB_OK;
}

int fqsb_rollback(fqsb_system* s)
{
    TRY(enter(s));
    if (!s->snap_valid) {
        return fail(FQSB_EASSERT, "no snapshot to roll back to");
    }
    const size_t n = (size_t)s->n;
    double* dst[5] = {s->S.u, s->S.v, s->S.a, s->S.yl, s->S.yr};
    for (int k = 0; k < 5; ++k) {
        CU(cudaMemcpyAsync(dst[k], s->snap_d[k], n * 8, cudaMemcpyDeviceToDevice, s->stream));
    }
    CU(cudaMemcpyAsync(s->S.idx, s->snap_idx, n * 8, cudaMemcpyDeviceToDevice, s->stream));
    CU(cudaMemcpyAsync(s->S.rng, s->snap_rng, n * 8, cudaMemcpyDeviceToDevice, s->stream));
    CU(cudaMemcpyAsync(s->S.u_frame, s->snap_uf, (size_t)s->R * 8, cudaMemcpyDeviceToDevice,
                       s->stream));
    CU(cudaMemcpyAsync(s->S.ctl, s->snap_ctl, (size_t)s->R * sizeof(Ctl),
                       cudaMemcpyDeviceToDevice, s->stream));
    invalidate_forces(s);
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

static int cells_args(fqsb_system* s, int64_t first, int64_t count)
{
    if (first < 0 || count < 1 || first + count > s->n) {
        return fail(FQSB_EASSERT, ASSERT_MSG("0 <= first, first + count <= size"));
    }
    return FQSB_OK;
}

int fqsb_export_cells(fqsb_system* s, int64_t first, int64_t count, void* buf, int on_device)
{
    TRY(enter(s));
    TRY(cells_args(s, first, count));
    const size_t bytes = (size_t)count * 7 * 8;
    u64* dbuf = (u64*)buf;
    if (!on_device) {
        TRY(scratch(s, bytes));
        dbuf = (u64*)s->d_scratch;
    }
    k_export_cells<<<grid_for(count), 256, 0, s->stream>>>(s->S, first, count, dbuf);
    CU(cudaGetLastError());
    s->launches++;
    if (!on_device) {
        CU(cudaMemcpyAsync(buf, dbuf, bytes, cudaMemcpyDeviceToHost, s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_import_cells(fqsb_system* s, int64_t first, int64_t count, const void* buf, int on_device)
{
    TRY(enter(s));
    TRY(cells_args(s, first, count));
    const size_t bytes = (size_t)count * 7 * 8;
    const u64* dbuf = (const u64*)buf;
    if (!on_device) {
        TRY(scratch(s, bytes));
        CU(cudaMemcpyAsync(s->d_scratch, buf, bytes, cudaMemcpyHostToDevice, s->stream));
        dbuf = (const u64*)s->d_scratch;
    }
    k_import_cells<<<grid_for(count), 256, 0, s->stream>>>(s->S, first, count, dbuf);
    CU(cudaGetLastError());
    s->launches++;
    invalidate_forces(s);
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

// advanceUniformly(du, false) with externally agreed displacements (detail.h:2027-2050)
int fqsb_advance_uniformly(fqsb_system* s, const double* du, const double* du_frame)
{
    TRY(enter(s));
    TRY(pull_ctl(s)); // (keeps the stream ordered; cheap)
    std::vector<double> uf((size_t)s->R);
    CU(cudaMemcpyAsync(uf.data(), s->S.u_frame, (size_t)s->R * 8, cudaMemcpyDeviceToHost,
                       s->stream));
    CU(cudaStreamSynchronize(s->stream));
    for (i64 r = 0; r < s->R; ++r) {
        uf[(size_t)r] += du_frame[r];
    }
    CU(cudaMemcpyAsync(s->S.u_frame, uf.data(), (size_t)s->R * 8, cudaMemcpyHostToDevice,
                       s->stream));
    CU(cudaMemcpyAsync(s->d_du, du, (size_t)s->R * 8, cudaMemcpyHostToDevice, s->stream));
    TRY(advance(s));
    return check_flags(s);
}

// raw per-realisation sums over the owned range: out [R][4]
//   what 1: {sum f^2, sum f_frame^2}   2: {sum v^2, sum f_frame}
//   what 3: {-, off-branch count, -, min displacement}   4: {sum (i-i_n), #(i != i_n), sum |i-i_n|}
int fqsb_reduce_sums(fqsb_system* s, int what, int direction, const int64_t* i_n, int64_t n,
                     double* out)
{
    TRY(enter(s));
    if (what < 1 || what > 4) {
        return fail(FQSB_EASSERT, "unknown reduction");
    }
    const i64* d_in = nullptr;
    if (what == 4) {
        if (n != s->n || !i_n) {
            return fail(FQSB_EASSERT, ASSERT_MSG("xt::has_shape(i_n, m_u.shape())"));
        }
        if (!s->d_in) {
            TRY(dev_alloc(s, &s->d_in, (size_t)s->n));
        }
        CU(cudaMemcpyAsync(s->d_in, i_n, (size_t)s->n * sizeof(i64), cudaMemcpyHostToDevice,
                           s->stream));
        d_in = s->d_in;
    }
    if (what == 3) {
        TRY(align(s, nullptr));
    }
    if (what == 1 && s->lr_gemm) {
        TRY(ensure_forces(s));
        what = 0;
    }
    TRY(reduce_to_host(s, what, direction, d_in));
    memcpy(out, s->h_out, (size_t)s->R * 4 * sizeof(double));
    return FQSB_OK;
}

// ---------------------------------------------------------------------------------------------
// External = RandomNormalForcing (detail.h:881-1000); the thermal classes of Line1d.h:261-330,
// 486-556 and Particles.h construct it right after the generator, before initSystem's refresh()
// ---------------------------------------------------------------------------------------------
int fqsb_enable_random_forcing(fqsb_system* s, double mean, double stddev, uint64_t seed_forcing,
                               int64_t seed_forcing_stride, const int64_t* dinc_init,
                               const int64_t* dinc, int64_t n)
{
    TRY(enter(s));
    if (s->thermal) {
        return fail(FQSB_EASSERT, "random forcing already enabled");
    }
    if (n != s->n) { // detail.h:978 has_shape
        return fail(FQSB_EASSERT, ASSERT_MSG("xt::has_shape(dinc, m_f_thermal.shape())"));
    }
    const bool combo_ok = s->P.pot == POT_CUSPY && (s->P.inter == INT_LAPLACE1D ||
                                                    s->P.inter == INT_QUARTIC1D ||
                                                    s->P.inter == INT_NONE);
    if (!combo_ok || s->par.minimisation == FQSB_MIN_OVERDAMPED) {
        return fail(FQSB_EUNSUPPORTED,
                    "random forcing exists for Cuspy x {Laplace1d, Quartic1d, no interactions}");
    }
    Thermal& T = s->th;
    T.mean = mean;
    T.sigma_sqrt2 = stddev * std::sqrt(2.0);
    TRY(dev_alloc(s, &T.state, (size_t)s->R));
    TRY(dev_alloc(s, &T.next, (size_t)s->n));
    {
        i64* d = nullptr;
        TRY(dev_alloc(s, &d, (size_t)s->n));
        T.dinc = d;
        CU(cudaMemcpyAsync(d, dinc, (size_t)s->n * sizeof(i64), cudaMemcpyHostToDevice, s->stream));
    }
    TRY(dev_alloc(s, &T.f_ext, (size_t)s->n));
    TRY(dev_alloc(s, &T.f_sys, (size_t)s->n));
    std::vector<u64> st((size_t)s->R);
    const u64 stride = seed_forcing_stride > 0 ? (u64)seed_forcing_stride : 1ULL;
    for (i64 r = 0; r < s->R; ++r) { // m_rng.seed(seed): prrng's default initseq
        st[(size_t)r] = pcg_seed_seq(seed_forcing + (u64)r * stride, FQSB_PCG_DEFAULT_INITSEQ,
                                     &T.inc_rng);
    }
    CU(cudaMemcpyAsync(T.state, st.data(), (size_t)s->R * sizeof(u64), cudaMemcpyHostToDevice,
                       s->stream));
    CU(cudaMemcpyAsync(T.next, dinc_init, (size_t)s->n * sizeof(i64), cudaMemcpyHostToDevice,
                       s->stream));
    CU(cudaMemsetAsync(T.f_ext, 0, (size_t)s->n * sizeof(double), s->stream));
    CU(cudaMemsetAsync(T.f_sys, 0, (size_t)s->n * sizeof(double), s->stream));
    s->S.f_thermal = T.f_sys;
    s->P.thermal = 1;
    s->thermal = true;
    TRY(thermal_update(s, 0, 0)); // initSystem -> refresh() -> updated_inc(), detail.h:1138
    invalidate_forces(s);
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

static int require_thermal(const fqsb_system* s, int64_t n, bool per_block)
{
    if (!s->thermal) {
        return fail(FQSB_EASSERT, "not a RandomForcing system");
    }
    if (per_block && n != s->n) {
        return fail(FQSB_EASSERT, ASSERT_MSG("xt::has_shape(arg, m_f_thermal.shape())"));
    }
    return FQSB_OK;
}

int fqsb_external_get_f_thermal(fqsb_system* s, double* out, int64_t n) // detail.h:967-970
{
    TRY(enter(s));
    TRY(require_thermal(s, n, true));
    CU(cudaMemcpyAsync(out, s->th.f_ext, (size_t)n * 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_external_set_f_thermal(fqsb_system* s, const double* f, int64_t n) // detail.h:976-980
{
    TRY(enter(s));
    TRY(require_thermal(s, n, true));
    CU(cudaMemcpyAsync(s->th.f_ext, f, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_external_get_next(fqsb_system* s, int64_t* out, int64_t n) // detail.h:986-989
{
    TRY(enter(s));
    TRY(require_thermal(s, n, true));
    CU(cudaMemcpyAsync(out, s->th.next, (size_t)n * 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_external_set_next(fqsb_system* s, const int64_t* next, int64_t n) // detail.h:995-999
{
    TRY(enter(s));
    TRY(require_thermal(s, n, true));
    CU(cudaMemcpyAsync(s->th.next, next, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_external_get_state(fqsb_system* s, uint64_t* out) // [R] detail.h:949-952
{
    TRY(enter(s));
    TRY(require_thermal(s, 0, false));
    CU(cudaMemcpyAsync(out, s->th.state, (size_t)s->R * 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

int fqsb_external_set_state(fqsb_system* s, const uint64_t* state) // [R] detail.h:958-961
{
    TRY(enter(s));
    TRY(require_thermal(s, 0, false));
    CU(cudaMemcpyAsync(s->th.state, state, (size_t)s->R * 8, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FQSB_OK;
}

} // extern "C"

#include "fqsb_slab.inl"
