// fqsb_device.cuh -- device-side building blocks of the B200-native integrator.
//
// Arithmetic contract: every per-block expression keeps the evaluation order of the
// reference (include/FrictionQPotSpringBlock/detail.h, cited per function as detail.h:LINE)
// and this translation unit is compiled with -fmad=false, so u, v, a and the forces are a
// deterministic IEEE-754 function of the inputs -- bit-identical to the CPU oracle
// (oracle/fqsb_oracle.c, -ffp-contract=off). Only reductions (residual norms) and libm
// functions (sin, log, pow) may differ in the last bits.
#pragma once

#include <cstdint>
#include <cstring>
#include <type_traits>
#include <cuda_runtime.h>

namespace fqsb {

typedef long long i64;
typedef unsigned long long u64;

// ---- parameters shared by all realisations of a handle (kernel argument, by value) --------
struct Par {
    int pot, inter, rank, dist;
    int rows, cols;
    int consumes; // distribution consumes pcg32 draws (everything but delta)
    int thermal;  // External = RandomNormalForcing (detail.h:881-1000): State::f_thermal is set
    // rows per CTA of the 2-D row-marching kernels (Verlet step / no-passing sweep), chosen per
    // handle so that the grid fills whole waves of resident CTAs (plan_band_rows)
    int s2_ty, s2_ty_np;
    // opt-in contracted arithmetic (fqsb_params.kernel bit 7): the resident Verlet kernels built
    // with FMA contraction (fqsb_resident.cu, FQSB_FMA_BUILD). Not bit-identical to the oracle.
    int fma;
    i64 N; // blocks per realisation
    i64 R; // realisations
    double m, inv_m, eta, mu, kappa, k1, k2, k_frame, dt;
    double c2; // (0.5 * dt) * dt, detail.h:1549 (a kernel parameter: costs the hot loops no register)
    double dpar[4];
    double offset;
    u64 seed, seed_stride;
    // slab decomposition: local block p is global block (seed_first + p) mod seed_period
    u64 seed_first, seed_period;
};

// ---- per-realisation control block (device global memory) ---------------------------------
enum Status : int {
    ST_RUNNING = 0,
    ST_CONVERGED = 1, // StopList criterion met -> quenched, returns 0
    ST_EVENT = 2,     // timeStepsUntilEvent: a well index changed -> returns step
    ST_TRUNCATED = 3, // minimise_truncate: A or S reached -> returns step
    ST_EXHAUSTED = 4, // max_iter reached -> returns max_iter + 1
    ST_NAN = 5,
    ST_IDLE = 6
};

enum Mode : int {
    MODE_FIXED = 0,
    MODE_MINIMISE = 1,
    MODE_UNTIL_EVENT = 2,
    MODE_TRUNCATE = 3,
    MODE_LOG = 4 // slab decomposition: record the per-step sums, decide on the host per batch
};

#define FQSB_NLOG 5 // logged per step: sum f^2, sum f_frame^2, hops, dS, dA

#define FQSB_RING 32
// resident stop modes: while the residual is >= tol it is only reduced every
// min(niter_tol, FQSB_SKIP_K) steps (k_resident, "residual sampling")
#define FQSB_SKIP_K 8

struct Ctl {
    int status;
    int init; // "first plastic event not seen yet" (detail.h:1758,1771)
    i64 steps; // steps done so far in the current call
    i64 S, A;  // sum |i - i_n|, #(i != i_n) (detail.h:1863-1864)
    i64 s_n;   // detail.h:1757,1777
    i64 inc;   // increment number m_inc (detail.h:1067)
    i64 qs_first, qs_last; // detail.h:1068-1069
    double residual;       // last residual inserted
    // GooseFEM::Iterate::StopList. Entry k is kept as the pair (num, den) with residual^2 =
    // num / den: num = sum f^2, den = sum f_frame^2 (or 1 when that is 0, detail.h:1516-1519)
    double ring[FQSB_RING];
    double ring_den[FQSB_RING];
    unsigned int count;     // streaming path: CTAs of this realisation that finished the step
    int flip;               // streaming path: which of the two u/v/a buffer sets is current
    int batch;              // blocked path: steps of the next launch (k, or the redo length)
    int pad1;
};

struct RunArgs {
    int mode;
    int track; // maintain S, A against i_n (time_activity / truncate)
    int flow;  // flowSteps: move the frame before every step
    int niter_tol;
    i64 max_steps; // total step budget of the call (max_iter, or n for fixed)
    i64 launch_steps; // resident path: upper bound of steps in this launch
    i64 A_truncate, S_truncate;
    double tol, tol2, v_frame;
    const i64* i_n; // device [R*N] or nullptr
    // slab decomposition: only blocks [own_lo, own_hi) of the local array enter the sums (the
    // rest are halo copies of a neighbour's blocks); MODE_LOG appends the sums of every step
    int own_lo, own_hi;
    double* log; // device [R][max_steps][FQSB_NLOG]
};

struct State {
    double *u, *v, *a;    // current state [R*N]
    double *u2, *v2, *a2; // ping-pong set of the streaming kernels
    double *yl, *yr;      // y[i], y[i+1] of the current well
    i64* idx;             // global well index i
    u64* rng;             // pcg32 state whose next draw is d_{i+2}
    double* u_frame;      // [R]
    Ctl* ctl;             // [R]
    int* err;             // [0] landscape underflow, [1] NaN
    const double* pref;   // LongRange prefactor table [N] (detail.h:829-844)
    double* part;         // streaming path: per-CTA partial sums
    int tiles;            // streaming path: CTAs per realisation
    const double* f_thermal; // System::m_f_thermal [R*N] (detail.h:1057), nullptr if athermal
};

// ---- External = RandomNormalForcing (detail.h:881-1000) ----------------------------------------
// One sequential prrng::pcg32 stream per realisation (m_rng.seed(seed_forcing): prrng's default
// initseq), consumed in block order by the blocks whose `next` increment has come.
#define FQSB_PCG_DEFAULT_INITSEQ 0xda3e39cb94b95bdbULL
struct Thermal {
    double mean, sigma_sqrt2; // normal(mu, sigma) = mu + sigma*sqrt(2) * erf_inv(2r - 1)
    u64 inc_rng;              // (initseq << 1) | 1
    u64* state;               // [R] m_rng state
    i64* next;                // [R*N] m_next
    const i64* dinc;          // [R*N] m_dinc
    double* f_ext;            // [R*N] RandomNormalForcing::m_f_thermal
    double* f_sys;            // [R*N] System::m_f_thermal (the copy made by updated_inc)
};

// ---- K2b (fqsb_blocked.cuh): temporally blocked tiles of a long 1-D line ----------------------
#define FQSB_BK_T 256       // threads per tile
#define FQSB_BK_CTAS 2    // resident tiles per SM: the load / store phase of one overlaps the steps of the other
#define FQSB_BK_MAXSTEPS 64 // upper bound of steps per launch (size of the per-step log)
#define FQSB_BK_PARK 8      // stop modes: steps whose per-thread sums are parked before a reduction (power of 2)
#define FQSB_BK_GROUP 32    // tiles per group of the two-level reduction of the per-step logs

// Slab member whose halo exchange rides on the tile kernel (fixed-step batches of a
// slab-decomposed line, fqsb_slab.inl): the tiles next to the member's halo regions READ them
// from the member's mailbox (filled by the neighbours' previous batch over NVLink) instead of from
// the state arrays, and the tiles that own the member's outermost blocks WRITE them straight into
// the neighbours' mailboxes in their write-back -- no push / import launches between batches.
struct BlockedFuse {
    int on;        // 0: plain launch
    int pull;      // read the halo regions from the mailbox (every batch of a call but the first)
    int batch;     // index of this batch within the call
    int n_readers; // tiles that pull (0 when !pull)
    int n_pushers; // tiles that push
    i64 hc;        // blocks per halo region (one side)
    u64* self;     // own mailbox (epoch flags "from prev" / "from next" at words 0 / 1)
    u64* prev;     // the neighbours' mailboxes (peer-mapped)
    u64* next;
    const u64* epoch;    // device: halo exchanges completed before this call
    unsigned int* count; // device [2]: readers done, pushers done (0 between launches)
    volatile int* h_status; // host-mapped: [0] peer timeout
    unsigned long long timeout_ns;
};

struct BlockedArgs {
    int own;    // owned blocks per tile (the last tile may own fewer)
    int H;      // halo blocks on each side (>= ksteps unless the system has no interactions)
    int ksteps; // steps per launch
    int ntiles;
    int nsteps; // fixed-step calls: steps of this launch (stop modes: Ctl::batch)
    int flip;   // fixed-step calls: which set is the input (stop modes: Ctl::flip)
    // second buffer set of the arrays that State does not already double (u2, v2, a2 do)
    double *yl2, *yr2;
    i64* idx2;
    u64* rng2;
    double* uf2; // [R]
    double* log;  // [R][ntiles][FQSB_BK_MAXSTEPS][FQSB_NLOG] per-step sums of every tile
    double* glog; // [R][ngroups][FQSB_BK_MAXSTEPS][FQSB_NLOG] ... of every group of tiles
    unsigned int* gcount; // [R][ngroups] tiles of a group that have finished (0 between launches)
    BlockedFuse fuse;
};

// ---- slab mailboxes (fqsb_slab.inl; also read / written by k_blocked in fused mode) -----------
#define FQSB_SLAB_FLAGS 32 // 8-byte words reserved for the epoch flags at the head of a mailbox

// mailbox layout (8-byte words): flags | halo mail [2 parity][2 side][7 planes][hc] |
// gather [2 parity][world][gcap]
__host__ __device__ __forceinline__ u64* slab_mail(u64* base, int parity, int side, i64 hc)
{
    return base + FQSB_SLAB_FLAGS + (i64)((parity * 2 + side) * 7) * hc;
}

#ifdef __CUDACC__
__device__ __forceinline__ u64 ld_acquire_sys(const u64* p)
{
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(u64* p, u64 v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif

// which tiles of a fused launch touch the exchange: `reader` = its load range (owned blocks, H
// halo blocks and the frozen cell per side, periodic in the n local blocks) meets a halo region
// [0, hc) / [n - hc, n); `pusher` = it owns blocks of [hc, 2 hc) / [n - 2 hc, n - hc)
__host__ __device__ inline void blocked_tile_roles(i64 n, i64 hc, i64 own, i64 H, i64 c,
                                                   bool* reader, bool* pusher)
{
    const i64 own0 = c * own;
    const i64 cnt = n - own0 < own ? n - own0 : own;
    const i64 a0 = own0, a1 = own0 + cnt;
    *pusher = (a0 < 2 * hc && hc < a1) || (a0 < n - hc && n - 2 * hc < a1);
    bool r = false;
    for (int k = -1; k <= 1; ++k) {
        const i64 l0 = own0 - H - 1 + k * n, l1 = own0 + cnt + H + 1 + k * n;
        r = r || (l0 < hc && 0 < l1) || (l0 < n && n - hc < l1);
    }
    *reader = r;
}


// ---- prrng::pcg32 ------------------------------------------------------
#define FQSB_PCG_MULT 0x5851f42d4c957f2dULL
#define FQSB_PCG_MULT_INV 0xc097ef87329e28a5ULL
#define FQSB_PCG_INC 1ULL // initseq = 0 for every block (Line1d.h:151)

__host__ __device__ __forceinline__ u64 pcg_next(u64 s) { return s * FQSB_PCG_MULT + FQSB_PCG_INC; }
__host__ __device__ __forceinline__ u64 pcg_prev(u64 s)
{
    return (s - FQSB_PCG_INC) * FQSB_PCG_MULT_INV;
}

__host__ __device__ __forceinline__ u64 pcg_seed(u64 initstate)
{
    u64 s = 0ULL;
    s = pcg_next(s);
    s += initstate;
    s = pcg_next(s);
    return s;
}

// the [0,1) double the generator emits from state `old` (32 random mantissa bits)
__host__ __device__ __forceinline__ double pcg_double(u64 old)
{
    unsigned int xs = (unsigned int)(((old >> 18u) ^ old) >> 27u);
    unsigned int rot = (unsigned int)(old >> 59u);
    unsigned int out = (xs >> rot) | (xs << ((0u - rot) & 31u));
    u64 bits = ((u64)out << 20) | 0x3ff0000000000000ULL;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((i64)bits) - 1.0;
#else
    double d;
    memcpy(&d, &bits, sizeof d);
    return d - 1.0;
#endif
}

// O(log n) jump (host side: state_at)
__host__ __device__ inline u64 pcg_advance(u64 s, i64 distance)
{
    u64 delta = (u64)distance;
    u64 cur_mult = FQSB_PCG_MULT, cur_plus = FQSB_PCG_INC, acc_mult = 1ULL, acc_plus = 0ULL;
    while (delta > 0) {
        if (delta & 1ULL) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1ULL) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1ULL;
    }
    return acc_mult * s + acc_plus;
}

// the same jump for a stream with an arbitrary increment (the thermal forcing stream)
__host__ __device__ inline u64 pcg_advance_inc(u64 s, u64 delta, u64 inc)
{
    u64 cur_mult = FQSB_PCG_MULT, cur_plus = inc, acc_mult = 1ULL, acc_plus = 0ULL;
    while (delta > 0) {
        if (delta & 1ULL) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1ULL) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1ULL;
    }
    return acc_mult * s + acc_plus;
}

__host__ __device__ inline u64 pcg_seed_seq(u64 initstate, u64 initseq, u64* inc_out)
{
    const u64 inc = (initseq << 1u) | 1u;
    u64 s = 0ULL;
    s = s * FQSB_PCG_MULT + inc;
    s += initstate;
    s = s * FQSB_PCG_MULT + inc;
    *inc_out = inc;
    return s;
}

// ---- erf_inv for prrng::pcg32::normal (thermal systems) ------------------------------------------
// normal(mu, sigma) = mu + sigma*sqrt(2)*erf_inv(2r - 1). With w = -log(1 - z^2):
//   erf_inv(z) = z * p(w - 3.125)          w < 6.25   (|z| < 0.99903)
//              = z * q(sqrt(w) - 3.25)     w < 16
//              = z * s(sqrt(w) - 4.35)     w <= 22    (prrng's doubles: |z| <= 1 - 2^-31, w < 20.8)
// (the decomposition of M. Giles, "Approximating the erfinv function", GPU Computing Gems 2011).
// The polynomials are Chebyshev interpolants of degree 24 / 20 / 12 fitted here against a
// 50-digit erfinv (tools/erfinv_fit.py): max relative error 2.8e-16 / 2.7e-16 / 4.6e-16 when
// evaluated in double. Even and odd parts are two independent FMA chains: the producer warp of the
// thermal kernel waits on the latency of this function once per step, not on its throughput.
// centre 3.125
__device__ const double kErfInvCentral[25] = {
    1.6536545626831027, 0.24015818242558834, -0.006033670871426851,
    -0.0007407025341546431, 0.00018673420801981186, -1.3882523393957483e-05,
    -1.3654691758785656e-06, 4.23478816822246e-07, -2.907039127564132e-08,
    -4.1126604371632185e-09, 1.051223377050429e-09, -5.414303283919504e-11,
    -1.2978805369932565e-11, 2.6305268312595183e-12, -8.07192593899004e-14,
    -4.0020031087558496e-14, 6.521333511502239e-15, -3.94018812230432e-17,
    -1.2215637192404172e-16, 1.5510787009902526e-17, 6.075050702072414e-19,
    -3.4734793888538036e-19, 1.999259988861535e-20, 3.194015548136271e-21,
    -3.5932028927020693e-22};

// centre 3.25
__device__ const double kErfInvMid[21] = {
    3.0838856104922208, 1.0052589676941655, 0.005370914553555033,
    -0.0037512085082247342, 0.00249144209795696, -0.0016882755354488555,
    0.0009532893415794137, -0.0003550378137852452, 2.4031512865758357e-05,
    6.828711739251955e-05, -4.732068066544697e-05, 1.2465028224217455e-05,
    2.93257845538535e-06, -3.985705945648173e-06, 1.4815977874022539e-06,
    -2.761716223816241e-08, -2.4549818963339823e-07, 1.3158663683192966e-07,
    -2.091453334773126e-08, -1.5305397964152548e-08, 7.680200479077053e-09};

// centre 4.35
__device__ const double kErfInvFar[13] = {
    4.193227791584444, 1.0101034080837366, 0.0005424991500763608,
    -0.000529058407104589, 0.00018294384014653055, -5.308491165471322e-05,
    1.62270677686947e-05, -6.545270788708373e-06, 3.507615965115579e-06,
    -1.9824451116752552e-06, 9.905280208516674e-07, -3.934390741898877e-07,
    9.900018726096847e-08};

template <int N>
__device__ __forceinline__ double poly_even_odd(const double (&c)[N], double t)
{
    const double t2 = t * t;
    constexpr int TOP_E = (N - 1) & ~1;            // highest even index
    constexpr int TOP_O = ((N - 1) & 1) ? N - 1 : N - 2; // highest odd index
    double e = c[TOP_E], o = c[TOP_O];
#pragma unroll
    for (int k = TOP_E - 2; k >= 0; k -= 2) {
        e = fma(e, t2, c[k]);
    }
#pragma unroll
    for (int k = TOP_O - 2; k >= 1; k -= 2) {
        o = fma(o, t2, c[k]);
    }
    return fma(o, t, e);
}

__device__ __forceinline__ double erf_inv_dev(double z)
{
    const double w = -log((1.0 - z) * (1.0 + z));
    double p;
    if (w < 6.25) {
        p = poly_even_odd(kErfInvCentral, w - 3.125);
    }
    else if (w < 16.0) {
        p = poly_even_odd(kErfInvMid, sqrt(w) - 3.25);
    }
    else {
        p = poly_even_odd(kErfInvFar, sqrt(w) - 4.35);
    }
    return p * z;
}

// `gamma` and `normal` spacings need long special-function code (a series / continued-fraction
// solve, erf_inv). Compiled into the hot kernels -- even out of line, on a path that never runs --
// it cost them ~20 % (deeper call graph: larger frames, more registers saved around the hop path).
// It therefore exists only in translation units that define FQSB_SLOW_DISTS: the helper kernels of
// fqsb_api.cu and the generic streaming kernels of fqsb_slowdist.cu, which systems with these two
// distributions are routed to (fqsb_stream.cu: launch_stream_step / launch_stream_sweep).
#ifdef FQSB_SLOW_DISTS
// ---- inverse of the regularised lower incomplete gamma function P(a, x) = p -----------------------
// prrng::pcg32::gamma(k, theta) = theta * boost::math::gamma_p_inv(k, r). boost is absent here, so
// this is an own double-precision solve (parity with the oracle's 80-bit one to ~1e-14): P by its
// series (x < a + 1) or Lentz's continued fraction for Q = 1 - P, the root by Halley's iteration
// from the Wilson-Hilferty / small-a starting points. Out of line: only well changes draw.
static __device__ __noinline__ double gamma_p_dev(double a, double x, double gln)
{
    if (x <= 0.0) {
        return 0.0;
    }
    if (x < a + 1.0) {
        double ap = a, del = 1.0 / a, sum = del;
        for (int n = 0; n < 1000; ++n) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) {
                break;
            }
        }
        return sum * exp(-x + a * log(x) - gln);
    }
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 1000; ++i) {
        const double an = -(double)i * ((double)i - a);
        b += 2.0;
        d = an * d + b;
        if (fabs(d) < tiny) {
            d = tiny;
        }
        c = b + an / c;
        if (fabs(c) < tiny) {
            c = tiny;
        }
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) {
            break;
        }
    }
    return 1.0 - exp(-x + a * log(x) - gln) * h;
}

static __device__ __noinline__ double gamma_p_inv_dev(double a, double p)
{
    if (!(a > 0.0) || p != p) {
        return __longlong_as_double(0x7ff8000000000000LL);
    }
    if (p <= 0.0) {
        return 0.0;
    }
    if (p >= 1.0) {
        return __longlong_as_double(0x7ff0000000000000LL);
    }
    const double a1 = a - 1.0, gln = lgamma(a);
    double x, lna1 = 0.0, afac = 0.0;
    if (a > 1.0) {
        lna1 = log(a1);
        afac = exp(a1 * (lna1 - 1.0) - gln);
        const double pp = p < 0.5 ? p : 1.0 - p;
        const double t = sqrt(-2.0 * log(pp));
        x = (2.30753 + t * 0.27061) / (1.0 + t * (0.99229 + t * 0.04481)) - t;
        if (p < 0.5) {
            x = -x;
        }
        const double w = 1.0 - 1.0 / (9.0 * a) - x / (3.0 * sqrt(a));
        x = fmax(1e-3, a * w * w * w);
    }
    else {
        const double t = 1.0 - a * (0.253 + a * 0.12);
        x = p < t ? pow(p / t, 1.0 / a) : 1.0 - log(1.0 - (p - t) / (1.0 - t));
    }
    for (int j = 0; j < 30; ++j) {
        if (x <= 0.0) {
            return 0.0;
        }
        const double err = gamma_p_dev(a, x, gln) - p;
        double t = a > 1.0 ? afac * exp(-(x - a1) + a1 * (log(x) - lna1))
                           : exp(-x + a1 * log(x) - gln);
        const double u = err / t;
        t = u / (1.0 - 0.5 * fmin(1.0, u * (a1 / x - 1.0)));
        x -= t;
        if (x <= 0.0) {
            x = 0.5 * (x + t);
        }
        if (fabs(t) < 1e-15 * x) {
            break;
        }
    }
    return x;
}

// prrng::pcg32::normal(mu, sigma) = mu + sigma * sqrt(2) * erf_inv(2 r - 1) (boost's erf_inv there,
// erf_inv_dev here: a few 1e-16 apart); out of line for the same reason
static __device__ __noinline__ double normal_from_draw_dev(double r, double mu, double sigma)
{
    return mu + (sigma * 1.4142135623730951) * erf_inv_dev(2.0 * r - 1.0);
}

#endif // FQSB_SLOW_DISTS

// ---- arithmetic of the yield landscape: never contracted ------------------------------------
// The landscape is part of the system's definition (pcg32 stream -> spacings -> cumulative sum),
// so its multiply-adds keep their two roundings even in the translation units that are built with
// FMA contraction for the dynamics (FQSB_FMA_BUILD).
#ifdef __CUDA_ARCH__
#define FQSB_XMUL(a, b) __dmul_rn((a), (b))
#define FQSB_XADD(a, b) __dadd_rn((a), (b))
#else
#define FQSB_XMUL(a, b) ((a) * (b))
#define FQSB_XADD(a, b) ((a) + (b))
#endif

// ---- distributions -> yield spacing ------------------------------------
enum { DIST_RANDOM = 0, DIST_DELTA = 1, DIST_EXPONENTIAL = 2, DIST_POWER = 3, DIST_GAMMA = 4,
       DIST_PARETO = 5, DIST_WEIBULL = 6, DIST_NORMAL = 7 };

__host__ __device__ __forceinline__ double spacing_from_draw(const Par& P, double r)
{
    if (P.dist == DIST_RANDOM) {
        return FQSB_XADD(FQSB_XMUL(r, P.dpar[0]), P.dpar[1]);
    }
    switch (P.dist) {
    case DIST_DELTA:
        return P.dpar[0] + P.dpar[1];
    case DIST_EXPONENTIAL:
        return FQSB_XADD(FQSB_XMUL(-log(1.0 - r), P.dpar[0]), P.dpar[1]);
    case DIST_POWER:
        return pow(1.0 - r, 1.0 / (P.dpar[0] + 1.0)) + P.dpar[1];
    case DIST_PARETO:
        return FQSB_XADD(FQSB_XMUL(P.dpar[1], pow(1.0 - r, -1.0 / P.dpar[0])), P.dpar[2]);
#if defined(__CUDA_ARCH__) && defined(FQSB_SLOW_DISTS)
    case DIST_NORMAL: // normal(mu, sigma) + offset
        return normal_from_draw_dev(r, P.dpar[0], P.dpar[1]) + P.dpar[2];
    case DIST_GAMMA: // gamma(k, theta) + offset
        return FQSB_XADD(FQSB_XMUL(P.dpar[1], gamma_p_inv_dev(P.dpar[0], r)), P.dpar[2]);
#endif
    default: // DIST_WEIBULL
        return FQSB_XADD(FQSB_XMUL(P.dpar[1], pow(-log(1.0 - r), 1.0 / P.dpar[0])), P.dpar[2]);
    }
}

// spacing the generator draws next from state st; st itself is not advanced
__host__ __device__ __forceinline__ double spacing_peek(const Par& P, u64 st)
{
    return spacing_from_draw(P, P.consumes ? pcg_double(st) : 0.0);
}

// ---- the yield landscape: prrng::pcg32_tensor_cumsum without a chunk ------------------------
// A block keeps only its current well: yl = y[i], yr = y[i+1], the global index i and the
// generator state st whose next draw is d_{i+2}. Leaving the well regenerates the neighbouring
// yield position from the pcg32 stream. Forward (one LCG step): y[j] = y[j-1] + d_j, exactly the
// sequential cumsum of the reference. Backward (one inverse-LCG step):
// y[i-1] = y[i] - d_i, which is the exact inverse only where the partial sums are exact in
// floating point (`random` / `delta` landscapes: multiples of 2^-31 -- all named configs); for the
// other distributions a left move re-associates the sum by an ulp, as prrng's own backward
// redraw does (the reference's tests only require allclose there). This replaces
// m_chunk->align(u) (detail.h:144,180,197) and align(p,u) (detail.h:1732).
// Returns the signed number of wells moved; sets *underflow when i would drop below 0.
// `index_now()` is only called when the block moves LEFT (the check that the landscape has a
// position there): callers whose index lives in global memory pay that load on backward moves only.
template <class IndexNow>
__host__ __device__ __forceinline__ int well_align_lazy(const Par& P, double u, double& yl,
                                                        double& yr, u64& st, IndexNow&& index_now,
                                                        int* underflow)
{
    int moved = 0;
    if (!(fabs(u) <= 1.7976931348623157e308)) { // NaN / inf: reported by the NaN check
        return 0;
    }
    while (u > yr) {
        double d = spacing_peek(P, st);
        if (P.consumes) {
            st = pcg_next(st);
        }
        yl = yr;
        yr = FQSB_XADD(yr, d);
        ++moved;
    }
    if (!(u > yl)) {
        const i64 i_now = index_now();
        do {
            if (i_now + moved <= 0) {
                *underflow = 1;
                break;
            }
            // y[i-1] = y[i] - d_i ; d_i is the draw two positions behind st
            u64 sb = st;
            if (P.consumes) {
                st = pcg_prev(st);
                sb = pcg_prev(st);
            }
            double d = spacing_peek(P, sb);
            yr = yl;
            yl = FQSB_XADD(yl, -d);
            --moved;
        } while (!(u > yl));
    }
    return moved;
}

__host__ __device__ __forceinline__ int well_align(const Par& P, double u, double& yl, double& yr,
                                                   u64& st, i64 i_now, int* underflow)
{
    return well_align_lazy(P, u, yl, yr, st, [i_now]() { return i_now; }, underflow);
}

// ---- correctly rounded a / b for a loop-invariant divisor ------------------------------------
// rcp = RN(1 / b) (IEEE division, computed once). q0 = RN(a * rcp) is within 2 ulp of a / b; one
// residual step (r = a - q*b exact through FMA, q += r*rcp) makes it faithful, a second one
// yields RN(a / b) (Markstein's theorem) -- the same bits as the IEEE division the reference
// performs, at 5 FP64 instructions instead of ~50. (Explicit fma(): this file is compiled with
// -fmad=false, which only forbids *implicit* contraction.)
__device__ __forceinline__ double div_by_invariant(double a, double b, double rcp)
{
    double q = a * rcp;
    double r = fma(-q, b, a);
    q = fma(r, rcp, q);
    r = fma(-q, b, a);
    return fma(r, rcp, q);
}

// ---- potentials -----------------------------------------------------------------------------
enum { POT_CUSPY = 0, POT_SEMISMOOTH = 1, POT_SMOOTH = 2 };

// UNIT: the caller guarantees mu == 1, m == 1 and k1 == 1 exactly, so the multiplications by
// those parameters (x * 1.0 == x in IEEE-754) are skipped without changing a single bit.
template <int POT, bool UNIT = false>
__device__ __forceinline__ double f_potential(const Par& P, double u, double yl, double yr)
{
    if (POT == POT_CUSPY) { // detail.h:164-169
        return UNIT ? (0.5 * (yl + yr) - u) : (0.5 * (yl + yr) - u) * P.mu;
    }
    else if (POT == POT_SEMISMOOTH) { // detail.h:261-276
        double xi = 0.5 * (yl + yr);
        const double mk = P.mu + P.kappa, rmk = 1.0 / mk;
        double u_r = div_by_invariant(P.mu * xi + P.kappa * yr, mk, rmk);
        double u_l = div_by_invariant(P.mu * xi + P.kappa * yl, mk, rmk);
        if (u < u_l) {
            return P.kappa * (u - yl);
        }
        else if (u <= u_r) {
            return P.mu * (0.5 * (yl + yr) - u);
        }
        return P.kappa * (u - yr);
    }
    else { // detail.h:402-408
        double umin = 0.5 * (yr + yl);
        double dy = 0.5 * (yr - yl);
        return -P.mu * dy / 3.14159265358979323846 *
               sin(3.14159265358979323846 * (u - umin) / dy);
    }
}

__device__ __forceinline__ double f_potential_rt(const Par& P, double u, double yl, double yr)
{
    switch (P.pot) {
    case POT_CUSPY:
        return f_potential<POT_CUSPY>(P, u, yl, yr);
    case POT_SEMISMOOTH:
        return f_potential<POT_SEMISMOOTH>(P, u, yl, yr);
    default:
        return f_potential<POT_SMOOTH>(P, u, yl, yr);
    }
}

// ---- interactions ---------------------------------------------------------------------------
enum { INT_NONE = 0, INT_LAPLACE1D = 1, INT_QUARTIC1D = 2, INT_QUARTICGRADIENT1D = 3,
       INT_LONGRANGE1D = 4, INT_LAPLACE2D = 5, INT_QUARTICGRADIENT2D = 6 };

// U(q): slip of the block with flat index q of the same realisation (shared memory, or
// recomputed from global memory). p = own flat index, (i, j) = its row/col for rank 2.
// WRAP = false: the caller's U() accepts q = -1 and q = N (ghost cells) for the 1-D stencils.
template <int INT, bool WRAP = true, bool UNIT = false, class UF>
__device__ __forceinline__ double f_interactions(const Par& P, UF&& U, const double* pref, int p,
                                                 int i, int j, double uc)
{
    const int N = (int)P.N;
    if (INT == INT_NONE) {
        return 0.0;
    }
    else if (INT == INT_LAPLACE1D) { // detail.h:480-486
        int l = (WRAP && p == 0) ? N - 1 : p - 1, r = (WRAP && p == N - 1) ? 0 : p + 1;
        return UNIT ? (U(l) - 2 * uc + U(r)) : (U(l) - 2 * uc + U(r)) * P.k1;
    }
    else if (INT == INT_QUARTIC1D) { // detail.h:784-803
        int l = (WRAP && p == 0) ? N - 1 : p - 1, r = (WRAP && p == N - 1) ? 0 : p + 1;
        double um = U(l), up = U(r);
        double dup = up - uc;
        double dun = um - uc;
        return P.k1 * (um - 2 * uc + up) + P.k2 * (dup * dup * dup + dun * dun * dun);
    }
    else if (INT == INT_QUARTICGRADIENT1D) { // detail.h:642-651
        int l = (WRAP && p == 0) ? N - 1 : p - 1, r = (WRAP && p == N - 1) ? 0 : p + 1;
        double um = U(l), up = U(r);
        double du = up - um;
        return (um - 2 * uc + up) * (P.k1 + (0.25 * P.k2) * du * du);
    }
    else if (INT == INT_LONGRANGE1D) { // detail.h:852-866 (same summation order)
        const int m = (N - N % 2) / 2;
        double fp = 0.0;
        for (int q = 0; q < N; ++q) {
            if (q == p) {
                continue;
            }
            int d = q > p ? q - p : p - q;
            if (d > m) {
                d = N - d;
            }
            fp += (U(q) - uc) * pref[d];
        }
        return fp;
    }
    else if (INT == INT_LAPLACE2D) { // detail.h:557-582
        const int R = P.rows, C = P.cols;
        int im = (i == 0 ? R - 1 : i - 1) * C, ip = (i == R - 1 ? 0 : i + 1) * C, ic = i * C;
        int jm = j == 0 ? C - 1 : j - 1, jp = j == C - 1 ? 0 : j + 1;
        double lap = U(im + j) + U(ip + j) + U(ic + jm) + U(ic + jp) - 4 * uc;
        return UNIT ? lap : lap * P.k1;
    }
    else { // QuarticGradient2d, detail.h:700-711
        const int R = P.rows, C = P.cols;
        int im = (i == 0 ? R - 1 : i - 1) * C, ip = (i == R - 1 ? 0 : i + 1) * C, ic = i * C;
        int jm = j == 0 ? C - 1 : j - 1, jp = j == C - 1 ? 0 : j + 1;
        double mk4_3 = P.k2 / 3.0;
        double mk4_23 = 2.0 * mk4_3;
        double u_pj = U(ip + j), u_mj = U(im + j), u_cp = U(ic + jp), u_cm = U(ic + jm);
        double l = u_pj + u_mj + u_cp + u_cm - 4 * uc;
        double dudx = 0.5 * (u_pj - u_mj);
        double dudy = 0.5 * (u_cp - u_cm);
        double d2udxdy = 0.25 * (U(ip + jp) - U(ip + jm) - U(im + jp) + U(im + jm));
        double d2udx2 = u_pj - 2 * uc + u_mj;
        double d2udy2 = u_cp - 2 * uc + u_cm;
        return l * (P.k1 + mk4_3) + mk4_23 * (dudx * dudx * d2udx2 + dudy * dudy * d2udy2 +
                                              2.0 * dudx * dudy * d2udxdy);
    }
}

template <class UF>
__device__ __forceinline__ double f_interactions_rt(const Par& P, UF&& U, const double* pref,
                                                    int p, int i, int j, double uc)
{
    switch (P.inter) {
    case INT_NONE:
        return 0.0;
    case INT_LAPLACE1D:
        return f_interactions<INT_LAPLACE1D>(P, U, pref, p, i, j, uc);
    case INT_QUARTIC1D:
        return f_interactions<INT_QUARTIC1D>(P, U, pref, p, i, j, uc);
    case INT_QUARTICGRADIENT1D:
        return f_interactions<INT_QUARTICGRADIENT1D>(P, U, pref, p, i, j, uc);
    case INT_LONGRANGE1D:
        return f_interactions<INT_LONGRANGE1D>(P, U, pref, p, i, j, uc);
    case INT_LAPLACE2D:
        return f_interactions<INT_LAPLACE2D>(P, U, pref, p, i, j, uc);
    default:
        return f_interactions<INT_QUARTICGRADIENT2D>(P, U, pref, p, i, j, uc);
    }
}

// ---- velocity-Verlet tail of timeStep (detail.h:1552-1565) -----------------------------------
// F = (f_frame + f_potential) + f_interactions at the new u; returns the final residual force
// f = F + f_damping and updates v, a in place (v_n, a_n are the values on entry).
template <bool UNIT = false>
__device__ __forceinline__ double verlet_tail(const Par& P, double F, double& v, double& a)
{
    const double vn = v, an = a;
    const double hdt = 0.5 * P.dt;
    const double meta = -P.eta;
    double vv = vn + P.dt * an;         // 1552
    double f = F + meta * vv;           // 1553: updated_v() -> f = f_frame+f_pot+f_int+f_damp
    double aa = UNIT ? f : f * P.inv_m; // 1555
    vv = vn + hdt * (an + aa);          // 1557
    f = F + meta * vv;                  // 1558
    aa = UNIT ? f : f * P.inv_m;        // 1560
    vv = vn + hdt * (an + aa);          // 1562
    f = F + meta * vv;                  // 1563
    aa = UNIT ? f : f * P.inv_m;        // 1565
    v = vv;
    a = aa;
    return f;
}

// the same with External = RandomNormalForcing: f = f_frame + f_pot + f_int + f_damp + f_thermal
// (detail.h:1326-1329, left to right)
template <bool UNIT = false>
__device__ __forceinline__ double verlet_tail_thermal(const Par& P, double F, double fth, double& v,
                                                      double& a)
{
    const double vn = v, an = a;
    const double hdt = 0.5 * P.dt;
    const double meta = -P.eta;
    double vv = vn + P.dt * an;
    double f = (F + meta * vv) + fth;
    double aa = UNIT ? f : f * P.inv_m;
    vv = vn + hdt * (an + aa);
    f = (F + meta * vv) + fth;
    aa = UNIT ? f : f * P.inv_m;
    vv = vn + hdt * (an + aa);
    f = (F + meta * vv) + fth;
    aa = UNIT ? f : f * P.inv_m;
    v = vv;
    a = aa;
    return f;
}

// ---- bulk asynchronous copies (the TMA engine: cp.async.bulk global -> shared, completion on an
//      mbarrier by byte count). Used by the row-marching 2-D kernels to stage whole rows in shared
//      memory several rows ahead of their use without spending registers on the look-ahead.
__device__ __forceinline__ unsigned smem_u32(const void* p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(u64* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

// make freshly initialised mbarriers visible to the async proxy
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// one arrival that also announces `bytes` of pending bulk-copy traffic for the current phase
__device__ __forceinline__ void mbar_arrive_expect_tx(u64* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

// bytes: multiple of 16; dst and src 16-byte aligned
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, u64* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(u64* bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0u;
}

__device__ __forceinline__ void mbar_arrive(u64* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(u64* bar, unsigned parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- GooseFEM::Iterate::StopList held in the lanes of a warp -----------
// lane l < n holds entry l as the pair (num, den), residual_l^2 = num / den; entries start at
// +inf. The criterion only COMPARES residuals (detail.h:1615,1748,1780,1874):
//     r_l < tol      <=>  num_l < tol^2 * den_l
//     r_{l+1} <= r_l <=>  num_{l+1} * den_l <= num_l * den_{l+1}
// so no square root or division is needed per step (they cost ~150 FP64-pipe instructions per
// warp, redundantly in every warp of a resident CTA). Equivalent in exact arithmetic; in floating
// point it can differ from sqrt()/sqrt() only where two residuals agree to the last bits, i.e.
// where the summation order of the norms already decides.
struct RingEntry {
    double num, den;
};

__device__ __forceinline__ RingEntry ring_entry(double sf, double sff)
{
    RingEntry e;
    e.num = sf;
    e.den = sff != 0.0 ? sff : 1.0; // residual() falls back to |f| when |f_frame| == 0
    return e;
}

__device__ __forceinline__ RingEntry ring_roll_insert(RingEntry ring, RingEntry x, int n, int lane)
{
    RingEntry nxt;
    nxt.num = __shfl_down_sync(0xffffffffu, ring.num, 1);
    nxt.den = __shfl_down_sync(0xffffffffu, ring.den, 1);
    return lane == n - 1 ? x : nxt;
}

__device__ __forceinline__ bool ring_stop(RingEntry ring, int n, int lane, double tol2, double tol4)
{
    RingEntry nxt;
    nxt.num = __shfl_down_sync(0xffffffffu, ring.num, 1);
    nxt.den = __shfl_down_sync(0xffffffffu, ring.den, 1);
    // std::is_sorted(..., greater): never r_{l+1} > r_l
    bool desc = (lane >= n - 1) || !(nxt.num * ring.den > ring.num * nxt.den);
    bool less1 = (lane >= n) || (ring.num < tol2 * ring.den); // all_less(tol): strict
    bool less2 = (lane >= n) || (ring.num < tol4 * ring.den); // all_less(tol * tol)
    bool descending = __all_sync(0xffffffffu, desc);
    bool all1 = __all_sync(0xffffffffu, less1);
    bool all2 = __all_sync(0xffffffffu, less2);
    return (descending && all1) || all2;
}

// ---- warp reductions (fixed butterfly order -> deterministic) --------------------------------
__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x += __shfl_xor_sync(0xffffffffu, x, o);
    }
    return x;
}

__device__ __forceinline__ int warp_sum(int x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x += __shfl_xor_sync(0xffffffffu, x, o);
    }
    return x;
}

__device__ __forceinline__ double warp_min(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x = fmin(x, __shfl_xor_sync(0xffffffffu, x, o));
    }
    return x;
}

// Sums two values over a warp with one butterfly: after the call every lane holds
// (sum of a, sum of b). Stage 1 splits the pair over even/odd lanes, stages 2..5 reduce one
// value per lane, the final shuffles broadcast. Fixed order -> deterministic.
__device__ __forceinline__ void warp_sum2(double& a, double& b)
{
    const int lane = threadIdx.x & 31;
    const bool odd = lane & 1;
    // even lanes collect a, odd lanes collect b
    double send = odd ? a : b;
    double keep = odd ? b : a;
    keep += __shfl_xor_sync(0xffffffffu, send, 1);
#pragma unroll
    for (int o = 16; o > 1; o >>= 1) {
        keep += __shfl_xor_sync(0xffffffffu, keep, o);
    }
    a = __shfl_sync(0xffffffffu, keep, 0);
    b = __shfl_sync(0xffffffffu, keep, 1);
}

// second level: lane l holds the value x_l of an interleaved array (a_0, b_0, a_1, b_1, ...)
// of 32 entries; returns (sum a, sum b) in every lane.
__device__ __forceinline__ void warp_sum_interleaved(double x, double& a, double& b)
{
#pragma unroll
    for (int o = 16; o > 1; o >>= 1) {
        x += __shfl_xor_sync(0xffffffffu, x, o);
    }
    a = __shfl_sync(0xffffffffu, x, 0);
    b = __shfl_sync(0xffffffffu, x, 1);
}

// residual() of detail.h:1512-1520 from the two sums of squares
__device__ __forceinline__ double residual_from_sums(double sf, double sff)
{
    double r_fres = sqrt(sf);
    double r_fext = sqrt(sff);
    return r_fext != 0.0 ? r_fres / r_fext : r_fres;
}

} // namespace fqsb
