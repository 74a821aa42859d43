/*
 * fqsb_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE).
 * See fqsb_oracle.h for scope, provenance and how the restatement is pinned.
 *
 * Every function cites the reference lines it follows as
 *   detail.h:LINE  = /root/reference/include/FrictionQPotSpringBlock/detail.h
 *   Line1d.h:LINE, Line2d.h:LINE likewise; "App. A.x" = SURVEY.md Appendix A.
 */
#define _GNU_SOURCE
#include "fqsb_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static __thread char g_err[512];

const char* orc_last_error(void) { return g_err; }

static int fail(int code, const char* msg)
{
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}

/* ------------------------------------------------------------------------------------------
 * prrng::pcg32 (App. A.1). 64-bit LCG state, XSH-RR output, 32 random mantissa bits.
 * ---------------------------------------------------------------------------------------- */
#define PCG_MULT 0x5851f42d4c957f2dULL

static inline uint32_t pcg_output(uint64_t old)
{
    uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xs >> rot) | (xs << ((-rot) & 31u));
}

static inline uint64_t pcg_next(uint64_t s, uint64_t inc) { return s * PCG_MULT + inc; }

static uint64_t pcg_mult_inv(void)
{
    /* Newton iteration for the inverse of an odd number modulo 2^64 */
    uint64_t x = PCG_MULT;
    for (int k = 0; k < 6; ++k) {
        x *= 2u - PCG_MULT * x;
    }
    return x;
}

static inline uint64_t pcg_prev(uint64_t s, uint64_t inc, uint64_t minv) { return (s - inc) * minv; }

static inline double pcg_double_from_state(uint64_t old)
{
    union {
        uint64_t u;
        double d;
    } x;
    x.u = ((uint64_t)pcg_output(old) << 20) | 0x3ff0000000000000ULL;
    return x.d - 1.0;
}

static uint64_t pcg_seed(uint64_t initstate, uint64_t initseq, uint64_t* inc_out)
{
    uint64_t inc = (initseq << 1u) | 1u;
    uint64_t s = 0u;
    s = pcg_next(s, inc);
    s += initstate;
    s = pcg_next(s, inc);
    *inc_out = inc;
    return s;
}

/* O(log n) LCG jump; negative distance = 2^64 - n (App. A.1 "advance") */
static uint64_t pcg_advance(uint64_t s, uint64_t inc, int64_t distance)
{
    uint64_t delta = (uint64_t)distance;
    uint64_t cur_mult = PCG_MULT, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
    while (delta > 0) {
        if (delta & 1u) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1u) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1u;
    }
    return acc_mult * s + acc_plus;
}

void orc_pcg32_draws(uint64_t initstate, uint64_t initseq, int64_t n, double* out)
{
    uint64_t inc;
    uint64_t s = pcg_seed(initstate, initseq, &inc);
    for (int64_t k = 0; k < n; ++k) {
        out[k] = pcg_double_from_state(s);
        s = pcg_next(s, inc);
    }
}

static double erf_inv_ld(double zd);

/* Inverse of the regularised lower incomplete gamma function P(a, x) = p -- what prrng's
 * pcg32::gamma maps its uniform draws through (boost::math::gamma_p_inv; boost is absent here,
 * "parity unpinned"). Restated in 80-bit long double: P by its series (x < a + 1) or Lentz's
 * continued fraction for Q = 1 - P, the root by Halley's iteration from the Wilson-Hilferty /
 * small-a starting points (Numerical Recipes, 3rd ed., section 6.2.1), run to 1e-19 relative. */
static long double gamma_p_ld(long double a, long double x, long double gln)
{
    if (x <= 0.0L) {
        return 0.0L;
    }
    if (x < a + 1.0L) {
        long double ap = a, del = 1.0L / a, sum = del;
        for (int n = 0; n < 2000; ++n) {
            ap += 1.0L;
            del *= x / ap;
            sum += del;
            if (fabsl(del) < fabsl(sum) * 1e-21L) {
                break;
            }
        }
        return sum * expl(-x + a * logl(x) - gln);
    }
    const long double tiny = 1e-4000L;
    long double b = x + 1.0L - a, c = 1.0L / tiny, d = 1.0L / b, h = d;
    for (int i = 1; i < 2000; ++i) {
        const long double an = -(long double)i * ((long double)i - a);
        b += 2.0L;
        d = an * d + b;
        if (fabsl(d) < tiny) {
            d = tiny;
        }
        c = b + an / c;
        if (fabsl(c) < tiny) {
            c = tiny;
        }
        d = 1.0L / d;
        const long double del = d * c;
        h *= del;
        if (fabsl(del - 1.0L) < 1e-21L) {
            break;
        }
    }
    return 1.0L - expl(-x + a * logl(x) - gln) * h;
}

static double gamma_p_inv_ld(double ad, double pd)
{
    if (!(ad > 0.0) || isnan(pd)) {
        return NAN;
    }
    if (pd <= 0.0) {
        return 0.0;
    }
    if (pd >= 1.0) {
        return INFINITY;
    }
    const long double a = ad, p = pd, a1 = a - 1.0L, gln = lgammal(a);
    long double x, lna1 = 0.0L, afac = 0.0L;
    if (a > 1.0L) {
        lna1 = logl(a1);
        afac = expl(a1 * (lna1 - 1.0L) - gln);
        const long double pp = p < 0.5L ? p : 1.0L - p;
        const long double t = sqrtl(-2.0L * logl(pp));
        x = (2.30753L + t * 0.27061L) / (1.0L + t * (0.99229L + t * 0.04481L)) - t;
        if (p < 0.5L) {
            x = -x;
        }
        const long double w = 1.0L - 1.0L / (9.0L * a) - x / (3.0L * sqrtl(a));
        x = fmaxl(1e-3L, a * w * w * w);
    }
    else {
        const long double t = 1.0L - a * (0.253L + a * 0.12L);
        x = p < t ? powl(p / t, 1.0L / a) : 1.0L - logl(1.0L - (p - t) / (1.0L - t));
    }
    for (int j = 0; j < 40; ++j) {
        if (x <= 0.0L) {
            return 0.0;
        }
        const long double err = gamma_p_ld(a, x, gln) - p;
        long double t = a > 1.0L ? afac * expl(-(x - a1) + a1 * (logl(x) - lna1))
                                 : expl(-x + a1 * logl(x) - gln);
        const long double u = err / t;
        t = u / (1.0L - 0.5L * fminl(1.0L, u * (a1 / x - 1.0L)));
        x -= t;
        if (x <= 0.0L) {
            x = 0.5L * (x + t);
        }
        if (fabsl(t) < 1e-19L * x) {
            break;
        }
    }
    return (double)x;
}

double orc_gamma_p_inv(double a, double p) { return gamma_p_inv_ld(a, p); }

/* distributions -> yield spacing (App. A.2; detail.h:31-66 lists the names) */
double orc_draw_to_spacing(double r, int32_t dist, const double* p)
{
    switch (dist) {
    case ORC_DIST_NORMAL: /* prrng::pcg32::normal(mu, sigma) (+ offset) */
        return p[0] + (p[1] * sqrt(2.0)) * erf_inv_ld(2.0 * r - 1.0) + p[2];
    case ORC_DIST_GAMMA: /* prrng::pcg32::gamma(k, theta) (+ offset) */
        return p[1] * gamma_p_inv_ld(p[0], r) + p[2];
    case ORC_DIST_RANDOM:
        return r * p[0] + p[1];
    case ORC_DIST_DELTA:
        return p[0] + p[1];
    case ORC_DIST_EXPONENTIAL:
        return -log(1.0 - r) * p[0] + p[1];
    case ORC_DIST_POWER:
        return pow(1.0 - r, 1.0 / (p[0] + 1.0)) + p[1];
    case ORC_DIST_PARETO:
        return p[1] * pow(1.0 - r, -1.0 / p[0]) + p[2];
    case ORC_DIST_WEIBULL:
        return p[1] * pow(-log(1.0 - r), 1.0 / p[0]) + p[2];
    default:
        return NAN;
    }
}

static void default_parameters(int32_t dist, int32_t n, const double* in, double* p)
{
    /* prrng defaults: random(scale=1,offset=0) delta(scale=1,offset=0) exponential(scale=1,
     * offset=0) power(k=1,offset=0) pareto(k=1,scale=1,offset=0) weibull(k=1,scale=1,offset=0) */
    double def[4] = {1.0, 0.0, 0.0, 0.0};
    if (dist == ORC_DIST_PARETO || dist == ORC_DIST_WEIBULL || dist == ORC_DIST_GAMMA) {
        def[1] = 1.0;
    }
    if (dist == ORC_DIST_NORMAL) { /* normal(mu = 0, sigma = 1), offset 0 */
        def[0] = 0.0;
        def[1] = 1.0;
    }
    for (int k = 0; k < 4; ++k) {
        p[k] = (k < n) ? in[k] : def[k];
    }
}

/* ------------------------------------------------------------------------------------------
 * prrng::pcg32_tensor_cumsum restated per block (App. A.3).
 * Canonical landscape: y[j] = y[j-1] + d_j (j >= origin+1), y[origin] = value.
 * At construction origin = 0 and y[0] = offset + d_0 (Line1d.h:148-157: cumsum then += offset;
 * identical in exact arithmetic, which holds for every golden: draws are multiples of 2^-31).
 * The stored stretch [base, base+len) grows on demand in both directions; below `origin` it is
 * continued by reversing the LCG and subtracting (as prrng's backward redraw re-associates).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double* y;
    int64_t head;      /* array position of entry `base` */
    int64_t base, len, cap;
    uint64_t st_begin; /* state whose next draw is d_base */
    uint64_t st_end;   /* state whose next draw is d_{base+len} */
    int64_t i;         /* global index at align: y[i] < u <= y[i+1] */
} block_t;

#define BLOCK_MAXLEN (1 << 12)

struct orc_system {
    orc_params par;
    double dpar[4];
    int64_t N, rows, cols;
    uint64_t inc_rng, minv;
    int consumes; /* distribution consumes randomness */
    block_t* blk;
    double *u, *v, *a, *v_n, *a_n;
    double *f, *f_pot, *f_int, *f_frame, *f_damp;
    double* pref; /* LongRange prefactor, detail.h:829-844 */
    int64_t inc, qs_first, qs_last;
    double u_frame, inv_m;
    int err; /* sticky landscape error */
    /* External = RandomNormalForcing<rank> (detail.h:881-1000); thermal == 0: External = void */
    int thermal;
    double th_mean, th_stddev;
    uint64_t th_state, th_inc; /* prrng::pcg32 m_rng */
    int64_t *th_next, *th_dinc;
    double *th_f_ext; /* RandomNormalForcing::m_f_thermal */
    double *f_thermal; /* System::m_f_thermal (copy made by updated_inc) */
};

static inline double blk_draw_fwd(const orc_system* s, uint64_t* st)
{
    double r = 0.0;
    if (s->consumes) {
        r = pcg_double_from_state(*st);
        *st = pcg_next(*st, s->inc_rng);
    }
    return orc_draw_to_spacing(r, s->par.distribution, s->dpar);
}

/* the spacing the generator would draw next from state st (state untouched) */
static inline double blk_peek(const orc_system* s, uint64_t st)
{
    double r = s->consumes ? pcg_double_from_state(st) : 0.0;
    return orc_draw_to_spacing(r, s->par.distribution, s->dpar);
}

static void blk_reserve(block_t* b, int64_t front, int64_t back)
{
    /* make room for `front` more entries before head and `back` more after the end */
    if (b->head >= front && b->head + b->len + back <= b->cap) {
        return;
    }
    int64_t need = b->len + front + back;
    int64_t cap = b->cap > 0 ? b->cap : 128;
    while (cap < 2 * need) {
        cap *= 2;
    }
    double* y = (double*)malloc((size_t)cap * sizeof(double));
    int64_t head = (cap - need) / 2 + front;
    if (b->len > 0) {
        memcpy(y + head, b->y + b->head, (size_t)b->len * sizeof(double));
    }
    free(b->y);
    b->y = y;
    b->head = head;
    b->cap = cap;
}

static void blk_append(const orc_system* s, block_t* b)
{
    if (b->len >= BLOCK_MAXLEN) {
        /* slide: forget the oldest half (can be regenerated backwards from st_begin) */
        int64_t drop = b->len / 2;
        if (s->consumes) {
            b->st_begin = pcg_advance(b->st_begin, s->inc_rng, drop);
        }
        b->head += drop;
        b->base += drop;
        b->len -= drop;
    }
    blk_reserve(b, 0, 1);
    double d = blk_draw_fwd(s, &b->st_end);
    b->y[b->head + b->len] = b->y[b->head + b->len - 1] + d;
    b->len += 1;
}

static int blk_prepend(const orc_system* s, block_t* b)
{
    if (b->base == 0) {
        return 0; /* nothing before the first draw */
    }
    if (b->len >= BLOCK_MAXLEN) {
        int64_t drop = b->len / 2;
        if (s->consumes) {
            b->st_end = pcg_advance(b->st_end, s->inc_rng, -drop);
        }
        b->len -= drop;
    }
    blk_reserve(b, 1, 0);
    /* y[base-1] = y[base] - d_base, where d_base is the next draw of st_begin */
    double d = blk_peek(s, b->st_begin);
    if (s->consumes) {
        b->st_begin = pcg_prev(b->st_begin, s->inc_rng, s->minv);
    }
    b->head -= 1;
    b->base -= 1;
    b->len += 1;
    b->y[b->head] = b->y[b->head + 1] - d;
    return 1;
}

static inline double blk_y(const block_t* b, int64_t j) { return b->y[b->head + (j - b->base)]; }

/* m_chunk->align(p, u): find i with y[i] < u <= y[i+1]   (App. A.3, detail.h:144,1732) */
static int blk_align(orc_system* s, block_t* b, double u)
{
    if (isnan(u)) {
        return 1; /* NaN is reported by timeStep itself (detail.h:1567) */
    }
    int64_t i = b->i;
    for (;;) {
        while (i + 1 >= b->base + b->len) {
            blk_append(s, b);
        }
        while (i < b->base) {
            if (!blk_prepend(s, b)) {
                s->err = 1;
                return 0;
            }
        }
        if (u > blk_y(b, i + 1)) {
            ++i;
        }
        else if (!(u > blk_y(b, i))) {
            --i;
            if (i < 0) {
                s->err = 1;
                b->i = 0;
                return 0;
            }
        }
        else {
            break;
        }
    }
    b->i = i;
    return 1;
}

static void align_all(orc_system* s)
{
    for (int64_t p = 0; p < s->N; ++p) {
        blk_align(s, &s->blk[p], s->u[p]);
    }
}

/* ------------------------------------------------------------------------------------------
 * potentials
 * ---------------------------------------------------------------------------------------- */
static void force_potential(orc_system* s)
{
    const int64_t N = s->N;
    const double mu = s->par.mu, kappa = s->par.kappa;
    align_all(s);
    switch (s->par.potential) {
    case ORC_POT_CUSPY: /* detail.h:139-170 */
        for (int64_t p = 0; p < N; ++p) {
            const block_t* b = &s->blk[p];
            s->f_pot[p] = 0.5 * (blk_y(b, b->i) + blk_y(b, b->i + 1)) - s->u[p];
        }
        for (int64_t p = 0; p < N; ++p) {
            s->f_pot[p] *= mu;
        }
        break;
    case ORC_POT_SEMISMOOTH: /* detail.h:236-277 */
        for (int64_t p = 0; p < N; ++p) {
            const block_t* b = &s->blk[p];
            double y0 = blk_y(b, b->i), y1 = blk_y(b, b->i + 1);
            double xi = 0.5 * (y0 + y1);
            double u_r = (mu * xi + kappa * y1) / (mu + kappa);
            double u_l = (mu * xi + kappa * y0) / (mu + kappa);
            double up = s->u[p];
            if (up < u_l) {
                s->f_pot[p] = kappa * (up - y0);
            }
            else if (up <= u_r) {
                s->f_pot[p] = mu * (0.5 * (y0 + y1) - up);
            }
            else {
                s->f_pot[p] = kappa * (up - y1);
            }
        }
        break;
    default: /* Smooth, detail.h:379-409 */
        for (int64_t p = 0; p < N; ++p) {
            const block_t* b = &s->blk[p];
            double y0 = blk_y(b, b->i), y1 = blk_y(b, b->i + 1);
            double up = s->u[p];
            double umin = 0.5 * (y1 + y0);
            double dy = 0.5 * (y1 - y0);
            s->f_pot[p] = -mu * dy / M_PI * sin(M_PI * (up - umin) / dy);
        }
        break;
    }
}

/* ------------------------------------------------------------------------------------------
 * interactions
 * ---------------------------------------------------------------------------------------- */
static inline int64_t wrap(int64_t i, int64_t n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

static void force_interactions(orc_system* s)
{
    const int64_t N = s->N;
    const double* u = s->u;
    double* f = s->f_int;
    const double k1 = s->par.k1, k2 = s->par.k2;

    switch (s->par.interactions) {
    case ORC_INT_NONE:
        break;
    case ORC_INT_LAPLACE1D: /* detail.h:474-487 */
        for (int64_t p = 1; p < N - 1; ++p) {
            f[p] = u[p - 1] - 2 * u[p] + u[p + 1];
        }
        if (N > 1) {
            f[0] = u[N - 1] - 2 * u[0] + u[1];
            f[N - 1] = u[N - 2] - 2 * u[N - 1] + u[0];
        }
        else {
            f[0] = u[0] - 2 * u[0] + u[0];
        }
        for (int64_t p = 0; p < N; ++p) {
            f[p] *= k1;
        }
        break;
    case ORC_INT_QUARTIC1D: /* detail.h:778-804 */
        for (int64_t p = 0; p < N; ++p) {
            double um = u[wrap(p - 1, N)], uc = u[p], up = u[wrap(p + 1, N)];
            double dup = up - uc;
            double dun = um - uc;
            f[p] = k1 * (um - 2 * uc + up) + k2 * (dup * dup * dup + dun * dun * dun);
        }
        break;
    case ORC_INT_QUARTICGRADIENT1D: { /* detail.h:639-652 */
        double k4_4 = 0.25 * k2;
        for (int64_t p = 0; p < N; ++p) {
            double um = u[wrap(p - 1, N)], uc = u[p], up = u[wrap(p + 1, N)];
            double du = up - um;
            f[p] = (um - 2 * uc + up) * (k1 + k4_4 * du * du);
        }
        break;
    }
    case ORC_INT_LONGRANGE1D: { /* detail.h:849-867 */
        const int64_t n = N;
        const int64_t m = (n - n % 2) / 2;
        for (int64_t p = 0; p < n; ++p) {
            double fp = 0.0;
            double up = u[p];
            for (int64_t i = 0; i < n; ++i) {
                if (i == p) {
                    continue;
                }
                int64_t d = llabs(i - p);
                if (d > m) {
                    d = n - d;
                }
                fp += (u[i] - up) * s->pref[d];
            }
            f[p] = fp;
        }
        break;
    }
    case ORC_INT_LAPLACE2D: { /* detail.h:545-583 */
        const int64_t R = s->rows, C = s->cols;
        for (int64_t i = 0; i < R; ++i) {
            const double* um = u + wrap(i - 1, R) * C;
            const double* up = u + wrap(i + 1, R) * C;
            const double* uc = u + i * C;
            for (int64_t j = 0; j < C; ++j) {
                f[i * C + j] =
                    um[j] + up[j] + uc[wrap(j - 1, C)] + uc[wrap(j + 1, C)] - 4 * uc[j];
            }
        }
        for (int64_t p = 0; p < N; ++p) {
            f[p] *= k1;
        }
        break;
    }
    default: { /* QuarticGradient2d, detail.h:682-750 */
        const int64_t R = s->rows, C = s->cols;
        double mk4_3 = k2 / 3.0;
        double mk4_23 = 2.0 * mk4_3;
        for (int64_t i = 0; i < R; ++i) {
            const double* um = u + wrap(i - 1, R) * C;
            const double* up = u + wrap(i + 1, R) * C;
            const double* uc = u + i * C;
            for (int64_t j = 0; j < C; ++j) {
                int64_t jm = wrap(j - 1, C), jp = wrap(j + 1, C);
                double l = up[j] + um[j] + uc[jp] + uc[jm] - 4 * uc[j];
                double dudx = 0.5 * (up[j] - um[j]);
                double dudy = 0.5 * (uc[jp] - uc[jm]);
                double d2udxdy = 0.25 * (up[jp] - up[jm] - um[jp] + um[jm]);
                double d2udx2 = up[j] - 2 * uc[j] + um[j];
                double d2udy2 = uc[jp] - 2 * uc[j] + uc[jm];
                f[i * C + j] =
                    l * (k1 + mk4_3) + mk4_23 * (dudx * dudx * d2udx2 + dudy * dudy * d2udy2 +
                                                 2.0 * dudx * dudy * d2udxdy);
            }
        }
        break;
    }
    }
}

/* detail.h:1321-1395 */
static void compute_force(orc_system* s)
{
    if (s->thermal) { /* detail.h:1326-1329 */
        for (int64_t p = 0; p < s->N; ++p) {
            s->f[p] = s->f_frame[p] + s->f_pot[p] + s->f_int[p] + s->f_damp[p] + s->f_thermal[p];
        }
        return;
    }
    for (int64_t p = 0; p < s->N; ++p) {
        s->f[p] = s->f_frame[p] + s->f_pot[p] + s->f_int[p] + s->f_damp[p];
    }
}

static void compute_force_frame(orc_system* s)
{
    for (int64_t p = 0; p < s->N; ++p) {
        s->f_frame[p] = s->par.k_frame * (s->u_frame - s->u[p]);
    }
}

static void compute_force_damping(orc_system* s)
{
    const double meta = -s->par.eta;
    for (int64_t p = 0; p < s->N; ++p) {
        s->f_damp[p] = meta * s->v[p];
    }
}

static void updated_u(orc_system* s)
{
    force_potential(s);
    force_interactions(s);
    compute_force_frame(s);
    compute_force(s);
}

static void updated_v(orc_system* s)
{
    compute_force_damping(s);
    compute_force(s);
}

/* ------------------------------------------------------------------------------------------
 * RandomNormalForcing (detail.h:881-1000) on top of prrng::pcg32::normal, whose source is
 * absent: normal(mu, sigma) = mu + sigma*sqrt(2) * erf_inv(2 r - 1), r = next_double(), with
 * boost::math::erf_inv evaluated under boost's default policy (double promoted to long double).
 * Restated as: solve erf(x) = z in 80-bit long double by Halley iterations on glibc's erfl /
 * erfcl (accurate to ~1 ulp of the 64-bit mantissa), then round to double. Both are correctly
 * rounded except when the exact value lies within ~2^-10 ulp of a rounding boundary.
 * Pinned on examples/Line1d_System_Cuspy_Laplace_RandomForcing.h5 (5e6 draws).
 * ---------------------------------------------------------------------------------------- */
static double erf_inv_ld(double zd)
{
    if (zd == 0.0) {
        return 0.0;
    }
    if (zd <= -1.0) {
        return -INFINITY;
    }
    if (zd >= 1.0) {
        return INFINITY;
    }
    const long double z = fabsl((long double)zd);
    const long double q = 1.0L - z; /* exact: z is a double in (0, 1) */
    const long double two_over_sqrtpi = 1.1283791670955125738961589031215452L;
    /* starting point (Winitzki's approximation), good to ~2e-3 relative */
    const long double a = 0.147L;
    const long double ln = logl((1.0L - z) * (1.0L + z));
    const long double t = 2.0L / (3.14159265358979323846264338327950288L * a) + 0.5L * ln;
    long double x = sqrtl(sqrtl(t * t - ln / a) - t);
    for (int it = 0; it < 6; ++it) {
        /* residual in the better conditioned of the two forms */
        long double r = (z < 0.5L) ? (erfl(x) - z) : (q - erfcl(x));
        long double d = two_over_sqrtpi * expl(-x * x);
        long double step = r / d;
        step = step / (1.0L + x * step); /* Halley: f''/f' = -2x */
        x -= step;
        if (fabsl(step) < 1e-21L * x) {
            break;
        }
    }
    return (double)(zd < 0.0 ? -x : x);
}

static double pcg_normal(uint64_t* st, uint64_t inc, double mean, double stddev)
{
    double r = pcg_double_from_state(*st);
    *st = pcg_next(*st, inc);
    return mean + (stddev * sqrt(2.0)) * erf_inv_ld(2.0 * r - 1.0);
}

void orc_pcg32_normal(uint64_t initstate, uint64_t initseq, int64_t n, double mean,
                      double stddev, double* out)
{
    uint64_t inc;
    uint64_t s = pcg_seed(initstate, initseq, &inc);
    for (int64_t k = 0; k < n; ++k) {
        out[k] = pcg_normal(&s, inc, mean, stddev);
    }
}

/* prrng::pcg32::randint(shape, high): unbiased bounded draw (pcg32 `boundedrand`) */
void orc_pcg32_randint(uint64_t initstate, uint64_t initseq, int64_t n, uint32_t high,
                       int64_t* out)
{
    uint64_t inc;
    uint64_t s = pcg_seed(initstate, initseq, &inc);
    const uint32_t threshold = (~high + 1u) % high;
    for (int64_t k = 0; k < n; ++k) {
        for (;;) {
            uint32_t r = pcg_output(s);
            s = pcg_next(s, inc);
            if (r >= threshold) {
                out[k] = (int64_t)(r % high);
                break;
            }
        }
    }
}

double orc_erf_inv(double z) { return erf_inv_ld(z); }

/* detail.h:1369-1375 + 931-943 */
static void updated_inc(orc_system* s)
{
    if (s->thermal) {
        for (int64_t p = 0; p < s->N; ++p) {
            if (s->inc >= s->th_next[p]) {
                s->th_f_ext[p] = pcg_normal(&s->th_state, s->th_inc, s->th_mean, s->th_stddev);
                s->th_next[p] += s->th_dinc[p];
            }
        }
        memcpy(s->f_thermal, s->th_f_ext, (size_t)s->N * sizeof(double));
    }
    compute_force(s);
}

/* ------------------------------------------------------------------------------------------
 * construction: Line1d.h:134-161 (and siblings), Line2d.h:88-117, detail.h:1096-1139
 * ---------------------------------------------------------------------------------------- */
int orc_create(const orc_params* par, orc_system** out)
{
    *out = NULL;
    if (par->rank != 1 && par->rank != 2) {
        return fail(ORC_EASSERT, "rank must be 1 or 2");
    }
    switch (par->distribution) {
    case ORC_DIST_RANDOM:
    case ORC_DIST_DELTA:
    case ORC_DIST_EXPONENTIAL:
    case ORC_DIST_POWER:
    case ORC_DIST_PARETO:
    case ORC_DIST_WEIBULL:
    case ORC_DIST_GAMMA:
    case ORC_DIST_NORMAL:
        break;
    default:
        return fail(ORC_EASSERT, "Unknown distribution");
    }
    orc_system* s = (orc_system*)calloc(1, sizeof *s);
    s->par = *par;
    s->rows = par->shape[0];
    s->cols = par->rank == 2 ? par->shape[1] : 1;
    s->N = s->rows * s->cols;
    const int64_t N = s->N;
    default_parameters(par->distribution, par->nparameters, par->parameters, s->dpar);
    s->consumes = par->distribution != ORC_DIST_DELTA;
    s->minv = pcg_mult_inv();
    s->inv_m = 1.0 / par->m;

    s->blk = (block_t*)calloc((size_t)N, sizeof(block_t));
    for (int64_t p = 0; p < N; ++p) {
        /* initstate = seed + flat index, initseq = 0 (Line1d.h:150-151, Line2d.h:28-34) */
        block_t* b = &s->blk[p];
        uint64_t st = pcg_seed(par->seed + (uint64_t)p, 0u, &s->inc_rng);
        b->st_begin = st;
        b->st_end = st;
        blk_reserve(b, 0, 1);
        double d = blk_draw_fwd(s, &b->st_end);
        b->y[b->head] = par->offset + d;
        b->base = 0;
        b->len = 1;
        b->i = 0;
    }

    double** arr[] = {&s->u,     &s->v,     &s->a,       &s->v_n,    &s->a_n,
                      &s->f,     &s->f_pot, &s->f_int,   &s->f_frame, &s->f_damp};
    for (size_t k = 0; k < sizeof arr / sizeof arr[0]; ++k) {
        *arr[k] = (double*)calloc((size_t)N, sizeof(double));
    }

    if (par->interactions == ORC_INT_LONGRANGE1D) { /* detail.h:829-844 */
        s->pref = (double*)calloc((size_t)N, sizeof(double));
        for (int64_t d = 1; d < N; ++d) {
            s->pref[d] = par->k1 / pow((double)d, par->k2 + 1.0);
        }
    }

    /* refresh(): detail.h:1310-1315 */
    updated_u(s);
    updated_v(s);
    compute_force(s);
    if (s->err) {
        orc_destroy(s);
        return fail(ORC_EASSERT, "u = 0 lies below the first yield position: lower the offset");
    }
    *out = s;
    return ORC_OK;
}

void orc_destroy(orc_system* s)
{
    if (!s) {
        return;
    }
    for (int64_t p = 0; p < s->N; ++p) {
        free(s->blk[p].y);
    }
    free(s->blk);
    free(s->u);
    free(s->v);
    free(s->a);
    free(s->v_n);
    free(s->a_n);
    free(s->f);
    free(s->f_pot);
    free(s->f_int);
    free(s->f_frame);
    free(s->f_damp);
    free(s->pref);
    free(s->th_next);
    free(s->th_dinc);
    free(s->th_f_ext);
    free(s->f_thermal);
    free(s);
}

/* Line1d.h:293-319 (System_Cuspy_Laplace_RandomForcing) and siblings: the athermal system plus
 * External = RandomNormalForcing; initSystem's refresh() already draws for dinc_init <= 0. */
int orc_create_thermal(const orc_params* par, double mean, double stddev, uint64_t seed_forcing,
                       const int64_t* dinc_init, const int64_t* dinc, orc_system** out)
{
    int rc = orc_create(par, out);
    if (rc) {
        return rc;
    }
    orc_system* s = *out;
    const size_t N = (size_t)s->N;
    s->thermal = 1;
    s->th_mean = mean;
    s->th_stddev = stddev;
    /* m_rng.seed(seed): initseq = prrng's default stream */
    s->th_state = pcg_seed(seed_forcing, 0xda3e39cb94b95bdbULL, &s->th_inc);
    s->th_next = (int64_t*)malloc(N * sizeof(int64_t));
    s->th_dinc = (int64_t*)malloc(N * sizeof(int64_t));
    memcpy(s->th_next, dinc_init, N * sizeof(int64_t));
    memcpy(s->th_dinc, dinc, N * sizeof(int64_t));
    s->th_f_ext = (double*)calloc(N, sizeof(double));
    s->f_thermal = (double*)calloc(N, sizeof(double));
    updated_inc(s);
    return ORC_OK;
}

int orc_thermal_get(const orc_system* s, double* f_thermal, int64_t* next, uint64_t* state)
{
    if (!s->thermal) {
        return fail(ORC_EASSERT, "not a RandomForcing system");
    }
    if (f_thermal) {
        memcpy(f_thermal, s->th_f_ext, (size_t)s->N * sizeof(double));
    }
    if (next) {
        memcpy(next, s->th_next, (size_t)s->N * sizeof(int64_t));
    }
    if (state) {
        *state = s->th_state;
    }
    return ORC_OK;
}

int orc_thermal_set(orc_system* s, const double* f_thermal, const int64_t* next,
                    const uint64_t* state)
{
    if (!s->thermal) {
        return fail(ORC_EASSERT, "not a RandomForcing system");
    }
    if (f_thermal) {
        memcpy(s->th_f_ext, f_thermal, (size_t)s->N * sizeof(double));
    }
    if (next) {
        memcpy(s->th_next, next, (size_t)s->N * sizeof(int64_t));
    }
    if (state) {
        s->th_state = *state;
    }
    return ORC_OK;
}

int64_t orc_size(const orc_system* s) { return s->N; }

static int check_landscape(orc_system* s)
{
    if (s->err) {
        s->err = 0;
        return fail(ORC_EASSERT, "yield landscape exhausted below its first entry");
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * setters / getters
 * ---------------------------------------------------------------------------------------- */
int orc_set_u(orc_system* s, const double* u) /* detail.h:1276-1281 */
{
    memcpy(s->u, u, (size_t)s->N * sizeof(double));
    updated_u(s);
    return check_landscape(s);
}

int orc_set_v(orc_system* s, const double* v) /* detail.h:1290-1295 */
{
    memcpy(s->v, v, (size_t)s->N * sizeof(double));
    updated_v(s);
    return ORC_OK;
}

int orc_set_a(orc_system* s, const double* a) /* detail.h:1301-1305 */
{
    memcpy(s->a, a, (size_t)s->N * sizeof(double));
    return ORC_OK;
}

int orc_set_u_frame(orc_system* s, double u_frame) /* detail.h:1253-1258 */
{
    s->u_frame = u_frame;
    compute_force_frame(s);
    compute_force(s);
    return ORC_OK;
}

int orc_set_inc(orc_system* s, int64_t inc) /* detail.h:1241-1247 */
{
    s->inc = inc;
    s->qs_first = inc;
    s->qs_last = inc;
    updated_inc(s);
    return ORC_OK;
}

int orc_set_t(orc_system* s, double t) /* detail.h:1231-1235 */
{
    s->inc = (int64_t)round(t / s->par.dt);
    double tt = (double)s->inc * s->par.dt;
    if (!(fabs(tt - t) <= 1e-8 + 1e-5 * fabs(t))) {
        return fail(ORC_EASSERT, "assertion failed (xt::allclose(this->t(), arg))");
    }
    return ORC_OK;
}

int orc_refresh(orc_system* s) /* detail.h:1310-1315 */
{
    updated_u(s);
    updated_v(s);
    updated_inc(s);
    return check_landscape(s);
}

int orc_quench(orc_system* s) /* detail.h:1527-1532 */
{
    memset(s->v, 0, (size_t)s->N * sizeof(double));
    memset(s->a, 0, (size_t)s->N * sizeof(double));
    updated_v(s);
    return ORC_OK;
}

int orc_get(const orc_system* s, int which, double* out)
{
    const double* src[] = {s->u, s->v, s->a, s->f, s->f_pot, s->f_frame, s->f_int, s->f_damp};
    if (which == 8 && s->thermal) { /* System::m_f_thermal */
        memcpy(out, s->f_thermal, (size_t)s->N * sizeof(double));
        return ORC_OK;
    }
    if (which < 0 || which > 7) {
        return fail(ORC_EASSERT, "bad array id");
    }
    memcpy(out, src[which], (size_t)s->N * sizeof(double));
    return ORC_OK;
}

double orc_u_frame(const orc_system* s) { return s->u_frame; }
int64_t orc_inc(const orc_system* s) { return s->inc; }
int64_t orc_qs_first(const orc_system* s) { return s->qs_first; }
int64_t orc_qs_last(const orc_system* s) { return s->qs_last; }

double orc_residual(const orc_system* s) /* detail.h:1512-1520 */
{
    double sf = 0.0, se = 0.0;
    for (int64_t p = 0; p < s->N; ++p) {
        sf += s->f[p] * s->f[p];
    }
    for (int64_t p = 0; p < s->N; ++p) {
        se += s->f_frame[p] * s->f_frame[p];
    }
    double r_fres = sqrt(sf);
    double r_fext = sqrt(se);
    if (r_fext != 0.0) {
        return r_fres / r_fext;
    }
    return r_fres;
}

double orc_temperature(const orc_system* s) /* detail.h:1500-1503 */
{
    double sv = 0.0;
    for (int64_t p = 0; p < s->N; ++p) {
        sv += s->v[p] * s->v[p];
    }
    return 0.5 * s->par.m * sv / (double)s->N;
}

/* ------------------------------------------------------------------------------------------
 * GooseFEM::Iterate::StopList (App. A.4)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double* r;
    int64_t n;
} stoplist;

static void sl_init(stoplist* l, int64_t n)
{
    l->n = n;
    l->r = (double*)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    for (int64_t k = 0; k < n; ++k) {
        l->r[k] = INFINITY;
    }
}

static void sl_roll_insert(stoplist* l, double x)
{
    for (int64_t k = 0; k + 1 < l->n; ++k) {
        l->r[k] = l->r[k + 1];
    }
    if (l->n > 0) {
        l->r[l->n - 1] = x;
    }
}

static int sl_descending(const stoplist* l)
{
    for (int64_t k = 0; k + 1 < l->n; ++k) {
        if (l->r[k + 1] > l->r[k]) {
            return 0;
        }
    }
    return 1;
}

static int sl_all_less(const stoplist* l, double tol)
{
    for (int64_t k = 0; k < l->n; ++k) {
        if (!(l->r[k] < tol)) {
            return 0;
        }
    }
    return 1;
}

static int sl_stop(const stoplist* l, double tol, double tol2)
{
    return (sl_descending(l) && sl_all_less(l, tol)) || sl_all_less(l, tol2);
}

/* ------------------------------------------------------------------------------------------
 * timeStep: detail.h:1539-1570, one loop per xtensor assignment
 * ---------------------------------------------------------------------------------------- */
static int time_step(orc_system* s)
{
    const int64_t N = s->N;
    const double dt = s->par.dt;
    const double c2 = 0.5 * dt * dt;
    const double hdt = 0.5 * dt;
    const double inv_m = s->inv_m;
    double *u = s->u, *v = s->v, *a = s->a, *v_n = s->v_n, *a_n = s->a_n, *f = s->f;

    s->inc++;
    if (s->thermal) { /* detail.h:1542-1544 */
        updated_inc(s);
    }
    memcpy(v_n, v, (size_t)N * sizeof(double));
    memcpy(a_n, a, (size_t)N * sizeof(double));

    for (int64_t p = 0; p < N; ++p) {
        u[p] = u[p] + dt * v[p] + c2 * a[p];
    }
    updated_u(s);

    for (int64_t p = 0; p < N; ++p) {
        v[p] = v_n[p] + dt * a_n[p];
    }
    updated_v(s);
    for (int64_t p = 0; p < N; ++p) {
        a[p] = f[p] * inv_m;
    }

    for (int rep = 0; rep < 2; ++rep) {
        for (int64_t p = 0; p < N; ++p) {
            v[p] = v_n[p] + hdt * (a_n[p] + a[p]);
        }
        updated_v(s);
        for (int64_t p = 0; p < N; ++p) {
            a[p] = f[p] * inv_m;
        }
    }

    int nan = 0;
    for (int64_t p = 0; p < N; ++p) {
        nan |= isnan(u[p]);
    }
    if (nan) {
        return fail(ORC_ENAN, "NaN entries found");
    }
    return check_landscape(s);
}

int orc_time_steps(orc_system* s, int64_t n) /* detail.h:1577-1583 */
{
    for (int64_t k = 0; k < n; ++k) {
        int rc = time_step(s);
        if (rc) {
            return rc;
        }
    }
    return ORC_OK;
}

int orc_flow_steps(orc_system* s, int64_t n, double v_frame) /* detail.h:1637-1645 */
{
    for (int64_t k = 0; k < n; ++k) {
        s->u_frame += v_frame * s->par.dt;
        int rc = time_step(s);
        if (rc) {
            return rc;
        }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * The FUSED CPU flavour of timeSteps ("best-case CPU"): the same
 * arithmetic as time_step() above -- every per-block expression in the reference's order, so the
 * results are bit-identical (tests/test_oracle_fused.py) -- but as two passes per step over four
 * arrays (new slips; then well test + forces + the three Verlet correctors per block) instead of
 * the ~69 array passes of the reference's xtensor expression structure (detail.h:1539-1570,
 * 1321-1395). Cuspy x Laplace1d without thermal forcing (the headline configuration); any other
 * system takes the faithful path. The force arrays are refreshed once, at the end.
 * ---------------------------------------------------------------------------------------- */
int orc_time_steps_fused(orc_system* s, int64_t n)
{
    if (s->par.potential != ORC_POT_CUSPY || s->par.interactions != ORC_INT_LAPLACE1D ||
        s->thermal || n <= 0) {
        return orc_time_steps(s, n);
    }
    const int64_t N = s->N;
    const double dt = s->par.dt, c2 = 0.5 * dt * dt, hdt = 0.5 * dt;
    const double inv_m = s->inv_m, meta = -s->par.eta, mu = s->par.mu, k = s->par.k1;
    const double kf = s->par.k_frame, uf = s->u_frame;
    double* yl = (double*)malloc((size_t)N * sizeof(double));
    double* yr = (double*)malloc((size_t)N * sizeof(double));
    double* un = (double*)malloc((size_t)N * sizeof(double));
    double *restrict u = s->u, *restrict v = s->v, *restrict a = s->a;
    for (int64_t p = 0; p < N; ++p) {
        const block_t* b = &s->blk[p];
        yl[p] = blk_y(b, b->i);
        yr[p] = blk_y(b, b->i + 1);
    }
    int nan = 0;
    for (int64_t it = 0; it < n; ++it) {
        s->inc++;
        for (int64_t p = 0; p < N; ++p) {
            un[p] = u[p] + dt * v[p] + c2 * a[p]; /* detail.h:1549 */
        }
        for (int64_t p = 0; p < N; ++p) { /* m_chunk->align(u), detail.h:144: rare well changes */
            const double uc = un[p];
            if (uc > yr[p] || !(uc > yl[p])) {
                block_t* b = &s->blk[p];
                blk_align(s, b, uc);
                yl[p] = blk_y(b, b->i);
                yr[p] = blk_y(b, b->i + 1);
            }
        }
#define FUSED_BLOCK(p, ul, ur) \
    { \
        const double uc = un[p]; \
        const double fi = ((ul) - 2 * uc + (ur)) * k;         /* detail.h:474-487 */ \
        const double fp = (0.5 * (yl[p] + yr[p]) - uc) * mu; /* detail.h:164-169 */ \
        const double ff = kf * (uf - uc);                    /* detail.h:1360 */ \
        const double F = ff + fp + fi;                       /* detail.h:1324 */ \
        const double vn = v[p], an = a[p]; \
        double vv = vn + dt * an; /* detail.h:1552-1565 */ \
        double f = F + meta * vv; \
        double aa = f * inv_m; \
        vv = vn + hdt * (an + aa); \
        f = F + meta * vv; \
        aa = f * inv_m; \
        vv = vn + hdt * (an + aa); \
        f = F + meta * vv; \
        aa = f * inv_m; \
        v[p] = vv; \
        a[p] = aa; \
    }
        /* branch-free interior (vectorised by the compiler: every lane keeps IEEE semantics and
           -ffp-contract=off forbids FMA, so the bits do not change), periodic ends apart */
        for (int64_t p = 1; p < N - 1; ++p) {
            FUSED_BLOCK(p, un[p - 1], un[p + 1])
        }
        if (N > 1) {
            FUSED_BLOCK(0, un[N - 1], un[1])
            FUSED_BLOCK(N - 1, un[N - 2], un[0])
        }
        else {
            FUSED_BLOCK(0, un[0], un[0])
        }
#undef FUSED_BLOCK
        double* tmp = u;
        u = un;
        un = tmp;
    }
    if (u != s->u) { /* odd number of steps: the current slips sit in the scratch array */
        memcpy(s->u, u, (size_t)N * sizeof(double));
        un = u;
    }
    for (int64_t p = 0; p < N; ++p) {
        nan |= isnan(s->u[p]);
    }
    free(yl);
    free(yr);
    free(un);
    updated_u(s); /* the stored force arrays of the final state */
    updated_v(s);
    if (nan) {
        return fail(ORC_ENAN, "NaN entries found");
    }
    return check_landscape(s);
}

static int any_index_changed(const orc_system* s, const int64_t* i_n)
{
    for (int64_t p = 0; p < s->N; ++p) {
        if (s->blk[p].i != i_n[p]) {
            return 1;
        }
    }
    return 0;
}

int orc_time_steps_until_event(orc_system* s, double tol, int64_t niter_tol, int64_t max_iter,
                               int64_t* ret) /* detail.h:1595-1622 */
{
    if (!(tol < 1.0)) {
        return fail(ORC_EASSERT, "assertion failed (tol < 1.0)");
    }
    double tol2 = tol * tol;
    stoplist res;
    sl_init(&res, niter_tol);
    int64_t* i_n = (int64_t*)malloc((size_t)s->N * sizeof(int64_t));
    orc_chunk_index_at_align(s, i_n);
    int64_t step;
    int rc = ORC_OK;
    for (step = 1; step < max_iter + 1; ++step) {
        rc = time_step(s);
        if (rc) {
            break;
        }
        if (any_index_changed(s, i_n)) {
            break;
        }
        sl_roll_insert(&res, orc_residual(s));
        if (sl_stop(&res, tol, tol2)) {
            orc_quench(s);
            step = 0;
            break;
        }
    }
    *ret = step;
    free(i_n);
    free(res.r);
    return rc;
}

static int64_t sum_abs_index_diff(const orc_system* s, const int64_t* i_n, int64_t* nchanged)
{
    int64_t S = 0, A = 0;
    for (int64_t p = 0; p < s->N; ++p) {
        int64_t d = s->blk[p].i - i_n[p];
        S += d < 0 ? -d : d;
        A += d != 0;
    }
    if (nchanged) {
        *nchanged = A;
    }
    return S;
}

/* overdamped no-passing sweeps: detail.h:1694-1753 (1-D Laplace in the reference; the 2-D
 * Laplace generalisation -- same local equilibrium with 4 neighbours -- is NEW, "parity
 * unpinned") */
static int minimise_overdamped(orc_system* s, double tol, double tol2, stoplist* res,
                               int64_t max_iter, int64_t* step_out)
{
    const int64_t N = s->N;
    const double k = s->par.k1, kf = s->par.k_frame, mu = s->par.mu;
    const int two_d = s->par.interactions == ORC_INT_LAPLACE2D;
    const double nneigh = two_d ? 4.0 : 2.0;
    s->qs_first = s->inc;
    s->qs_last = s->inc;
    int64_t step;
    for (step = 1; step < max_iter + 1; ++step) {
        memcpy(s->v_n, s->u, (size_t)N * sizeof(double));
        const double* old = s->v_n;
        for (int64_t p = 0; p < N; ++p) {
            double uneigh;
            if (!two_d) {
                if (p == 0) {
                    uneigh = old[N - 1] + old[N > 1 ? 1 : 0];
                }
                else if (p == N - 1) {
                    uneigh = old[N - 2] + old[0];
                }
                else {
                    uneigh = old[p - 1] + old[p + 1];
                }
            }
            else {
                int64_t i = p / s->cols, j = p % s->cols;
                uneigh = old[wrap(i - 1, s->rows) * s->cols + j] +
                         old[wrap(i + 1, s->rows) * s->cols + j] +
                         old[i * s->cols + wrap(j - 1, s->cols)] +
                         old[i * s->cols + wrap(j + 1, s->cols)];
            }
            block_t* b = &s->blk[p];
            int64_t i = b->i;
            double umin, u;
            for (;;) {
                while (i + 1 >= b->base + b->len) {
                    blk_append(s, b);
                }
                umin = 0.5 * (blk_y(b, i) + blk_y(b, i + 1));
                u = (k * uneigh + kf * s->u_frame + mu * umin) / (nneigh * k + kf + mu);
                b->i = i;
                if (!blk_align(s, b, u)) {
                    return check_landscape(s);
                }
                if (b->i == i) {
                    break;
                }
                i = b->i;
            }
            s->u[p] = u;
            s->f_frame[p] = kf * (s->u_frame - u);
            s->f_pot[p] = mu * (umin - u);
        }
        force_interactions(s);
        for (int64_t p = 0; p < N; ++p) {
            s->f[p] = s->f_pot[p] + s->f_int[p] + s->f_frame[p];
        }
        sl_roll_insert(res, orc_residual(s));
        if (sl_stop(res, tol, tol2)) {
            orc_quench(s);
            *step_out = 0;
            return ORC_OK;
        }
    }
    *step_out = step;
    return ORC_OK;
}

int orc_minimise(orc_system* s, double tol, int64_t niter_tol, int64_t max_iter,
                 int time_activity, int max_iter_is_error, int64_t* ret) /* detail.h:1676-1792 */
{
    if (!(tol < 1.0)) {
        return fail(ORC_EASSERT, "assertion failed (tol < 1.0)");
    }
    double tol2 = tol * tol;
    stoplist res;
    sl_init(&res, niter_tol);
    int64_t step = 0;
    int rc = ORC_OK;

    if (s->par.minimisation == ORC_MIN_NONE) { /* detail.h:1691-1693 */
        free(res.r);
        return fail(ORC_EUNSUPPORTED, "Minimisation not implementated");
    }
    if (s->par.minimisation == ORC_MIN_OVERDAMPED) {
        if (time_activity) {
            free(res.r);
            return fail(ORC_EASSERT, "assertion failed (!time_activity)");
        }
        if (s->par.potential != ORC_POT_CUSPY || (s->par.interactions != ORC_INT_LAPLACE1D &&
                                                  s->par.interactions != ORC_INT_LAPLACE2D)) {
            free(res.r);
            return fail(ORC_EUNSUPPORTED, "Minimisation not implementated");
        }
        rc = minimise_overdamped(s, tol, tol2, &res, max_iter, &step);
    }
    else {
        int64_t* i_n = NULL;
        int64_t sum = 0, sum_n = 0;
        int init = 1;
        if (time_activity) {
            i_n = (int64_t*)malloc((size_t)s->N * sizeof(int64_t));
            orc_chunk_index_at_align(s, i_n);
        }
        for (step = 1; step < max_iter + 1; ++step) {
            rc = time_step(s);
            if (rc) {
                break;
            }
            sl_roll_insert(&res, orc_residual(s));
            if (time_activity) {
                sum = sum_abs_index_diff(s, i_n, NULL);
                if (sum != sum_n) {
                    if (init) {
                        init = 0;
                        s->qs_first = s->inc;
                    }
                    s->qs_last = s->inc;
                }
                sum_n = sum;
            }
            if (sl_stop(&res, tol, tol2)) {
                orc_quench(s);
                step = 0;
                break;
            }
        }
        free(i_n);
    }
    free(res.r);
    *ret = step;
    if (rc) {
        return rc;
    }
    if (step != 0 && max_iter_is_error) {
        return fail(ORC_ENOCONV, "No convergence found");
    }
    return ORC_OK;
}

int orc_minimise_truncate(orc_system* s, const int64_t* i_n, int64_t A_truncate,
                          int64_t S_truncate, double tol, int64_t niter_tol, int64_t max_iter,
                          int time_activity, int max_iter_is_error,
                          int64_t* ret) /* detail.h:1833-1893 */
{
    if (!(tol < 1.0)) {
        return fail(ORC_EASSERT, "assertion failed (tol < 1.0)");
    }
    if (!time_activity) {
        return fail(ORC_EASSERT, "assertion failed (time_activity)");
    }
    if (s->par.minimisation != ORC_MIN_DYNAMIC) {
        return fail(ORC_EUNSUPPORTED, "Minimisation not implementated");
    }
    double tol2 = tol * tol;
    stoplist res;
    sl_init(&res, niter_tol);
    int64_t sum = 0, sum_n = 0, a = 0;
    int init = 1;
    int64_t step;
    int rc = ORC_OK;
    int truncated = 0;
    for (step = 1; step < max_iter + 1; ++step) {
        rc = time_step(s);
        if (rc) {
            break;
        }
        sl_roll_insert(&res, orc_residual(s));
        sum = sum_abs_index_diff(s, i_n, &a);
        if (sum != sum_n) {
            if (init) {
                init = 0;
                s->qs_first = s->inc;
            }
            s->qs_last = s->inc;
        }
        sum_n = sum;
        if (sl_stop(&res, tol, tol2)) {
            orc_quench(s);
            step = 0;
            truncated = 1;
            break;
        }
        if (A_truncate > 0 && a >= A_truncate) {
            truncated = 1;
            break;
        }
        if (S_truncate > 0 && sum >= S_truncate) {
            truncated = 1;
            break;
        }
    }
    free(res.r);
    *ret = step;
    if (rc) {
        return rc;
    }
    if (!truncated && max_iter_is_error) {
        return fail(ORC_ENOCONV, "No convergence found");
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * event driven protocol
 * ---------------------------------------------------------------------------------------- */
int orc_max_uniform_displacement(orc_system* s, int direction, double* out)
{
    /* detail.h:1901-1905 + Cuspy 176-187, SemiSmooth 282-334, Smooth 414-422 */
    if (direction != 1 && direction != -1) {
        return fail(ORC_EASSERT, "assertion failed (direction == 1 || direction == -1)");
    }
    const int64_t N = s->N;
    if (s->par.potential == ORC_POT_SMOOTH) {
        return fail(ORC_EUNSUPPORTED, "Operation not possible.");
    }
    align_all(s);
    double best = INFINITY;
    if (s->par.potential == ORC_POT_CUSPY) {
        for (int64_t p = 0; p < N; ++p) {
            const block_t* b = &s->blk[p];
            double d = direction > 0 ? blk_y(b, b->i + 1) - s->u[p] : s->u[p] - blk_y(b, b->i);
            if (d < best) {
                best = d;
            }
        }
        *out = best;
        return ORC_OK;
    }
    const double mu = s->par.mu, kappa = s->par.kappa;
    for (int64_t p = 0; p < N; ++p) {
        const block_t* b = &s->blk[p];
        double y0 = blk_y(b, b->i), y1 = blk_y(b, b->i + 1);
        double xi = 0.5 * (y0 + y1);
        double u_r = (mu * xi + kappa * y1) / (mu + kappa);
        double u_l = (mu * xi + kappa * y0) / (mu + kappa);
        double up = s->u[p];
        if (up < u_l) {
            *out = 0.0;
            return ORC_OK;
        }
        else if (up <= u_r) {
            double d = direction > 0 ? u_r - up : up - u_l;
            if (d < best) {
                best = d;
            }
        }
        else {
            *out = 0.0;
            return ORC_OK;
        }
    }
    *out = best;
    return ORC_OK;
}

/* detail.h:2027-2050 */
static double advance_uniformly(orc_system* s, double du, int input_is_frame)
{
    double du_particles, du_frame;
    const double kf = s->par.k_frame, mu = s->par.mu;
    if (input_is_frame) {
        du_frame = du;
        du_particles = du * kf / (kf + mu);
    }
    else {
        du_particles = du;
        du_frame = du * (kf + mu) / kf;
    }
    for (int64_t p = 0; p < s->N; ++p) {
        s->u[p] += du_particles;
    }
    s->u_frame += du_frame;
    updated_u(s);
    return input_is_frame ? du_particles : du_frame;
}

int orc_event_driven_step(orc_system* s, double eps, int kick, int direction,
                          double* out) /* detail.h:1933-1960 */
{
    if (direction != 1 && direction != -1) {
        return fail(ORC_EASSERT, "assertion failed (direction == 1 || direction == -1)");
    }
    double du = 0.0;
    if (!kick) {
        int rc = orc_max_uniform_displacement(s, direction, &du);
        if (rc) {
            return rc;
        }
        if (du < 0.5 * eps) {
            *out = 0.0;
            return ORC_OK;
        }
        *out = advance_uniformly(s, direction > 0 ? du - 0.5 * eps : 0.5 * eps - du, 0);
    }
    else {
        *out = advance_uniformly(s, direction > 0 ? eps : -eps, 0);
    }
    return check_landscape(s);
}

int orc_trigger(orc_system* s, int64_t p, double eps, int direction)
{
    /* detail.h:1972-1977 returns before updated_u();
     * Cuspy::trigger aligns first (193-205), SemiSmooth/Smooth do not (340-350, 428-438) */
    if (p < 0 || p >= s->N) {
        return fail(ORC_EASSERT, "assertion failed ((size_type)p < m_N)");
    }
    if (s->par.potential == ORC_POT_CUSPY) {
        align_all(s);
    }
    const block_t* b = &s->blk[p];
    if (direction > 0) {
        s->u[p] = blk_y(b, b->i + 1) + 0.5 * eps;
    }
    else {
        s->u[p] = blk_y(b, b->i) - 0.5 * eps;
    }
    return check_landscape(s);
}

int orc_advance_to_fixed_force(orc_system* s, double f_frame,
                               int allow_plastic) /* detail.h:1988-1995 */
{
    int64_t* i_n = (int64_t*)malloc((size_t)s->N * sizeof(int64_t));
    orc_chunk_index_at_align(s, i_n);
    double mean = 0.0;
    for (int64_t p = 0; p < s->N; ++p) {
        mean += s->f_frame[p];
    }
    mean /= (double)s->N;
    advance_uniformly(s, (f_frame - mean) / s->par.mu, 0);
    int changed = any_index_changed(s, i_n);
    free(i_n);
    if (!allow_plastic && changed) {
        return fail(ORC_EASSERT,
                    "assertion failed (allow_plastic || xt::all(xt::equal(m_chunk->index_at_align(), i_n)))");
    }
    return check_landscape(s);
}

/* ------------------------------------------------------------------------------------------
 * chunk surface
 * ---------------------------------------------------------------------------------------- */
int orc_chunk_index_at_align(const orc_system* s, int64_t* out)
{
    for (int64_t p = 0; p < s->N; ++p) {
        out[p] = s->blk[p].i;
    }
    return ORC_OK;
}

int orc_chunk_left_of_align(const orc_system* s, double* out)
{
    for (int64_t p = 0; p < s->N; ++p) {
        out[p] = blk_y(&s->blk[p], s->blk[p].i);
    }
    return ORC_OK;
}

int orc_chunk_right_of_align(const orc_system* s, double* out)
{
    for (int64_t p = 0; p < s->N; ++p) {
        out[p] = blk_y(&s->blk[p], s->blk[p].i + 1);
    }
    return ORC_OK;
}

int orc_chunk_yield(orc_system* s, int64_t first, int64_t n, double* out)
{
    if (first < 0 || n < 1) {
        return fail(ORC_EASSERT, "yield window out of range");
    }
    /* generated on the fly from the stored stretch; the block itself is left untouched */
    for (int64_t p = 0; p < s->N; ++p) {
        const block_t* b = &s->blk[p];
        double* o = out + p * n;
        const int64_t lo = b->base, hi = b->base + b->len - 1, last = first + n - 1;
        for (int64_t j = (first > lo ? first : lo); j <= (last < hi ? last : hi); ++j) {
            o[j - first] = blk_y(b, j);
        }
        if (last > hi) {
            uint64_t st = b->st_end;
            double y = blk_y(b, hi);
            for (int64_t j = hi + 1; j <= last; ++j) {
                y = y + blk_draw_fwd(s, &st);
                if (j >= first) {
                    o[j - first] = y;
                }
            }
        }
        if (first < lo) {
            uint64_t st = b->st_begin;
            double y = blk_y(b, lo);
            for (int64_t j = lo - 1; j >= first; --j) {
                y = y - blk_peek(s, st); /* y[j] = y[j+1] - d_{j+1} */
                if (s->consumes) {
                    st = pcg_prev(st, s->inc_rng, s->minv);
                }
                if (j <= last) {
                    o[j - first] = y;
                }
            }
        }
    }
    return ORC_OK;
}

int orc_chunk_state_at(orc_system* s, const int64_t* index, uint64_t* state)
{
    for (int64_t p = 0; p < s->N; ++p) {
        const block_t* b = &s->blk[p];
        state[p] = s->consumes ? pcg_advance(b->st_begin, s->inc_rng, index[p] - b->base)
                               : b->st_begin;
    }
    return ORC_OK;
}

int orc_chunk_restore(orc_system* s, const uint64_t* state, const double* value,
                      const int64_t* index)
{
    for (int64_t p = 0; p < s->N; ++p) {
        block_t* b = &s->blk[p];
        b->len = 0;
        b->head = 0;
        blk_reserve(b, 0, 1);
        b->base = index[p];
        b->st_begin = state[p];
        b->st_end = state[p];
        (void)blk_draw_fwd(s, &b->st_end); /* d_index is implied by `value` */
        b->y[b->head] = value[p];
        b->len = 1;
        b->i = index[p];
    }
    updated_u(s); /* re-align to the current u */
    return check_landscape(s);
}

/* ------------------------------------------------------------------------------------------
 * bounded multi-threaded CPU arm (bench.py cpu_baseline / --impl reference):
 * `nsys` independent lines (seed = seed0 + r*N), one realisation per thread round-robin,
 * prepared by minimise() + eventDrivenStep(kick) and then advanced by timeSteps(n).
 * ---------------------------------------------------------------------------------------- */
struct orc_ensemble {
    orc_params par;
    int64_t nsys;
    int nthreads;
    int fused; /* timeSteps through orc_time_steps_fused */
    int ftz;   /* flush denormals (x86 MXCSR FTZ | DAZ) in the worker threads: a long un-kicked
                  run decays into denormal velocities, which x86 executes several times slower;
                  timing only -- results then differ from the reference below 1e-308 */
    orc_system** sys;
};

typedef struct {
    orc_ensemble* e;
    int tid;
    int64_t nsteps;
    int prepare; /* 1: create + minimise + kick, 2: kick only, 0: timeSteps */
    double checksum;
} ens_arg;

static void* ens_worker(void* vp)
{
    ens_arg* a = (ens_arg*)vp;
    orc_ensemble* e = a->e;
    double cs = 0.0;
#if defined(__x86_64__) || defined(__i386__)
    if (e->ftz) {
        __builtin_ia32_ldmxcsr(__builtin_ia32_stmxcsr() | 0x8040u);
    }
#endif
    for (int64_t r = a->tid; r < e->nsys; r += e->nthreads) {
        if (a->prepare == 2) {
            if (e->sys[r]) {
                double du;
                orc_event_driven_step(e->sys[r], 1e-3, 0, 1, &du);
                orc_event_driven_step(e->sys[r], 1e-3, 1, 1, &du);
            }
        }
        else if (a->prepare) {
            orc_params par = e->par;
            int64_t n = par.shape[0] * (par.rank == 2 ? par.shape[1] : 1);
            par.seed = e->par.seed + (uint64_t)(r * n);
            orc_system* s = NULL;
            if (orc_create(&par, &s) != ORC_OK) {
                continue;
            }
            int64_t ret;
            double du;
            orc_minimise(s, 1e-5, 10, 1000000000LL, 0, 0, &ret);
            orc_event_driven_step(s, 1e-3, 0, 1, &du);
            orc_event_driven_step(s, 1e-3, 1, 1, &du);
            e->sys[r] = s;
        }
        else if (e->sys[r]) {
            orc_system* s = e->sys[r];
            if (e->fused) {
                orc_time_steps_fused(s, a->nsteps);
            }
            else {
                orc_time_steps(s, a->nsteps);
            }
            for (int64_t p = 0; p < s->N; ++p) {
                cs += s->u[p];
            }
        }
    }
    a->checksum = cs;
    return NULL;
}

static double ens_run(orc_ensemble* e, int prepare, int64_t nsteps, double* checksum)
{
    pthread_t* th = (pthread_t*)malloc((size_t)e->nthreads * sizeof(pthread_t));
    ens_arg* args = (ens_arg*)calloc((size_t)e->nthreads, sizeof(ens_arg));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < e->nthreads; ++t) {
        args[t] = (ens_arg){e, t, nsteps, prepare, 0.0};
        pthread_create(&th[t], NULL, ens_worker, &args[t]);
    }
    double cs = 0.0;
    for (int t = 0; t < e->nthreads; ++t) {
        pthread_join(th[t], NULL);
        cs += args[t].checksum;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    free(args);
    if (checksum) {
        *checksum = cs;
    }
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

orc_ensemble* orc_ensemble_create(const orc_params* par, int64_t nsys, int nthreads)
{
    orc_ensemble* e = (orc_ensemble*)calloc(1, sizeof *e);
    e->par = *par;
    e->nsys = nsys;
    e->nthreads = nthreads < 1 ? 1 : nthreads;
    e->sys = (orc_system**)calloc((size_t)nsys, sizeof(orc_system*));
    ens_run(e, 1, 0, NULL);
    return e;
}

double orc_ensemble_time_steps(orc_ensemble* e, int64_t nsteps, double* checksum)
{
    return ens_run(e, 0, nsteps, checksum);
}

void orc_ensemble_configure(orc_ensemble* e, int fused, int ftz)
{
    e->fused = fused;
    e->ftz = ftz;
}

/* eventDrivenStep(eps, false) + eventDrivenStep(eps, true) on every system (untimed by callers) */
void orc_ensemble_kick(orc_ensemble* e) { ens_run(e, 2, 0, NULL); }

void orc_ensemble_destroy(orc_ensemble* e)
{
    if (!e) {
        return;
    }
    for (int64_t r = 0; r < e->nsys; ++r) {
        orc_destroy(e->sys[r]);
    }
    free(e->sys);
    free(e);
}
