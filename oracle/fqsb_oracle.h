/*
 * fqsb_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE).
 *
 * A dependency-free C restatement of the hot path of tdegeus/FrictionQPotSpringBlock
 * (include/FrictionQPotSpringBlock/detail.h, Line1d.h, Line2d.h) plus the behaviour of the
 * two absent third-party libraries it calls on that path:
 *   - prrng  (tdegeus/prrng, pinned ">=1.11.1" by the reference's environment.yaml:
 *             pcg32, pcg32_tensor_cumsum, lower_bound, alignment) and
 *   - GooseFEM::Iterate::StopList (tdegeus/GooseFEM, unpinned).
 * Their sources are NOT under /root/reference, so their published algorithms are restated
 * here and the restatement is PINNED on the reference's own committed
 * golden files examples/ *.h5 (tests/test_oracle_golden.py: S integer-exact, x_frame and
 * f_frame np.allclose, as asserted by examples/Line1d_Cuspy_Laplace.py:64-67).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use this library, and only as the checker / the timed CPU arm. The product
 * (frictionqpotspringblock_b200) never links, imports or calls it.
 *
 * Arithmetic: every per-block expression keeps the reference's evaluation order and is
 * compiled with -ffp-contract=off, so the values are a deterministic IEEE-754 function of
 * the inputs; the array-pass structure of detail.h (one loop per xtensor assignment) is
 * kept too, because it is also what bench.py times as the CPU baseline.
 */
#ifndef FQSB_ORACLE_H
#define FQSB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* potential (detail.h:113-439) */
enum { ORC_POT_CUSPY = 0, ORC_POT_SEMISMOOTH = 1, ORC_POT_SMOOTH = 2 };
/* interactions (detail.h:446-868); NONE = Particles-like */
enum {
    ORC_INT_NONE = 0,
    ORC_INT_LAPLACE1D = 1,
    ORC_INT_QUARTIC1D = 2,
    ORC_INT_QUARTICGRADIENT1D = 3,
    ORC_INT_LONGRANGE1D = 4,
    ORC_INT_LAPLACE2D = 5,
    ORC_INT_QUARTICGRADIENT2D = 6
};
/* minimisation (detail.h:1005-1020, 1691-1753) */
enum { ORC_MIN_DYNAMIC = 0, ORC_MIN_OVERDAMPED = 1, ORC_MIN_NONE = 2 };
/* prrng::distribution (detail.h:31-66) */
enum {
    ORC_DIST_RANDOM = 0,
    ORC_DIST_DELTA = 1,
    ORC_DIST_EXPONENTIAL = 2,
    ORC_DIST_POWER = 3,
    ORC_DIST_GAMMA = 4,
    ORC_DIST_PARETO = 5,
    ORC_DIST_WEIBULL = 6,
    ORC_DIST_NORMAL = 7
};

/* return codes */
enum { ORC_OK = 0, ORC_ENAN = 1, ORC_ENOCONV = 2, ORC_EASSERT = 3, ORC_EUNSUPPORTED = 4 };

typedef struct {
    int32_t potential;
    int32_t interactions;
    int32_t minimisation;
    int32_t rank;        /* 1 or 2 */
    int64_t shape[2];    /* [N,1] or [rows,cols] */
    double m, eta, mu, kappa;
    double k1;           /* k_interactions | a1 | k2 */
    double k2;           /* a2 | k4 | alpha */
    double k_frame, dt;
    uint64_t seed;
    int32_t distribution;
    int32_t nparameters;
    double parameters[4];
    double offset;
    int64_t nchunk;
} orc_params;

typedef struct orc_system orc_system;

const char* orc_last_error(void);

int orc_create(const orc_params* par, orc_system** out);
void orc_destroy(orc_system* s);

int64_t orc_size(const orc_system* s);

/* setters: detail.h:1231-1315 */
int orc_set_u(orc_system* s, const double* u);
int orc_set_v(orc_system* s, const double* v);
int orc_set_a(orc_system* s, const double* a);
int orc_set_u_frame(orc_system* s, double u_frame);
int orc_set_inc(orc_system* s, int64_t inc);
int orc_set_t(orc_system* s, double t);
int orc_refresh(orc_system* s);
int orc_quench(orc_system* s);

/* getters: detail.h:1402-1520. which: 0 u,1 v,2 a,3 f,4 f_potential,5 f_frame,
 * 6 f_interactions,7 f_damping */
int orc_get(const orc_system* s, int which, double* out);
double orc_u_frame(const orc_system* s);
int64_t orc_inc(const orc_system* s);
double orc_residual(const orc_system* s);
double orc_temperature(const orc_system* s);
int64_t orc_qs_first(const orc_system* s);
int64_t orc_qs_last(const orc_system* s);

/* dynamics: detail.h:1539-1645 */
int orc_time_steps(orc_system* s, int64_t n);
/* the fused single-pass CPU flavour of the same steps (bit-identical results) */
int orc_time_steps_fused(orc_system* s, int64_t n);
int orc_flow_steps(orc_system* s, int64_t n, double v_frame);
int orc_time_steps_until_event(orc_system* s, double tol, int64_t niter_tol, int64_t max_iter,
                               int64_t* ret);
/* detail.h:1676-1893 */
int orc_minimise(orc_system* s, double tol, int64_t niter_tol, int64_t max_iter,
                 int time_activity, int max_iter_is_error, int64_t* ret);
int orc_minimise_truncate(orc_system* s, const int64_t* i_n, int64_t A_truncate,
                          int64_t S_truncate, double tol, int64_t niter_tol, int64_t max_iter,
                          int time_activity, int max_iter_is_error, int64_t* ret);
/* detail.h:1901-2050 */
int orc_max_uniform_displacement(orc_system* s, int direction, double* out);
int orc_event_driven_step(orc_system* s, double eps, int kick, int direction, double* out);
int orc_trigger(orc_system* s, int64_t p, double eps, int direction);
int orc_advance_to_fixed_force(orc_system* s, double f_frame, int allow_plastic);

/* prrng pcg32_tensor_cumsum surface */
int orc_chunk_index_at_align(const orc_system* s, int64_t* out);
int orc_chunk_left_of_align(const orc_system* s, double* out);
int orc_chunk_right_of_align(const orc_system* s, double* out);
/* y[p, j] for global indices j = first .. first+n-1 (first >= 0), row-major [N, n] */
int orc_chunk_yield(orc_system* s, int64_t first, int64_t n, double* out);
/* state of block p's generator positioned so that its next draw is global draw index[p] */
int orc_chunk_state_at(orc_system* s, const int64_t* index, uint64_t* state);
/* restart every block's sequence: y[index[p]] = value[p], generator state[p] as state_at */
int orc_chunk_restore(orc_system* s, const uint64_t* state, const double* value,
                      const int64_t* index);

/* External = RandomNormalForcing (detail.h:881-1000; Line1d.h:261-330, 486-556;
 * Particles.h System_Cuspy_RandomForcing): the system of `par` plus a normally distributed force
 * per block that is redrawn from ONE sequential pcg32(seed_forcing) stream, in block order,
 * whenever inc >= next[p] (next starts at dinc_init and advances by dinc[p]). */
int orc_create_thermal(const orc_params* par, double mean, double stddev, uint64_t seed_forcing,
                       const int64_t* dinc_init, const int64_t* dinc, orc_system** out);
/* external.f_thermal / external.next / external.state (python/main.cpp:255-268); NULL = skip */
int orc_thermal_get(const orc_system* s, double* f_thermal, int64_t* next, uint64_t* state);
int orc_thermal_set(orc_system* s, const double* f_thermal, const int64_t* next,
                    const uint64_t* state);

/* free-standing helpers used by the tests */
void orc_pcg32_normal(uint64_t initstate, uint64_t initseq, int64_t n, double mean,
                      double stddev, double* out);
void orc_pcg32_randint(uint64_t initstate, uint64_t initseq, int64_t n, uint32_t high,
                       int64_t* out);
double orc_erf_inv(double z);
void orc_pcg32_draws(uint64_t initstate, uint64_t initseq, int64_t n, double* out);
double orc_draw_to_spacing(double r, int32_t distribution, const double* par);
double orc_gamma_p_inv(double a, double p);

/* bounded multi-threaded CPU arm for bench.py: `nsys` independent lines (seed = seed0 + r*N),
 * one realisation per thread round-robin; create() prepares them with minimise() and one
 * eventDrivenStep kick (untimed); time_steps() advances all of them by timeSteps(nsteps) and
 * returns the wall seconds that took. */
typedef struct orc_ensemble orc_ensemble;
orc_ensemble* orc_ensemble_create(const orc_params* par, int64_t nsys, int nthreads);
double orc_ensemble_time_steps(orc_ensemble* e, int64_t nsteps, double* checksum);
void orc_ensemble_kick(orc_ensemble* e);
/* fused: timeSteps as the single-pass flavour; ftz: flush denormals in the workers (timing only) */
void orc_ensemble_configure(orc_ensemble* e, int fused, int ftz);
void orc_ensemble_destroy(orc_ensemble* e);

#ifdef __cplusplus
}
#endif
#endif
