/*
 * fqsb.h -- C ABI of the B200-native FrictionQPotSpringBlock integrator (libfqsb.so).
 *
 * This is the drop-in boundary of the hot path. The reference has no FFI of its own: it is a
 * header-only C++ template library (include/FrictionQPotSpringBlock/detail.h) consumed
 * directly by python/main.cpp. Each entry point below therefore names the reference
 * member function it replaces ("ref:" = path:line under the reference tree); a C++ host class
 * with the reference's names (include/fqsb.hpp) and the Python module
 * (frictionqpotspringblock_b200) forward 1:1 to these calls, exactly as python/main.cpp:46-226
 * forwards to detail::System.
 *
 * One handle = an ENSEMBLE of `nrealisations` >= 1 independent systems of identical
 * parameters (a single reference `System_*` object is nrealisations == 1). Realisation r owns
 * the pcg32 initstates  seed + r*seed_stride + p  (p = flat block index), so it equals the
 * reference object constructed with seed = seed + r*seed_stride (ref: Line1d.h:148-151).
 * Per-block arrays are row-major [nrealisations][size]; per-realisation scalars are arrays of
 * length nrealisations.
 *
 * Conventions
 *  - plain pointers are HOST pointers unless the name says `_device`;
 *  - every function returns an fqsb_status; the message of the last failure on the calling
 *    thread is fqsb_last_error() and equals the reference's exception text;
 *  - one handle = one owner thread = one CUDA stream; calls are synchronous on return;
 *  - there is no CPU fallback: without a CUDA device fqsb_create fails with FQSB_ECUDA.
 */
#ifndef FQSB_H
#define FQSB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FQSB_ABI_VERSION 4

typedef enum {
    FQSB_OK = 0,
    FQSB_ENAN = 1,         /* "NaN entries found"               ref: detail.h:1567-1569 */
    FQSB_ENOCONV = 2,      /* "No convergence found"            ref: detail.h:1787-1789 */
    FQSB_EASSERT = 3,      /* "...assertion failed (...)"       ref: config.h:19-24     */
    FQSB_EUNSUPPORTED = 4, /* "Minimisation not implementated", "Operation not possible." */
    FQSB_ECUDA = 5         /* CUDA runtime failure / no device */
} fqsb_status;

/* ref: detail.h:113-439 */
typedef enum { FQSB_POT_CUSPY = 0, FQSB_POT_SEMISMOOTH = 1, FQSB_POT_SMOOTH = 2 } fqsb_potential;

/* ref: detail.h:446-868 */
typedef enum {
    FQSB_INT_NONE = 0,
    FQSB_INT_LAPLACE1D = 1,
    FQSB_INT_QUARTIC1D = 2,
    FQSB_INT_QUARTICGRADIENT1D = 3,
    FQSB_INT_LONGRANGE1D = 4,
    FQSB_INT_LAPLACE2D = 5,
    FQSB_INT_QUARTICGRADIENT2D = 6
} fqsb_interactions;

/* ref: detail.h:1005-1020 (Overdamped = the "minimise_nopassing" of older releases) */
typedef enum {
    FQSB_MIN_DYNAMIC = 0,
    FQSB_MIN_OVERDAMPED = 1,
    FQSB_MIN_NONE = 2 /* thermal systems: minimise throws, ref: detail.h:1691-1693 */
} fqsb_minimisation;

/* ref: detail.h:31-66 (prrng::distribution) */
typedef enum {
    FQSB_DIST_RANDOM = 0,
    FQSB_DIST_DELTA = 1,
    FQSB_DIST_EXPONENTIAL = 2,
    FQSB_DIST_POWER = 3,
    FQSB_DIST_GAMMA = 4, /* (k, theta, offset): theta * gamma_p_inv(k, r) + offset; own solver */
    FQSB_DIST_PARETO = 5,
    FQSB_DIST_WEIBULL = 6,
    FQSB_DIST_NORMAL = 7 /* (mu, sigma, offset): mu + sigma*sqrt(2)*erf_inv(2r - 1) + offset */
} fqsb_distribution;

/* Constructor arguments of every Line1d / Line2d System_* class
 * (ref: Line1d.h:134-147,199-211,348-362,394-407,451-465,585-599,643-657; Line2d.h:88-101,133-147) */
typedef struct {
    int32_t potential;    /* fqsb_potential */
    int32_t interactions; /* fqsb_interactions */
    int32_t minimisation; /* fqsb_minimisation */
    int32_t rank;         /* 1 (Line1d) or 2 (Line2d) */
    int64_t shape[2];     /* [N, 1] or [rows, cols] */
    double m, eta, mu, kappa;
    double k1; /* k_interactions | a1 | k2 */
    double k2; /* a2 | k4 | alpha */
    double k_frame, dt;
    uint64_t seed;
    int32_t distribution; /* fqsb_distribution */
    int32_t nparameters;
    double parameters[4];
    double offset;  /* default -100 in the reference */
    int64_t nchunk; /* default 5000; only sizes fqsb_chunk_data (the device keeps no chunk) */
    /* --- new surface (the reference has no ensemble class) --- */
    int64_t nrealisations; /* 0 is read as 1 */
    int64_t seed_stride;   /* 0 is read as prod(shape) */
    int32_t device;        /* CUDA ordinal; -1 = current device */
    int32_t kernel;        /* low 4 bits: 0 auto, 1 force resident (one CTA per realisation), 2 force
                              streaming (one step per launch), 3 force the temporally blocked tiles
                              (1-D nearest-neighbour lines; bits 8..15 steps per launch, bits 16..31
                              owned blocks per tile, 0 = planner's choice); bits 4..6 resident variant;
                              bit 7 (FQSB_KERNEL_FMA): opt in to the resident kernels built with FMA
                              contraction (Cuspy lines with Laplace / Quartic / QuarticGradient
                              interactions; ~1.4x fewer FP64 instructions per step). The yield
                              landscape stays exact; u, v, a then agree with the reference to
                              rounding (~1e-13 relative after 1000 steps) instead of bit for bit.
                              Default off: bit-identical arithmetic. */
    /* slab decomposition: local block p is global block (seed_first + p) mod seed_period and draws
     * the global block's pcg32 stream; seed_period == 0 disables the mapping */
    int64_t seed_first;
    int64_t seed_period;
} fqsb_params;

#define FQSB_KERNEL_FMA 0x80

typedef struct fqsb_system fqsb_system;

/* library ------------------------------------------------------------------------------- */
const char* fqsb_last_error(void);
int fqsb_abi_version(void);
const char* fqsb_version(void); /* ref: config.h:209-212 */
int fqsb_device_count(void);

/* lifetime (ref: Line1d.h:134-161 ctor body + detail.h:1096-1139 initSystem) ------------- */
int fqsb_create(const fqsb_params* params, fqsb_system** out);
void fqsb_destroy(fqsb_system* s);
int fqsb_get_params(const fqsb_system* s, fqsb_params* out);
int64_t fqsb_size(const fqsb_system* s);          /* blocks per realisation, ref: detail.h:1155 */
int64_t fqsb_nrealisations(const fqsb_system* s);
/* run all later work of this handle on an existing cudaStream_t (e.g. torch's current one) */
int fqsb_set_stream(fqsb_system* s, void* cuda_stream);
void* fqsb_get_stream(const fqsb_system* s);

/* state in (ref: detail.h:1231-1315) ---------------------------------------------------- */
int fqsb_set_u(fqsb_system* s, const double* u, int64_t n);  /* + updated_u(),  ref: 1276-1281 */
int fqsb_set_v(fqsb_system* s, const double* v, int64_t n);  /* + updated_v(),  ref: 1290-1295 */
int fqsb_set_a(fqsb_system* s, const double* a, int64_t n);  /*                 ref: 1301-1305 */
int fqsb_set_u_frame(fqsb_system* s, const double* u_frame); /* [R]             ref: 1253-1258 */
int fqsb_set_inc(fqsb_system* s, const int64_t* inc);        /* [R]             ref: 1241-1247 */
int fqsb_set_t(fqsb_system* s, const double* t);             /* [R]             ref: 1231-1235 */
int fqsb_refresh(fqsb_system* s);                            /*                 ref: 1310-1315 */
int fqsb_quench(fqsb_system* s);                             /*                 ref: 1527-1532 */

/* state out (ref: detail.h:1402-1520) --------------------------------------------------- */
typedef enum {
    FQSB_U = 0,
    FQSB_V = 1,
    FQSB_A = 2,
    FQSB_F = 3,
    FQSB_F_POTENTIAL = 4,
    FQSB_F_FRAME = 5,
    FQSB_F_INTERACTIONS = 6,
    FQSB_F_DAMPING = 7
} fqsb_array;
int fqsb_get(fqsb_system* s, int which, double* out, int64_t n); /* ref: 1402-1468 */
/* device-resident view of the same arrays (valid until the next call on the handle) */
int fqsb_get_device(fqsb_system* s, int which, const double** out_device);
int fqsb_get_u_frame(fqsb_system* s, double* out);     /* [R]  ref: 1264-1267 */
int fqsb_get_inc(fqsb_system* s, int64_t* out);        /* [R]  ref: 1486-1489 */
int fqsb_get_t(fqsb_system* s, double* out);           /* [R]  ref: 1477-1480 */
int fqsb_residual(fqsb_system* s, double* out);        /* [R]  ref: 1512-1520 */
int fqsb_temperature(fqsb_system* s, double* out);     /* [R]  ref: 1500-1503 */
int fqsb_mean_f_frame(fqsb_system* s, double* out);    /* [R]  np.mean(system.f_frame) of the examples */
int fqsb_qs_activity(fqsb_system* s, int64_t* first, int64_t* last); /* [R] ref: 1802-1818 */

/* dynamics (ref: detail.h:1539-1645) ---------------------------------------------------- */
int fqsb_time_steps(fqsb_system* s, int64_t n);                  /* timeStep/timeSteps, ref: 1539-1583 */
int fqsb_flow_steps(fqsb_system* s, int64_t n, double v_frame);  /* ref: 1637-1645 */
/* ret [R]: step of the first well change, 0 if converged, max_iter+1 otherwise. ref: 1595-1622 */
int fqsb_time_steps_until_event(fqsb_system* s, double tol, int64_t niter_tol, int64_t max_iter,
                                int64_t* ret);

/* `system.u = u; system.v = v; system.a = a; system.timeSteps(nsteps); out = system.u ...` as ONE
 * call whose host<->device copies overlap the kernels: the realisations are cut into chunks that
 * run on internal streams of the handle (copy-in, updated_u(), resident kernel, copy-out per
 * chunk). Same results as the separate calls. u, v, a [R*size] host (pinned for real overlap) or
 * NULL = keep; out_u, out_v, out_a [R*size] or NULL; out_mean_f_frame [R] or NULL.
 * ref: detail.h:1276-1305 (setters), 1577-1583 (timeSteps), 1402-1468 (getters) */
int fqsb_run_from_host(fqsb_system* s, const double* u, const double* v, const double* a, int64_t n,
                       int64_t nsteps, double* out_u, double* out_v, double* out_a,
                       double* out_mean_f_frame);

/* minimisation (ref: detail.h:1676-1893) ------------------------------------------------ */
/* ret [R]: 0 if converged, max_iter+1 otherwise (only reachable with max_iter_is_error == 0).
 * Dynamic systems: velocity-Verlet until the StopList criterion (ref: 1754-1785);
 * Overdamped systems: no-passing Jacobi sweeps (ref: 1694-1753). */
int fqsb_minimise(fqsb_system* s, double tol, int64_t niter_tol, int64_t max_iter,
                  int time_activity, int max_iter_is_error, int64_t* ret);
/* i_n [R][size]; ret [R]. ref: 1833-1893 */
int fqsb_minimise_truncate(fqsb_system* s, const int64_t* i_n, int64_t A_truncate,
                           int64_t S_truncate, double tol, int64_t niter_tol, int64_t max_iter,
                           int time_activity, int max_iter_is_error, int64_t* ret);

/* event-driven protocol (ref: detail.h:1901-2050) --------------------------------------- */
int fqsb_max_uniform_displacement(fqsb_system* s, int direction, double* out); /* [R] ref: 1901-1905 */
int fqsb_event_driven_step(fqsb_system* s, double eps, int kick, int direction,
                           double* du_frame);                                   /* [R] ref: 1933-1960 */
int fqsb_trigger(fqsb_system* s, int64_t realisation, int64_t p, double eps, int direction); /* ref: 1972-1977 */
int fqsb_advance_to_fixed_force(fqsb_system* s, const double* f_frame, int allow_plastic);   /* [R] ref: 1988-1995 */

/* yield landscape = the prrng::pcg32_tensor_cumsum object `system.chunk`
 * (absent third-party; call sites ref: detail.h:144-160,183-186,1602,1609,1725-1733,
 *  tests/test_Line1d.py:83-86,309-326) -------------------------------------------------- */
int fqsb_chunk_index_at_align(fqsb_system* s, int64_t* out, int64_t n); /* global well index i */
int fqsb_chunk_left_of_align(fqsb_system* s, double* out, int64_t n);   /* y[i]   */
int fqsb_chunk_right_of_align(fqsb_system* s, double* out, int64_t n);  /* y[i+1] */
/* chunk.align(u): the wells follow the positions u [R*size] (the system's own slips are untouched;
 * the next updated_u() re-aligns to them) */
int fqsb_chunk_align(fqsb_system* s, const double* u, int64_t n);
/* y[p, first[p] + j], j < nyield, row-major [R*size][nyield]; first [R*size] >= 0 */
int fqsb_chunk_data(fqsb_system* s, const int64_t* first, int64_t nyield, double* out);
/* pcg32 state positioned so that the next draw is global draw index[p] */
int fqsb_chunk_state_at(fqsb_system* s, const int64_t* index, uint64_t* state, int64_t n);
/* restart: y[index[p]] = value[p] with generator state[p] (as from state_at), then re-align */
int fqsb_chunk_restore(fqsb_system* s, const uint64_t* state, const double* value,
                       const int64_t* index, int64_t n);
/* signed well-index change since `i_n` summed per realisation (the examples' S) and the number
 * of blocks that changed (A): i_n [R*size] (n = R*size is checked), out_S [R], out_A [R] (either
 * may be NULL). ref: examples/Line1d_Cuspy_Laplace.py:59-61 */
int fqsb_avalanche(fqsb_system* s, const int64_t* i_n, int64_t n, int64_t* out_S, int64_t* out_A);
/* the same without moving R*size indices over PCIe twice per event:
 * fqsb_mark_indices keeps a device-side copy of the current well indices (the `i_n =
 * system.chunk.index_at_align` of the examples); fqsb_avalanche_since_mark reduces S = sum(i - i_n)
 * and A = #(i != i_n) per realisation on the device and returns R x 16 bytes. */
int fqsb_mark_indices(fqsb_system* s);
int fqsb_avalanche_since_mark(fqsb_system* s, int64_t* out_S, int64_t* out_A);
/* per-realisation event record of the last minimise(time_activity = 1) / minimise_truncate call,
 * kept by the stepping kernel itself: S_abs = sum |i - i_n|, A = #(i != i_n) (ref:
 * detail.h:1768-1778, 1863-1864), first / last = quasistaticActivityFirst / Last (ref: 1802-1818).
 * Any pointer may be NULL. R x 32 bytes. */
int fqsb_event_record(fqsb_system* s, int64_t* S_abs, int64_t* A, int64_t* first, int64_t* last);

/* thermal systems: External = RandomNormalForcing (ref: detail.h:881-1000; classes
 * Line1d.h:261-330 System_Cuspy_Laplace_RandomForcing, Line1d.h:486-556
 * System_Cuspy_Quartic_RandomForcing, Particles.h System_Cuspy_RandomForcing) ------------------
 * Call once, right after fqsb_create (params.minimisation = FQSB_MIN_NONE): every block gets a
 * force drawn from normal(mean, stddev) that is redrawn whenever inc >= next[p] (next starts at
 * dinc_init[p] and then advances by dinc[p]); the draws of one realisation come, in block order,
 * from ONE sequential prrng::pcg32(seed_forcing + r*seed_forcing_stride) stream (stride 0 is read
 * as 1). dinc_init, dinc: [R*size]. The residual becomes f_frame + f_potential + f_interactions +
 * f_damping + f_thermal (ref: detail.h:1326-1329). */
int fqsb_enable_random_forcing(fqsb_system* s, double mean, double stddev, uint64_t seed_forcing,
                               int64_t seed_forcing_stride, const int64_t* dinc_init,
                               const int64_t* dinc, int64_t n);
/* `system.external` (ref: python/main.cpp:250-275) */
int fqsb_external_get_f_thermal(fqsb_system* s, double* out, int64_t n);      /* ref: 967-970 */
int fqsb_external_set_f_thermal(fqsb_system* s, const double* f, int64_t n);  /* ref: 976-980 */
int fqsb_external_get_next(fqsb_system* s, int64_t* out, int64_t n);          /* ref: 986-989 */
int fqsb_external_set_next(fqsb_system* s, const int64_t* next, int64_t n);   /* ref: 995-999 */
int fqsb_external_get_state(fqsb_system* s, uint64_t* out);                   /* [R] ref: 949-952 */
int fqsb_external_set_state(fqsb_system* s, const uint64_t* state);           /* [R] ref: 958-961 */

/* slab decomposition primitives (new surface; one very large line / interface over several GPUs,
 * driven by frictionqpotspringblock_b200/slab.py) ----------------------- */
/* only blocks [lo, hi) of the local array enter reductions (the rest are halo copies) */
int fqsb_set_owned_range(fqsb_system* s, int64_t lo, int64_t hi);
/* k steps (Verlet, or no-passing sweeps) without a stop decision; log [R][k][5] receives per step
 * {sum f^2, sum f_frame^2, #well changes, dS, dA} over the owned range */
int fqsb_logged_steps(fqsb_system* s, int64_t k, double* log);
int fqsb_snapshot(fqsb_system* s);
int fqsb_rollback(fqsb_system* s);
/* full state (u, v, a, y_l, y_r, idx, rng: 7 planes of 8-byte words) of `count` consecutive
 * blocks; `buf` is a device pointer when on_device != 0 */
int fqsb_export_cells(fqsb_system* s, int64_t first, int64_t count, void* buf, int on_device);
int fqsb_import_cells(fqsb_system* s, int64_t first, int64_t count, const void* buf, int on_device);
/* u += du[r], u_frame += du_frame[r], re-align (advanceUniformly, ref: detail.h:2027-2050) */
int fqsb_advance_uniformly(fqsb_system* s, const double* du, const double* du_frame);
/* raw sums over the owned range, out [R][4] (see fqsb_api.cu); i_n [n] is read for what == 4 */
int fqsb_reduce_sums(fqsb_system* s, int what, int direction, const int64_t* i_n, int64_t n,
                     double* out);

/* slab decomposition inside the library (BASELINE configs #3, #5) --------
 * ONE very large line / interface over G GPUs: member g (a handle with nrealisations == 1,
 * params.kernel = 2, params.seed_first / seed_period set) owns a contiguous range of rows and
 * integrates them extended by `halo` rows per side. Batches of k <= halo steps run without
 * communication; after a batch every member stores its outermost owned rows straight into its
 * neighbours' memory (NVLink peer stores + st.release.sys epoch flags, no NCCL, no host staging)
 * and its k x 5 log of per-step sums into every member's mailbox; the per-step stop decision of
 * the reference (ref: detail.h:1754-1785) is replayed on the rank-ordered global sums, identically
 * on every member. Members live in ONE process (`members` = all G handles, peer access between
 * the devices) or in one process per GPU (`members` = the local handle, nmembers = 1; mailboxes
 * shared through 64-byte CUDA IPC handles that the caller moves between the processes). */
int fqsb_slab_init(fqsb_system* s, int rank, int world, int64_t halo_cells, int kmax);
int fqsb_slab_ipc_handle(fqsb_system* s, void* out64);
/* locals [world] (entries may be NULL) : members living in this process; ipc_handles [world][64]
 * (may be NULL): fqsb_slab_ipc_handle of the others */
int fqsb_slab_connect(fqsb_system* s, fqsb_system* const* locals, const void* ipc_handles);
/* out [10]: rank, world, halo_cells, own_lo, own_hi, batches, batches redone (criterion fired
 * inside), CUDA graphs in use, speculative batches discarded, 1 if the members run the blocked kernel */
int fqsb_slab_info(fqsb_system* s, int64_t* out);
int fqsb_slab_exchange(fqsb_system** members, int nmembers);
/* timeSteps (flow == 0) / flowSteps (ref: detail.h:1577-1583, 1637-1645), `batch` steps per exchange */
int fqsb_slab_time_steps(fqsb_system** members, int nmembers, int64_t n, int64_t batch, int flow,
                         double v_frame);
/* minimise, dynamic or overdamped (ref: detail.h:1676-1792); any niter_tol >= 1.
 * *ret: 0 converged, steps + 1 otherwise; *steps: steps taken */
int fqsb_slab_minimise(fqsb_system** members, int nmembers, double tol, int64_t niter_tol,
                       int64_t max_iter, int64_t batch, int max_iter_is_error, int64_t* ret,
                       int64_t* steps);
/* sums over the whole system (members added in rank order), out [4]:
 * what 1 {sum f^2, sum f_frame^2}, 2 {sum v^2, sum f_frame}, 3 {-, off-branch count, -, min
 * displacement}, 4 {sum (i - i_mark), #(i != i_mark), sum |i - i_mark|} */
int fqsb_slab_sums(fqsb_system** members, int nmembers, int what, int direction, double* out);
int fqsb_slab_mark_indices(fqsb_system** members, int nmembers);
/* ref: detail.h:1933-1960 */
int fqsb_slab_event_driven_step(fqsb_system** members, int nmembers, double eps, int kick,
                                int direction, double* du_frame);
/* host-only: replay of the StopList criterion (GooseFEM::Iterate::StopList + detail.h:1780) over
 * a batch log [k][5]; ring_num / ring_den [niter_tol] carry the ring between calls (start: +inf,
 * 1). Returns the 1-based stopping step, 0 if none, -1 on NaN. */
int64_t fqsb_slab_first_stop(const double* log, int64_t k, double tol, int64_t niter_tol,
                             double* ring_num, double* ring_den);

/* host-only: geometry the temporally blocked kernel (1-D nearest-neighbour lines beyond one CTA,
 * BASELINE config #3) would use for a line of n_blocks x n_realisations; no reference counterpart
 * (the reference steps one block at a time, ref: detail.h:1539-1569). steps_per_launch / own_hint
 * 0 = planner's choice. halo_cells > 0: the line is a slab member with that many halo blocks per
 * side, and out[6], out[7] count the tiles that read / write the neighbours' mailboxes when the
 * exchange rides on the tile kernel.
 * out [8]: blocks per thread, owned blocks per tile, tile halo, steps per launch, tiles,
 * tiles per SM (ceil), reader tiles, pusher tiles */
int fqsb_plan_blocked(int64_t n_blocks, int64_t n_realisations, int has_interactions,
                      int steps_per_launch, int own_hint, int64_t halo_cells, int64_t* out);

/* host staging helpers (pinned memory for the e2e path) ---------------------------------- */
void* fqsb_host_alloc(size_t bytes);
void fqsb_host_free(void* p);

/* instrumentation: kernels launched / steps executed by this handle since creation */
int64_t fqsb_launch_count(const fqsb_system* s);
int64_t fqsb_step_count(const fqsb_system* s);
/* name of the stepping kernel the last dynamics call used ("resident", "stream", ...) */
const char* fqsb_last_kernel(const fqsb_system* s);
/* device time (CUDA events on the handle's stream) spent in the stepping-kernel launches of the
 * last dynamics call, and how many launches that was */
double fqsb_last_kernel_seconds(const fqsb_system* s);
int64_t fqsb_last_kernel_launches(const fqsb_system* s);

#ifdef __cplusplus
}
#endif
#endif /* FQSB_H */
