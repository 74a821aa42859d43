/*
 * fqsb.hpp -- C++ host classes with the reference's names and signatures over the C ABI.
 *
 * Mirrors namespace FrictionQPotSpringBlock::{Line1d,Line2d} of the reference
 * (include/FrictionQPotSpringBlock/Line1d.h:112-677, Line2d.h:77-162) and the public members of
 * detail::System (detail.h:1141-1995). Header-only, depends on nothing but <fqsb.h>; link with
 * libfqsb.so. Every failure is a std::runtime_error carrying the reference's text
 * (config.h:19-24), as in the reference.
 *
 * Array getters return a reference to a host mirror that is refreshed by the call (the
 * reference returns a reference to its own storage, detail.h:1402-1468); setters copy, as
 * `xt::noalias(m_u) = arg` does (detail.h:1279).
 */
#ifndef FQSB_HPP
#define FQSB_HPP

#include <array>
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "fqsb.h"

namespace FrictionQPotSpringBlock {

inline std::string version() { return fqsb_version(); } // config.h:209-212

namespace detail {

// detail.h:31-66
inline int string_to_distribution(const std::string& str)
{
    static const char* names[] = {"random", "delta", "exponential", "power",
                                  "gamma",  "pareto", "weibull",    "normal"};
    for (int k = 0; k < 8; ++k) {
        if (str == names[k]) {
            return k;
        }
    }
    throw std::runtime_error("Unknown distribution: " + str);
}

inline void check(int rc)
{
    if (rc != FQSB_OK) {
        throw std::runtime_error(fqsb_last_error());
    }
}

/** detail::RandomNormalForcing<1> (detail.h:881-1000) as seen through `system.external()`:
 *  a view on the forcing state that lives on the device next to the system. */
class RandomNormalForcing {
    fqsb_system* m_h = nullptr;
    mutable std::vector<double> m_f_thermal;
    mutable std::vector<int64_t> m_next;

public:
    RandomNormalForcing() = default;
    explicit RandomNormalForcing(fqsb_system* h) : m_h(h) {}

    uint64_t state() const // detail.h:949-952
    {
        uint64_t x;
        check(fqsb_external_get_state(m_h, &x));
        return x;
    }
    void set_state(uint64_t state) { check(fqsb_external_set_state(m_h, &state)); } // 958-961
    const std::vector<double>& f_thermal() const // detail.h:967-970
    {
        m_f_thermal.resize(static_cast<size_t>(fqsb_size(m_h)));
        check(fqsb_external_get_f_thermal(m_h, m_f_thermal.data(),
                                          static_cast<int64_t>(m_f_thermal.size())));
        return m_f_thermal;
    }
    void set_f_thermal(const std::vector<double>& f_thermal) // detail.h:976-980
    {
        check(fqsb_external_set_f_thermal(m_h, f_thermal.data(),
                                          static_cast<int64_t>(f_thermal.size())));
    }
    const std::vector<int64_t>& next() const // detail.h:986-989
    {
        m_next.resize(static_cast<size_t>(fqsb_size(m_h)));
        check(fqsb_external_get_next(m_h, m_next.data(), static_cast<int64_t>(m_next.size())));
        return m_next;
    }
    void set_next(const std::vector<int64_t>& next) // detail.h:995-999
    {
        check(fqsb_external_set_next(m_h, next.data(), static_cast<int64_t>(next.size())));
    }
};

/** What the reference's constructors have no argument for (new surface): how many independent
 *  realisations one handle integrates, on which device, with which kernel, and -- for the members
 *  of a slab-decomposed system -- which global blocks the local ones are. Consumed by the NEXT
 *  system constructed on this thread (see Ensemble<S> and Slab<S>); default = one reference
 *  system on the current device. */
struct Options {
    int64_t nrealisations = 1;
    int64_t seed_stride = 0;
    int device = -1;
    int kernel = 0;
    /** opt in to the resident kernels built with FMA contraction (FQSB_KERNEL_FMA): same yield
     *  landscape, trajectories equal to rounding instead of bit for bit; default off */
    bool contracted = false;
    int64_t seed_first = 0;
    int64_t seed_period = 0;
};

inline Options& next_options()
{
    thread_local Options o;
    return o;
}

struct OptionsSetter {
    explicit OptionsSetter(const Options& o) { next_options() = o; }
};

/** `system.chunk()`: the prrng::pcg32_tensor_cumsum object of the reference (detail.h:1146-1149,
 *  python/main.cpp:65-70) as far as the library and its tests use it. The device keeps no chunk --
 *  only the current well and the generator state of every block -- so `data()` / `start()` are a
 *  window of `chunk_size()` yield positions regenerated on demand, positioned like prrng's
 *  alignment(buffer = 2, margin = 30) would (Line1d.h:154). */
class Chunk {
    fqsb_system* m_h = nullptr;
    int64_t m_nchunk = 0;
    mutable std::vector<int64_t> m_start, m_i, m_tmp;
    mutable std::vector<double> m_y, m_data;
    mutable std::vector<uint64_t> m_state;

    size_t n() const { return static_cast<size_t>(fqsb_size(m_h) * fqsb_nrealisations(m_h)); }

    void update_start() const
    {
        const size_t N = n();
        m_start.resize(N, 0);
        m_i.resize(N);
        check(fqsb_chunk_index_at_align(m_h, m_i.data(), static_cast<int64_t>(N)));
        for (size_t p = 0; p < N; ++p) {
            const int64_t loc = m_i[p] - m_start[p];
            if (loc < 2 || loc >= m_nchunk - 1 - 2) {
                m_start[p] = m_i[p] > 30 ? m_i[p] - 30 : 0;
            }
        }
    }

public:
    Chunk() = default;
    Chunk(fqsb_system* h, int64_t nchunk) : m_h(h), m_nchunk(nchunk) {}

    size_t chunk_size() const { return static_cast<size_t>(m_nchunk); }

    const std::vector<int64_t>& index_at_align() const
    {
        m_i.resize(n());
        check(fqsb_chunk_index_at_align(m_h, m_i.data(), static_cast<int64_t>(m_i.size())));
        return m_i;
    }
    const std::vector<double>& left_of_align() const
    {
        m_y.resize(n());
        check(fqsb_chunk_left_of_align(m_h, m_y.data(), static_cast<int64_t>(m_y.size())));
        return m_y;
    }
    const std::vector<double>& right_of_align() const
    {
        m_y.resize(n());
        check(fqsb_chunk_right_of_align(m_h, m_y.data(), static_cast<int64_t>(m_y.size())));
        return m_y;
    }
    const std::vector<int64_t>& start() const
    {
        update_start();
        return m_start;
    }
    const std::vector<int64_t>& chunk_index_at_align() const
    {
        update_start();
        m_tmp.resize(m_i.size());
        for (size_t p = 0; p < m_i.size(); ++p) {
            m_tmp[p] = m_i[p] - m_start[p];
        }
        return m_tmp;
    }
    /** yield positions [size][chunk_size], row-major */
    const std::vector<double>& data() const
    {
        update_start();
        m_data.resize(m_start.size() * static_cast<size_t>(m_nchunk));
        check(fqsb_chunk_data(m_h, m_start.data(), m_nchunk, m_data.data()));
        return m_data;
    }
    void align(const std::vector<double>& u)
    {
        check(fqsb_chunk_align(m_h, u.data(), static_cast<int64_t>(u.size())));
    }
    const std::vector<uint64_t>& state_at(const std::vector<int64_t>& index) const
    {
        m_state.resize(index.size());
        check(fqsb_chunk_state_at(m_h, index.data(), m_state.data(),
                                  static_cast<int64_t>(index.size())));
        return m_state;
    }
    void restore(const std::vector<uint64_t>& state, const std::vector<double>& value,
                 const std::vector<int64_t>& index)
    {
        if (state.size() != value.size() || state.size() != index.size()) {
            throw std::runtime_error("assertion failed (state, value, index of equal shape)");
        }
        check(fqsb_chunk_restore(m_h, state.data(), value.data(), index.data(),
                                 static_cast<int64_t>(state.size())));
        m_start = index;
    }
};

/** detail::System (detail.h:1046-2051), all work on the GPU. One realisation unless constructed
 *  through Ensemble<S>: arrays are then [nrealisations][size] and the scalar accessors return the
 *  first realisation (the `*_all` ones return every realisation's). */
class System {
protected:
    fqsb_system* m_h = nullptr;
    fqsb_params m_par{};
    std::vector<size_t> m_shape;
    mutable std::vector<double> m_mirror[8];
    mutable std::vector<int64_t> m_index;
    mutable std::vector<double> m_scalar;
    mutable std::vector<int64_t> m_iscalar;
    Chunk m_chunk;

    size_t R() const { return static_cast<size_t>(m_par.nrealisations); }
    double* dbuf() const
    {
        m_scalar.resize(R());
        return m_scalar.data();
    }
    int64_t* ibuf() const
    {
        m_iscalar.resize(R());
        return m_iscalar.data();
    }
    const double* spread(double x) const
    {
        m_scalar.assign(R(), x);
        return m_scalar.data();
    }

    void initSystem(int potential, int interactions, int minimisation,
                    const std::vector<size_t>& shape, double m, double eta, double mu,
                    double kappa, double k1, double k2, double k_frame, double dt, uint64_t seed,
                    const std::string& distribution, const std::vector<double>& parameters,
                    double offset, size_t nchunk)
    {
        m_shape = shape;
        m_par.potential = potential;
        m_par.interactions = interactions;
        m_par.minimisation = minimisation;
        m_par.rank = static_cast<int32_t>(shape.size());
        m_par.shape[0] = static_cast<int64_t>(shape[0]);
        m_par.shape[1] = shape.size() > 1 ? static_cast<int64_t>(shape[1]) : 1;
        m_par.m = m;
        m_par.eta = eta;
        m_par.mu = mu;
        m_par.kappa = kappa;
        m_par.k1 = k1;
        m_par.k2 = k2;
        m_par.k_frame = k_frame;
        m_par.dt = dt;
        m_par.seed = seed;
        m_par.distribution = string_to_distribution(distribution);
        if (parameters.size() > 4) {
            throw std::runtime_error("at most 4 distribution parameters");
        }
        m_par.nparameters = static_cast<int32_t>(parameters.size());
        for (int k = 0; k < m_par.nparameters; ++k) {
            m_par.parameters[k] = parameters[static_cast<size_t>(k)];
        }
        m_par.offset = offset;
        m_par.nchunk = static_cast<int64_t>(nchunk);
        const Options opt = next_options();
        next_options() = Options{}; // consumed
        m_par.nrealisations = opt.nrealisations > 0 ? opt.nrealisations : 1;
        m_par.seed_stride = opt.seed_stride;
        m_par.device = opt.device;
        m_par.kernel = opt.kernel | (opt.contracted ? FQSB_KERNEL_FMA : 0);
        m_par.seed_first = opt.seed_first;
        m_par.seed_period = opt.seed_period;
        check(fqsb_create(&m_par, &m_h));
        m_chunk = Chunk(m_h, m_par.nchunk);
    }

    // External = RandomNormalForcing (Line1d.h:316-318), before initSystem's refresh()
    void initForcing(double mean, double stddev, uint64_t seed_forcing,
                     const std::vector<int64_t>& dinc_init, const std::vector<int64_t>& dinc)
    {
        if (dinc_init.size() != dinc.size()) {
            throw std::runtime_error("assertion failed (xt::has_shape(dinc_init, dinc.shape()))");
        }
        check(fqsb_enable_random_forcing(m_h, mean, stddev, seed_forcing, 1, dinc_init.data(),
                                         dinc.data(), static_cast<int64_t>(dinc.size())));
        m_external = RandomNormalForcing(m_h);
    }

    RandomNormalForcing m_external;

    const std::vector<double>& array(int which) const
    {
        auto& buf = m_mirror[which];
        buf.resize(this->size() * R());
        check(fqsb_get(m_h, which, buf.data(), static_cast<int64_t>(buf.size())));
        return buf;
    }

    System() = default;

public:
    System(const System&) = delete;
    System& operator=(const System&) = delete;
    virtual ~System() { fqsb_destroy(m_h); }

    fqsb_system* handle() const { return m_h; }
    size_t size() const { return static_cast<size_t>(fqsb_size(m_h)); } // detail.h:1155
    const std::vector<size_t>& shape() const { return m_shape; }        // detail.h:1164
    double dt() const { return m_par.dt; }
    double mu() const { return m_par.mu; }
    double eta() const { return m_par.eta; }
    double m() const { return m_par.m; }
    double k_frame() const { return m_par.k_frame; }

    size_t nrealisations() const { return R(); }
    /** Generator of the yield landscape (detail.h:1146-1149) */
    Chunk& chunk() { return m_chunk; }
    const Chunk& chunk() const { return m_chunk; }

    // detail.h:1231-1315
    void set_t(double arg) { check(fqsb_set_t(m_h, spread(arg))); }
    void set_inc(int64_t arg)
    {
        m_iscalar.assign(R(), arg);
        check(fqsb_set_inc(m_h, m_iscalar.data()));
    }
    void set_u_frame(double arg) { check(fqsb_set_u_frame(m_h, spread(arg))); }
    void set_u_frame_all(const std::vector<double>& arg)
    {
        if (arg.size() != R()) {
            throw std::runtime_error("assertion failed (one frame position per realisation)");
        }
        check(fqsb_set_u_frame(m_h, arg.data()));
    }
    void set_u(const std::vector<double>& arg)
    {
        check(fqsb_set_u(m_h, arg.data(), static_cast<int64_t>(arg.size())));
    }
    void set_v(const std::vector<double>& arg)
    {
        check(fqsb_set_v(m_h, arg.data(), static_cast<int64_t>(arg.size())));
    }
    void set_a(const std::vector<double>& arg)
    {
        check(fqsb_set_a(m_h, arg.data(), static_cast<int64_t>(arg.size())));
    }
    void refresh() { check(fqsb_refresh(m_h)); }
    void quench() { check(fqsb_quench(m_h)); }

    // detail.h:1402-1520
    const std::vector<double>& u() const { return array(FQSB_U); }
    const std::vector<double>& v() const { return array(FQSB_V); }
    const std::vector<double>& a() const { return array(FQSB_A); }
    const std::vector<double>& f() const { return array(FQSB_F); }
    const std::vector<double>& f_potential() const { return array(FQSB_F_POTENTIAL); }
    const std::vector<double>& f_frame() const { return array(FQSB_F_FRAME); }
    const std::vector<double>& f_interactions() const { return array(FQSB_F_INTERACTIONS); }
    const std::vector<double>& f_damping() const { return array(FQSB_F_DAMPING); }
    const std::vector<double>& u_frame_all() const
    {
        check(fqsb_get_u_frame(m_h, dbuf()));
        return m_scalar;
    }
    double u_frame() const { return u_frame_all()[0]; }
    double t() const
    {
        check(fqsb_get_t(m_h, dbuf()));
        return m_scalar[0];
    }
    const std::vector<int64_t>& inc_all() const
    {
        check(fqsb_get_inc(m_h, ibuf()));
        return m_iscalar;
    }
    int64_t inc() const { return inc_all()[0]; }
    double temperature() const
    {
        check(fqsb_temperature(m_h, dbuf()));
        return m_scalar[0];
    }
    const std::vector<double>& residual_all() const
    {
        check(fqsb_residual(m_h, dbuf()));
        return m_scalar;
    }
    double residual() const { return residual_all()[0]; }
    /** np.mean(system.f_frame) per realisation, reduced on the device */
    const std::vector<double>& mean_f_frame_all() const
    {
        check(fqsb_mean_f_frame(m_h, dbuf()));
        return m_scalar;
    }
    size_t quasistaticActivityFirst() const
    {
        check(fqsb_qs_activity(m_h, ibuf(), nullptr));
        return static_cast<size_t>(m_iscalar[0]);
    }
    size_t quasistaticActivityLast() const
    {
        check(fqsb_qs_activity(m_h, nullptr, ibuf()));
        return static_cast<size_t>(m_iscalar[0]);
    }

    // the chunk surface used by the library itself (detail.h:1602,1609): global well index
    const std::vector<int64_t>& index_at_align() const
    {
        m_index.resize(this->size() * R());
        check(fqsb_chunk_index_at_align(m_h, m_index.data(), static_cast<int64_t>(m_index.size())));
        return m_index;
    }

    // device-side avalanche bookkeeping: i_n kept on the device
    void mark_indices() { check(fqsb_mark_indices(m_h)); }
    /** S = sum(i - i_n) and A = #(i != i_n) per realisation since mark_indices() */
    void avalanche_since_mark(std::vector<int64_t>& S, std::vector<int64_t>& A) const
    {
        S.resize(R());
        A.resize(R());
        check(fqsb_avalanche_since_mark(m_h, S.data(), A.data()));
    }

    /** set_u / set_v / set_a, timeSteps(n) and the read-back of u (+ mean f_frame per realisation)
     *  as ONE call whose copies overlap the kernels (fqsb_run_from_host); nullptr = keep / skip */
    void run_from_host(size_t n, const double* u, const double* v, const double* a, double* out_u,
                       double* out_mean_f_frame = nullptr)
    {
        check(fqsb_run_from_host(m_h, u, v, a, static_cast<int64_t>(this->size() * R()),
                                 static_cast<int64_t>(n), out_u, nullptr, nullptr,
                                 out_mean_f_frame));
    }

    // detail.h:1539-1645
    void timeStep() { check(fqsb_time_steps(m_h, 1)); }
    void timeSteps(size_t n) { check(fqsb_time_steps(m_h, static_cast<int64_t>(n))); }
    void flowSteps(size_t n, double v_frame)
    {
        check(fqsb_flow_steps(m_h, static_cast<int64_t>(n), v_frame));
    }
    size_t timeStepsUntilEvent(double tol = 1e-5, size_t niter_tol = 10, size_t max_iter = 1e9)
    {
        check(fqsb_time_steps_until_event(m_h, tol, static_cast<int64_t>(niter_tol),
                                          static_cast<int64_t>(max_iter), ibuf()));
        return static_cast<size_t>(m_iscalar[0]);
    }

    // detail.h:1676-1893
    const std::vector<int64_t>& minimise_all(double tol = 1e-5, size_t niter_tol = 10,
                                             size_t max_iter = 1e9, bool time_activity = false,
                                             bool max_iter_is_error = true)
    {
        check(fqsb_minimise(m_h, tol, static_cast<int64_t>(niter_tol),
                            static_cast<int64_t>(max_iter), time_activity, max_iter_is_error,
                            ibuf()));
        return m_iscalar;
    }
    size_t minimise(double tol = 1e-5, size_t niter_tol = 10, size_t max_iter = 1e9,
                    bool time_activity = false, bool max_iter_is_error = true)
    {
        return static_cast<size_t>(
            minimise_all(tol, niter_tol, max_iter, time_activity, max_iter_is_error)[0]);
    }
    size_t minimise_truncate(const std::vector<int64_t>& i_n, size_t A_truncate = 0,
                             size_t S_truncate = 0, double tol = 1e-5, size_t niter_tol = 10,
                             size_t max_iter = 1e9, bool time_activity = true,
                             bool max_iter_is_error = true)
    {
        if (i_n.size() != this->size() * R()) {
            throw std::runtime_error("assertion failed (xt::has_shape(i_n, m_u.shape()))");
        }
        check(fqsb_minimise_truncate(m_h, i_n.data(), static_cast<int64_t>(A_truncate),
                                     static_cast<int64_t>(S_truncate), tol,
                                     static_cast<int64_t>(niter_tol),
                                     static_cast<int64_t>(max_iter), time_activity,
                                     max_iter_is_error, ibuf()));
        return static_cast<size_t>(m_iscalar[0]);
    }

    // detail.h:1901-1995
    double maxUniformDisplacement(int direction = 1)
    {
        check(fqsb_max_uniform_displacement(m_h, direction, dbuf()));
        return m_scalar[0];
    }
    const std::vector<double>& eventDrivenStep_all(double eps, bool kick, int direction = 1)
    {
        check(fqsb_event_driven_step(m_h, eps, kick, direction, dbuf()));
        return m_scalar;
    }
    double eventDrivenStep(double eps, bool kick, int direction = 1)
    {
        return eventDrivenStep_all(eps, kick, direction)[0];
    }
    void trigger(size_t p, double eps, int direction = 1)
    {
        check(fqsb_trigger(m_h, 0, static_cast<int64_t>(p), eps, direction));
    }
    void advanceToFixedForce(double f_frame, bool allow_plastic = false)
    {
        check(fqsb_advance_to_fixed_force(m_h, spread(f_frame), allow_plastic));
    }
};

} // namespace detail

namespace Line1d {

#define FQSB_SHAPE1 std::vector<size_t>{shape[0]}

/** Line1d.h:112-162 */
class System_Cuspy_Laplace : public detail::System {
public:
    System_Cuspy_Laplace(double m, double eta, double mu, double k_interactions, double k_frame,
                         double dt, const std::array<size_t, 1>& shape, uint64_t seed,
                         const std::string& distribution, const std::vector<double>& parameters,
                         double offset = -100.0, size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_CUSPY, FQSB_INT_LAPLACE1D, FQSB_MIN_DYNAMIC, FQSB_SHAPE1, m, eta, mu,
                   0.0, k_interactions, 0.0, k_frame, dt, seed, distribution, parameters, offset,
                   nchunk);
    }
};

/** Line1d.h:173-238: minimisation only; the dynamics are hidden as in the reference */
class System_Cuspy_Laplace_Nopassing : public detail::System {
public:
    System_Cuspy_Laplace_Nopassing(double mu, double k_interactions, double k_frame,
                                   const std::array<size_t, 1>& shape, uint64_t seed,
                                   const std::string& distribution,
                                   const std::vector<double>& parameters, double offset = -100.0,
                                   size_t nchunk = 5000, double eta = 0.0, double dt = 0.0)
    {
        initSystem(FQSB_POT_CUSPY, FQSB_INT_LAPLACE1D, FQSB_MIN_OVERDAMPED, FQSB_SHAPE1, 1.0, eta,
                   mu, 0.0, k_interactions, 0.0, k_frame, dt, seed, distribution, parameters,
                   offset, nchunk);
    }

protected:
    using detail::System::flowSteps;
    using detail::System::timeStep;
    using detail::System::timeSteps;
    using detail::System::timeStepsUntilEvent;
};

/** Line1d.h:336-377 */
class System_SemiSmooth_Laplace : public detail::System {
public:
    System_SemiSmooth_Laplace(double m, double eta, double mu, double kappa, double k_interactions,
                              double k_frame, double dt, const std::array<size_t, 1>& shape,
                              uint64_t seed, const std::string& distribution,
                              const std::vector<double>& parameters, double offset = -100.0,
                              size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_SEMISMOOTH, FQSB_INT_LAPLACE1D, FQSB_MIN_DYNAMIC, FQSB_SHAPE1, m, eta,
                   mu, kappa, k_interactions, 0.0, k_frame, dt, seed, distribution, parameters,
                   offset, nchunk);
    }
};

/** Line1d.h:383-422 */
class System_Smooth_Laplace : public detail::System {
public:
    System_Smooth_Laplace(double m, double eta, double mu, double k_interactions, double k_frame,
                          double dt, const std::array<size_t, 1>& shape, uint64_t seed,
                          const std::string& distribution, const std::vector<double>& parameters,
                          double offset = -100.0, size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_SMOOTH, FQSB_INT_LAPLACE1D, FQSB_MIN_DYNAMIC, FQSB_SHAPE1, m, eta, mu,
                   0.0, k_interactions, 0.0, k_frame, dt, seed, distribution, parameters, offset,
                   nchunk);
    }
};

/** Line1d.h:428-480 */
class System_Cuspy_Quartic : public detail::System {
public:
    System_Cuspy_Quartic(double m, double eta, double mu, double a1, double a2, double k_frame,
                         double dt, const std::array<size_t, 1>& shape, uint64_t seed,
                         const std::string& distribution, const std::vector<double>& parameters,
                         double offset = -100.0, size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_CUSPY, FQSB_INT_QUARTIC1D, FQSB_MIN_DYNAMIC, FQSB_SHAPE1, m, eta, mu,
                   0.0, a1, a2, k_frame, dt, seed, distribution, parameters, offset, nchunk);
    }
};

/** Thermal systems hide the athermal protocol (Line1d.h:321-329) */
#define FQSB_HIDE_ATHERMAL \
protected: \
    using detail::System::eventDrivenStep; \
    using detail::System::quasistaticActivityFirst; \
    using detail::System::quasistaticActivityLast;

/** Line1d.h:261-330 */
class System_Cuspy_Laplace_RandomForcing : public detail::System {
public:
    System_Cuspy_Laplace_RandomForcing(double m, double eta, double mu, double k_interactions,
                                       double k_frame, double dt, double mean, double stddev,
                                       uint64_t seed_forcing,
                                       const std::vector<int64_t>& dinc_init,
                                       const std::vector<int64_t>& dinc,
                                       const std::array<size_t, 1>& shape, uint64_t seed,
                                       const std::string& distribution,
                                       const std::vector<double>& parameters,
                                       double offset = -100.0, size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_CUSPY, FQSB_INT_LAPLACE1D, FQSB_MIN_NONE, FQSB_SHAPE1, m, eta, mu, 0.0,
                   k_interactions, 0.0, k_frame, dt, seed, distribution, parameters, offset,
                   nchunk);
        initForcing(mean, stddev, seed_forcing, dinc_init, dinc);
    }
    const detail::RandomNormalForcing& external() const { return m_external; } // detail.h:1218
    detail::RandomNormalForcing& external() { return m_external; }
    FQSB_HIDE_ATHERMAL
};

/** Line1d.h:486-556 */
class System_Cuspy_Quartic_RandomForcing : public detail::System {
public:
    System_Cuspy_Quartic_RandomForcing(double m, double eta, double mu, double a1, double a2,
                                       double k_frame, double dt, double mean, double stddev,
                                       uint64_t seed_forcing,
                                       const std::vector<int64_t>& dinc_init,
                                       const std::vector<int64_t>& dinc,
                                       const std::array<size_t, 1>& shape, uint64_t seed,
                                       const std::string& distribution,
                                       const std::vector<double>& parameters,
                                       double offset = -100.0, size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_CUSPY, FQSB_INT_QUARTIC1D, FQSB_MIN_NONE, FQSB_SHAPE1, m, eta, mu, 0.0,
                   a1, a2, k_frame, dt, seed, distribution, parameters, offset, nchunk);
        initForcing(mean, stddev, seed_forcing, dinc_init, dinc);
    }
    const detail::RandomNormalForcing& external() const { return m_external; }
    detail::RandomNormalForcing& external() { return m_external; }
    FQSB_HIDE_ATHERMAL
};

#undef FQSB_HIDE_ATHERMAL

/** Line1d.h:562-614 */
class System_Cuspy_QuarticGradient : public detail::System {
public:
    System_Cuspy_QuarticGradient(double m, double eta, double mu, double k2, double k4,
                                 double k_frame, double dt, const std::array<size_t, 1>& shape,
                                 uint64_t seed, const std::string& distribution,
                                 const std::vector<double>& parameters, double offset = -100.0,
                                 size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_CUSPY, FQSB_INT_QUARTICGRADIENT1D, FQSB_MIN_DYNAMIC, FQSB_SHAPE1, m,
                   eta, mu, 0.0, k2, k4, k_frame, dt, seed, distribution, parameters, offset,
                   nchunk);
    }
};

/** Line1d.h:620-672 */
class System_Cuspy_LongRange : public detail::System {
public:
    System_Cuspy_LongRange(double m, double eta, double mu, double k_interactions, double alpha,
                           double k_frame, double dt, const std::array<size_t, 1>& shape,
                           uint64_t seed, const std::string& distribution,
                           const std::vector<double>& parameters, double offset = -100.0,
                           size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_CUSPY, FQSB_INT_LONGRANGE1D, FQSB_MIN_DYNAMIC, FQSB_SHAPE1, m, eta, mu,
                   0.0, k_interactions, alpha, k_frame, dt, seed, distribution, parameters, offset,
                   nchunk);
    }
};

#undef FQSB_SHAPE1

} // namespace Line1d

/** Particles.h:93-311: independent particles (no interactions) */
namespace Particles {

#define FQSB_SHAPE1 std::vector<size_t>{shape[0]}

/** Particles.h:93-135 */
class System_Cuspy : public detail::System {
public:
    System_Cuspy(double m, double eta, double mu, double k_frame, double dt,
                 const std::array<size_t, 1>& shape, uint64_t seed,
                 const std::string& distribution, const std::vector<double>& parameters,
                 double offset = -100.0, size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_CUSPY, FQSB_INT_NONE, FQSB_MIN_DYNAMIC, FQSB_SHAPE1, m, eta, mu, 0.0,
                   0.0, 0.0, k_frame, dt, seed, distribution, parameters, offset, nchunk);
    }
};

/** Particles.h:161-230 */
class System_Cuspy_RandomForcing : public detail::System {
public:
    System_Cuspy_RandomForcing(double m, double eta, double mu, double k_frame, double dt,
                               double mean, double stddev, uint64_t seed_forcing,
                               const std::vector<int64_t>& dinc_init,
                               const std::vector<int64_t>& dinc,
                               const std::array<size_t, 1>& shape, uint64_t seed,
                               const std::string& distribution,
                               const std::vector<double>& parameters, double offset = -100.0,
                               size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_CUSPY, FQSB_INT_NONE, FQSB_MIN_NONE, FQSB_SHAPE1, m, eta, mu, 0.0, 0.0,
                   0.0, k_frame, dt, seed, distribution, parameters, offset, nchunk);
        initForcing(mean, stddev, seed_forcing, dinc_init, dinc);
    }
    const detail::RandomNormalForcing& external() const { return m_external; }
    detail::RandomNormalForcing& external() { return m_external; }

protected:
    using detail::System::eventDrivenStep;
    using detail::System::quasistaticActivityFirst;
    using detail::System::quasistaticActivityLast;
};

/** Particles.h:233-270 (runs the SemiSmooth_Laplace kernels with k_interactions = 0: the
 *  interaction term is an exact zero) */
class System_SemiSmooth : public detail::System {
public:
    System_SemiSmooth(double m, double eta, double mu, double kappa, double k_frame, double dt,
                      const std::array<size_t, 1>& shape, uint64_t seed,
                      const std::string& distribution, const std::vector<double>& parameters,
                      double offset = -100.0, size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_SEMISMOOTH, FQSB_INT_LAPLACE1D, FQSB_MIN_DYNAMIC, FQSB_SHAPE1, m, eta,
                   mu, kappa, 0.0, 0.0, k_frame, dt, seed, distribution, parameters, offset,
                   nchunk);
    }
};

/** Particles.h:276-311 (Smooth_Laplace kernels with k_interactions = 0) */
class System_Smooth : public detail::System {
public:
    System_Smooth(double m, double eta, double mu, double k_frame, double dt,
                  const std::array<size_t, 1>& shape, uint64_t seed,
                  const std::string& distribution, const std::vector<double>& parameters,
                  double offset = -100.0, size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_SMOOTH, FQSB_INT_LAPLACE1D, FQSB_MIN_DYNAMIC, FQSB_SHAPE1, m, eta, mu,
                   0.0, 0.0, 0.0, k_frame, dt, seed, distribution, parameters, offset, nchunk);
    }
};

#undef FQSB_SHAPE1

} // namespace Particles

namespace Line2d {

#define FQSB_SHAPE2 std::vector<size_t>{shape[0], shape[1]}

/** Line2d.h:77-117 */
class System_Cuspy_Laplace : public detail::System {
public:
    System_Cuspy_Laplace(double m, double eta, double mu, double k_interactions, double k_frame,
                         double dt, const std::array<size_t, 2>& shape, uint64_t seed,
                         const std::string& distribution, const std::vector<double>& parameters,
                         double offset = -100.0, size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_CUSPY, FQSB_INT_LAPLACE2D, FQSB_MIN_DYNAMIC, FQSB_SHAPE2, m, eta, mu,
                   0.0, k_interactions, 0.0, k_frame, dt, seed, distribution, parameters, offset,
                   nchunk);
    }
};

/** Line2d.h:123-162 */
class System_Cuspy_QuarticGradient : public detail::System {
public:
    System_Cuspy_QuarticGradient(double m, double eta, double mu, double k2, double k4,
                                 double k_frame, double dt, const std::array<size_t, 2>& shape,
                                 uint64_t seed, const std::string& distribution,
                                 const std::vector<double>& parameters, double offset = -100.0,
                                 size_t nchunk = 5000)
    {
        initSystem(FQSB_POT_CUSPY, FQSB_INT_QUARTICGRADIENT2D, FQSB_MIN_DYNAMIC, FQSB_SHAPE2, m,
                   eta, mu, 0.0, k2, k4, k_frame, dt, seed, distribution, parameters, offset,
                   nchunk);
    }
};

/** 2-D generalisation of Line1d::System_Cuspy_Laplace_Nopassing (overdamped no-passing sweeps on
 *  the 5-point lattice; new: the reference has the 1-D class only, Line1d.h:173-238) */
class System_Cuspy_Laplace_Nopassing : public detail::System {
public:
    System_Cuspy_Laplace_Nopassing(double mu, double k_interactions, double k_frame,
                                   const std::array<size_t, 2>& shape, uint64_t seed,
                                   const std::string& distribution,
                                   const std::vector<double>& parameters, double offset = -100.0,
                                   size_t nchunk = 5000, double eta = 0.0, double dt = 0.0)
    {
        initSystem(FQSB_POT_CUSPY, FQSB_INT_LAPLACE2D, FQSB_MIN_OVERDAMPED, FQSB_SHAPE2, 1.0, eta,
                   mu, 0.0, k_interactions, 0.0, k_frame, dt, seed, distribution, parameters,
                   offset, nchunk);
    }

private:
    using detail::System::flowSteps;
    using detail::System::timeStep;
    using detail::System::timeSteps;
    using detail::System::timeStepsUntilEvent;
};

#undef FQSB_SHAPE2

} // namespace Line2d

/** `nrealisations` independent systems of class S in one device-resident handle (new surface: the
 *  reference has no batch class). Realisation r is the reference system constructed with
 *  seed + r * seed_stride (seed_stride defaults to the number of blocks). Arrays are
 *  [nrealisations][size]; use the `*_all` accessors for per-realisation scalars.
 *
 *      Ensemble<Line1d::System_Cuspy_Laplace> ens({16384}, m, eta, mu, k, k_frame, dt, shape, ...);
 */
template <class S>
class Ensemble : private detail::OptionsSetter, public S {
public:
    template <class... Args>
    explicit Ensemble(const detail::Options& options, Args&&... args)
        : detail::OptionsSetter(options), S(std::forward<Args>(args)...)
    {
    }
};

/** ONE very large line / interface of class S spread over several GPUs of this process (SURVEY.md
 *  section 8e): member g integrates a contiguous range of rows extended by `halo` rows per side;
 *  halo rows travel as NVLink peer stores between the members' GPUs inside libfqsb.so
 *  (fqsb_slab_*), the per-step stop decision of the reference is replayed per batch.
 *
 *      auto make = [&](const std::array<size_t, 2>& local) {
 *          return std::make_unique<Line2d::System_Cuspy_Laplace>(m, eta, mu, k, k_frame, dt, local,
 *                                                                seed, "random", par, -50.0); };
 *      Slab<Line2d::System_Cuspy_Laplace, 2> slab({0, 1, 2, 3}, 32, {4096, 4096}, make);
 */
template <class S, size_t Rank>
class Slab {
    std::vector<std::unique_ptr<S>> m_members;
    std::vector<fqsb_system*> m_handles;
    std::array<size_t, Rank> m_shape;
    int64_t m_batch = 0;
    size_t m_size = 1;

public:
    template <class Make>
    Slab(const std::vector<int>& devices, size_t halo, const std::array<size_t, Rank>& shape,
         Make&& make, bool overdamped = false, int kernel = -1)
        : m_shape(shape)
    {
        const size_t G = devices.size();
        size_t unit = 1;
        for (size_t d = 1; d < Rank; ++d) {
            unit *= shape[d];
        }
        m_size = unit * shape[0];
        const size_t base = shape[0] / G, extra = shape[0] % G;
        if (G == 0 || base < halo || halo < 1) {
            throw std::runtime_error("every member must own at least `halo` rows");
        }
        if (kernel < 0) { // 1-D dynamic lines: temporally blocked kernel; otherwise streaming
            kernel = (Rank == 1 && !overdamped) ? 0 : 2;
        }
        m_batch = static_cast<int64_t>(overdamped ? halo - 1 : halo);
        if ((kernel & 15) != 2 && m_batch > 64) {
            m_batch = 64;
        }
        for (size_t g = 0; g < G; ++g) {
            const size_t lo = g * base + (g < extra ? g : extra);
            const size_t cnt = base + (g < extra ? 1 : 0);
            detail::Options opt;
            opt.device = devices[g];
            opt.kernel = kernel;
            opt.seed_first = static_cast<int64_t>(((lo + shape[0] - halo) % shape[0]) * unit);
            opt.seed_period = static_cast<int64_t>(m_size);
            std::array<size_t, Rank> local = shape;
            local[0] = cnt + 2 * halo;
            detail::next_options() = opt;
            m_members.push_back(make(local));
            m_handles.push_back(m_members.back()->handle());
            detail::check(fqsb_slab_init(m_handles.back(), static_cast<int>(g), static_cast<int>(G),
                                         static_cast<int64_t>(halo * unit),
                                         static_cast<int>(halo)));
        }
        for (size_t g = 0; g < G; ++g) {
            detail::check(fqsb_slab_connect(m_handles[g], m_handles.data(), nullptr));
        }
        detail::check(fqsb_slab_exchange(m_handles.data(), static_cast<int>(G)));
    }

    size_t size() const { return m_size; }
    size_t members() const { return m_members.size(); }
    S& member(size_t g) { return *m_members[g]; }
    int n() const { return static_cast<int>(m_handles.size()); }

    void set_u_frame(double x)
    {
        for (auto& m : m_members) {
            m->set_u_frame(x);
        }
    }
    double u_frame() const { return m_members[0]->u_frame(); }
    int64_t inc() const { return m_members[0]->inc(); }
    void timeSteps(size_t nsteps)
    {
        detail::check(fqsb_slab_time_steps(m_handles.data(), n(), static_cast<int64_t>(nsteps),
                                           m_batch, 0, 0.0));
    }
    void flowSteps(size_t nsteps, double v_frame)
    {
        detail::check(fqsb_slab_time_steps(m_handles.data(), n(), static_cast<int64_t>(nsteps),
                                           m_batch, 1, v_frame));
    }
    size_t minimise(double tol = 1e-5, size_t niter_tol = 10, size_t max_iter = 1e9,
                    bool max_iter_is_error = true)
    {
        int64_t ret = 0, steps = 0;
        detail::check(fqsb_slab_minimise(m_handles.data(), n(), tol,
                                         static_cast<int64_t>(niter_tol),
                                         static_cast<int64_t>(max_iter), m_batch,
                                         max_iter_is_error, &ret, &steps));
        return static_cast<size_t>(ret);
    }
    double eventDrivenStep(double eps, bool kick, int direction = 1)
    {
        double x = 0.0;
        detail::check(
            fqsb_slab_event_driven_step(m_handles.data(), n(), eps, kick, direction, &x));
        return x;
    }
    double residual()
    {
        double s[4];
        detail::check(fqsb_slab_sums(m_handles.data(), n(), 1, 1, s));
        const double r_fres = std::sqrt(s[0]), r_fext = std::sqrt(s[1]);
        return r_fext != 0.0 ? r_fres / r_fext : r_fres; // detail.h:1512-1520
    }
    void mark_indices() { detail::check(fqsb_slab_mark_indices(m_handles.data(), n())); }
    /** S = sum(i - i_n), A = #(i != i_n) of the whole system since mark_indices() */
    void avalanche_since_mark(int64_t& S_out, int64_t& A_out)
    {
        double s[4];
        detail::check(fqsb_slab_sums(m_handles.data(), n(), 4, 1, s));
        S_out = static_cast<int64_t>(std::llround(s[0]));
        A_out = static_cast<int64_t>(std::llround(s[1]));
    }
};

} // namespace FrictionQPotSpringBlock

#endif /* FQSB_HPP */
