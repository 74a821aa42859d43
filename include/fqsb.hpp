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
#include <cstdint>
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

/** detail::System (detail.h:1046-2051) for one realisation, all work on the GPU. */
class System {
protected:
    fqsb_system* m_h = nullptr;
    fqsb_params m_par{};
    std::vector<size_t> m_shape;
    mutable std::vector<double> m_mirror[8];
    mutable std::vector<int64_t> m_index;

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
        m_par.nparameters = static_cast<int32_t>(parameters.size() < 4 ? parameters.size() : 4);
        for (int k = 0; k < m_par.nparameters; ++k) {
            m_par.parameters[k] = parameters[static_cast<size_t>(k)];
        }
        m_par.offset = offset;
        m_par.nchunk = static_cast<int64_t>(nchunk);
        m_par.nrealisations = 1;
        m_par.seed_stride = 0;
        m_par.device = -1;
        m_par.kernel = 0;
        check(fqsb_create(&m_par, &m_h));
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
        buf.resize(this->size());
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

    // detail.h:1231-1315
    void set_t(double arg) { check(fqsb_set_t(m_h, &arg)); }
    void set_inc(int64_t arg) { check(fqsb_set_inc(m_h, &arg)); }
    void set_u_frame(double arg) { check(fqsb_set_u_frame(m_h, &arg)); }
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
    double u_frame() const
    {
        double x;
        check(fqsb_get_u_frame(m_h, &x));
        return x;
    }
    double t() const
    {
        double x;
        check(fqsb_get_t(m_h, &x));
        return x;
    }
    int64_t inc() const
    {
        int64_t x;
        check(fqsb_get_inc(m_h, &x));
        return x;
    }
    double temperature() const
    {
        double x;
        check(fqsb_temperature(m_h, &x));
        return x;
    }
    double residual() const
    {
        double x;
        check(fqsb_residual(m_h, &x));
        return x;
    }
    size_t quasistaticActivityFirst() const
    {
        int64_t first;
        check(fqsb_qs_activity(m_h, &first, nullptr));
        return static_cast<size_t>(first);
    }
    size_t quasistaticActivityLast() const
    {
        int64_t last;
        check(fqsb_qs_activity(m_h, nullptr, &last));
        return static_cast<size_t>(last);
    }

    // the chunk surface used by the library itself (detail.h:1602,1609): global well index
    const std::vector<int64_t>& index_at_align() const
    {
        m_index.resize(this->size());
        check(fqsb_chunk_index_at_align(m_h, m_index.data(), static_cast<int64_t>(m_index.size())));
        return m_index;
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
        int64_t ret;
        check(fqsb_time_steps_until_event(m_h, tol, static_cast<int64_t>(niter_tol),
                                          static_cast<int64_t>(max_iter), &ret));
        return static_cast<size_t>(ret);
    }

    // detail.h:1676-1893
    size_t minimise(double tol = 1e-5, size_t niter_tol = 10, size_t max_iter = 1e9,
                    bool time_activity = false, bool max_iter_is_error = true)
    {
        int64_t ret;
        check(fqsb_minimise(m_h, tol, static_cast<int64_t>(niter_tol),
                            static_cast<int64_t>(max_iter), time_activity, max_iter_is_error,
                            &ret));
        return static_cast<size_t>(ret);
    }
    size_t minimise_truncate(const std::vector<int64_t>& i_n, size_t A_truncate = 0,
                             size_t S_truncate = 0, double tol = 1e-5, size_t niter_tol = 10,
                             size_t max_iter = 1e9, bool time_activity = true,
                             bool max_iter_is_error = true)
    {
        if (i_n.size() != this->size()) {
            throw std::runtime_error("assertion failed (xt::has_shape(i_n, m_u.shape()))");
        }
        int64_t ret;
        check(fqsb_minimise_truncate(m_h, i_n.data(), static_cast<int64_t>(A_truncate),
                                     static_cast<int64_t>(S_truncate), tol,
                                     static_cast<int64_t>(niter_tol),
                                     static_cast<int64_t>(max_iter), time_activity,
                                     max_iter_is_error, &ret));
        return static_cast<size_t>(ret);
    }

    // detail.h:1901-1995
    double maxUniformDisplacement(int direction = 1)
    {
        double x;
        check(fqsb_max_uniform_displacement(m_h, direction, &x));
        return x;
    }
    double eventDrivenStep(double eps, bool kick, int direction = 1)
    {
        double x;
        check(fqsb_event_driven_step(m_h, eps, kick, direction, &x));
        return x;
    }
    void trigger(size_t p, double eps, int direction = 1)
    {
        check(fqsb_trigger(m_h, 0, static_cast<int64_t>(p), eps, direction));
    }
    void advanceToFixedForce(double f_frame, bool allow_plastic = false)
    {
        check(fqsb_advance_to_fixed_force(m_h, &f_frame, allow_plastic));
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

#undef FQSB_SHAPE2

} // namespace Line2d
} // namespace FrictionQPotSpringBlock

#endif /* FQSB_HPP */
