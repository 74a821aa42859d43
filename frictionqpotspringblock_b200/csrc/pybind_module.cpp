// pybind_module.cpp -- pybind11 module `_FrictionQPotSpringBlock` over the C++ host classes of
// include/fqsb.hpp, bound the way the reference binds its templates
// (/root/reference/python/main.cpp:46-226: mySystemNd + mySystemNdAthermal + mySystemNdDynamics),
// with plain numpy arrays instead of xtensor-python. The ctypes package
// (frictionqpotspringblock_b200) is the primary Python surface; this module shows that the
// reference's own binding layer compiles against the B200 host classes unchanged in structure.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "fqsb.hpp"

namespace py = pybind11;
namespace M = FrictionQPotSpringBlock;
using M::detail::System;

static py::array_t<double> as_array(const System& s, const std::vector<double>& v)
{
    std::vector<py::ssize_t> shape(s.shape().begin(), s.shape().end());
    py::array_t<double> out(shape);
    std::copy(v.begin(), v.end(), out.mutable_data());
    return out;
}

static std::vector<double> as_vector(
    const py::array_t<double, py::array::c_style | py::array::forcecast>& a)
{
    return std::vector<double>(a.data(), a.data() + a.size());
}

template <class Binder>
void mySystemNd(Binder& cls) // main.cpp:46-151
{
    cls.def_property_readonly("size", &System::size, "Number of particles");
    cls.def_property_readonly("shape", &System::shape, "Shape of the system");
    cls.def_property_readonly("dt", &System::dt, "Time step (parameter)");
    cls.def_property_readonly("mu", &System::mu, "Curvature of each well (parameter)");
    cls.def_property_readonly("eta", &System::eta, "Damping coefficient (parameter)");
    cls.def_property_readonly("m", &System::m, "Mass of each particle (parameter)");
    cls.def_property_readonly("k_frame", &System::k_frame, "Loading frame stiffness (parameter)");
    cls.def_property(
        "u", [](const System& s) { return as_array(s, s.u()); },
        [](System& s, const py::array_t<double, py::array::c_style | py::array::forcecast>& a) {
            s.set_u(as_vector(a));
        },
        "Particle slip ('positions').");
    cls.def_property(
        "v", [](const System& s) { return as_array(s, s.v()); },
        [](System& s, const py::array_t<double, py::array::c_style | py::array::forcecast>& a) {
            s.set_v(as_vector(a));
        },
        "Particle velocities.");
    cls.def_property(
        "a", [](const System& s) { return as_array(s, s.a()); },
        [](System& s, const py::array_t<double, py::array::c_style | py::array::forcecast>& a) {
            s.set_a(as_vector(a));
        },
        "Particle accelerations.");
    cls.def_property("inc", &System::inc, &System::set_inc, "Increment");
    cls.def_property("t", &System::t, &System::set_t, "Time");
    cls.def_property("u_frame", &System::u_frame, &System::set_u_frame, "Frame position");
    cls.def_property_readonly("f", [](const System& s) { return as_array(s, s.f()); });
    cls.def_property_readonly("f_potential",
                              [](const System& s) { return as_array(s, s.f_potential()); });
    cls.def_property_readonly("f_frame", [](const System& s) { return as_array(s, s.f_frame()); });
    cls.def_property_readonly("f_interactions",
                              [](const System& s) { return as_array(s, s.f_interactions()); });
    cls.def_property_readonly("f_damping",
                              [](const System& s) { return as_array(s, s.f_damping()); });
    cls.def_property_readonly("temperature", &System::temperature, "Temperature");
    cls.def_property_readonly("residual", &System::residual, "Residual");
    cls.def_property_readonly("index_at_align", [](const System& s) {
        const auto& i = s.index_at_align();
        std::vector<py::ssize_t> shape(s.shape().begin(), s.shape().end());
        py::array_t<int64_t> out(shape);
        std::copy(i.begin(), i.end(), out.mutable_data());
        return out;
    });
    cls.def_property_readonly(
        "chunk", [](System& s) -> M::detail::Chunk& { return s.chunk(); },
        py::return_value_policy::reference_internal, "Chunk of random numbers"); // main.cpp:65-70
    cls.def("refresh", &System::refresh, "refresh");
    cls.def("quench", &System::quench, "quench");
    cls.def("maxUniformDisplacement", &System::maxUniformDisplacement, py::arg("direction") = 1);
    cls.def("trigger", &System::trigger, py::arg("p"), py::arg("eps"), py::arg("direction") = 1);
    cls.def("advanceToFixedForce", &System::advanceToFixedForce, py::arg("f_frame"),
            py::arg("allow_plastic") = false);
}

template <class Binder>
void mySystemNdAthermal(Binder& cls) // main.cpp:153-203
{
    cls.def("minimise", &System::minimise, py::arg("tol") = 1e-5, py::arg("niter_tol") = 10,
            py::arg("max_iter") = size_t(1e9), py::arg("time_activity") = false,
            py::arg("max_iter_is_error") = true, py::call_guard<py::gil_scoped_release>());
    cls.def(
        "minimise_truncate",
        [](System& s, const py::array_t<int64_t, py::array::c_style | py::array::forcecast>& i_n,
           size_t A_truncate, size_t S_truncate, double tol, size_t niter_tol, size_t max_iter,
           bool time_activity, bool max_iter_is_error) {
            std::vector<int64_t> v(i_n.data(), i_n.data() + i_n.size());
            return s.minimise_truncate(v, A_truncate, S_truncate, tol, niter_tol, max_iter,
                                       time_activity, max_iter_is_error);
        },
        py::arg("i_n"), py::arg("A_truncate") = 0, py::arg("S_truncate") = 0,
        py::arg("tol") = 1e-5, py::arg("niter_tol") = 10, py::arg("max_iter") = size_t(1e9),
        py::arg("time_activity") = true, py::arg("max_iter_is_error") = true);
    cls.def("eventDrivenStep", &System::eventDrivenStep, py::arg("eps"), py::arg("kick"),
            py::arg("direction") = 1);
    cls.def_property_readonly("quasistaticActivityFirst", &System::quasistaticActivityFirst);
    cls.def_property_readonly("quasistaticActivityLast", &System::quasistaticActivityLast);
}

template <class Binder>
void mySystemNdDynamics(Binder& cls) // main.cpp:211-226
{
    cls.def("timeStep", &System::timeStep, "timeStep");
    cls.def("timeSteps", &System::timeSteps, py::arg("n"),
            py::call_guard<py::gil_scoped_release>());
    cls.def("timeStepsUntilEvent", &System::timeStepsUntilEvent, py::arg("tol") = 1e-5,
            py::arg("niter_tol") = 10, py::arg("max_iter") = size_t(1e9));
    cls.def("flowSteps", &System::flowSteps, py::arg("n"), py::arg("v_frame"));
}

template <class Binder, class S>
void mySystemNdExternal(Binder& cls) // main.cpp:205-209
{
    cls.def_property_readonly(
        "external", [](S& s) -> M::detail::RandomNormalForcing& { return s.external(); },
        py::return_value_policy::reference_internal, "Class adding external force");
}

// constructor argument lists of main.cpp:493-520 etc.
#define FQSB_COMMON_ARGS \
    py::arg("shape"), py::arg("seed"), py::arg("distribution"), py::arg("parameters"), \
        py::arg("offset") = -100.0, py::arg("nchunk") = 5000

PYBIND11_MODULE(_FrictionQPotSpringBlock, m)
{
    m.doc() = "Spring-block friction model with local disordered potential energy landscape "
              "(B200-native engine behind the reference's binding layer)";
    m.def("version", &M::version, "Return version string.");
    py::class_<System> base(m, "System");
    mySystemNd(base);
    mySystemNdAthermal(base);

    { // main.cpp:246-276
        py::module sm = m.def_submodule("detail", "detail");
        { // the prrng::pcg32_tensor_cumsum surface the systems expose as `system.chunk`
            using Ch = M::detail::Chunk;
            py::class_<Ch> ch(sm, "Chunk");
            auto ints = [](const std::vector<int64_t>& v) {
                return py::array_t<int64_t>(static_cast<py::ssize_t>(v.size()), v.data());
            };
            auto dbls = [](const std::vector<double>& v) {
                return py::array_t<double>(static_cast<py::ssize_t>(v.size()), v.data());
            };
            ch.def_property_readonly("chunk_size", &Ch::chunk_size);
            ch.def_property_readonly("index_at_align",
                                     [ints](const Ch& c) { return ints(c.index_at_align()); });
            ch.def_property_readonly(
                "chunk_index_at_align",
                [ints](const Ch& c) { return ints(c.chunk_index_at_align()); });
            ch.def_property_readonly("start", [ints](const Ch& c) { return ints(c.start()); });
            ch.def_property_readonly("left_of_align",
                                     [dbls](const Ch& c) { return dbls(c.left_of_align()); });
            ch.def_property_readonly("right_of_align",
                                     [dbls](const Ch& c) { return dbls(c.right_of_align()); });
            ch.def_property_readonly("data", [](const Ch& c) {
                const auto& v = c.data();
                const py::ssize_t nc = static_cast<py::ssize_t>(c.chunk_size());
                py::array_t<double> out({static_cast<py::ssize_t>(v.size()) / nc, nc});
                std::copy(v.begin(), v.end(), out.mutable_data());
                return out;
            });
            ch.def("align",
                   [](Ch& c, const py::array_t<double, py::array::c_style | py::array::forcecast>& u) {
                       c.align(as_vector(u));
                   },
                   py::arg("u"));
            ch.def("state_at",
                   [](const Ch& c,
                      const py::array_t<int64_t, py::array::c_style | py::array::forcecast>& index) {
                       const auto& v = c.state_at(
                           std::vector<int64_t>(index.data(), index.data() + index.size()));
                       return py::array_t<uint64_t>(static_cast<py::ssize_t>(v.size()), v.data());
                   },
                   py::arg("index"));
            ch.def("restore",
                   [](Ch& c,
                      const py::array_t<uint64_t, py::array::c_style | py::array::forcecast>& state,
                      const py::array_t<double, py::array::c_style | py::array::forcecast>& value,
                      const py::array_t<int64_t, py::array::c_style | py::array::forcecast>& index) {
                       c.restore(std::vector<uint64_t>(state.data(), state.data() + state.size()),
                                 as_vector(value),
                                 std::vector<int64_t>(index.data(), index.data() + index.size()));
                   },
                   py::arg("state"), py::arg("value"), py::arg("index"));
        }
        using S = M::detail::RandomNormalForcing;
        py::class_<S> cls(sm, "RandomNormalForcing_1");
        cls.def_property("state", &S::state, &S::set_state, "State of RNG");
        cls.def_property(
            "f_thermal",
            [](const S& s) {
                const auto& v = s.f_thermal();
                return py::array_t<double>(static_cast<py::ssize_t>(v.size()), v.data());
            },
            [](S& s, const py::array_t<double, py::array::c_style | py::array::forcecast>& a) {
                s.set_f_thermal(as_vector(a));
            },
            "Random force");
        cls.def_property(
            "next",
            [](const S& s) {
                const auto& v = s.next();
                return py::array_t<int64_t>(static_cast<py::ssize_t>(v.size()), v.data());
            },
            [](S& s, const py::array_t<int64_t, py::array::c_style | py::array::forcecast>& a) {
                s.set_next(std::vector<int64_t>(a.data(), a.data() + a.size()));
            },
            "Next draw increment");
        cls.def("__repr__", [](const S&) {
            return "<FrictionQPotSpringBlock.detail.RandomNormalForcing_1>";
        });
    }

    {
        py::module sm = m.def_submodule("Line1d", "Line1d");
        namespace SM = M::Line1d;
        using S1 = const std::array<size_t, 1>&;
        using Str = const std::string&;
        using Par = const std::vector<double>&;
        {
            py::class_<SM::System_Cuspy_Laplace, System> cls(sm, "System_Cuspy_Laplace");
            cls.def(py::init<double, double, double, double, double, double, S1, uint64_t, Str,
                             Par, double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("k_interactions"),
                    py::arg("k_frame"), py::arg("dt"), FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
        }
        {
            py::class_<SM::System_Cuspy_Laplace_Nopassing, System> cls(
                sm, "System_Cuspy_Laplace_Nopassing");
            cls.def(py::init<double, double, double, S1, uint64_t, Str, Par, double, size_t,
                             double, double>(),
                    py::arg("mu"), py::arg("k_interactions"), py::arg("k_frame"),
                    FQSB_COMMON_ARGS, py::arg("eta") = 0.0, py::arg("dt") = 0.0);
        }
        {
            py::class_<SM::System_SemiSmooth_Laplace, System> cls(sm, "System_SemiSmooth_Laplace");
            cls.def(py::init<double, double, double, double, double, double, double, S1, uint64_t,
                             Str, Par, double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("kappa"),
                    py::arg("k_interactions"), py::arg("k_frame"), py::arg("dt"),
                    FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
        }
        {
            py::class_<SM::System_Smooth_Laplace, System> cls(sm, "System_Smooth_Laplace");
            cls.def(py::init<double, double, double, double, double, double, S1, uint64_t, Str,
                             Par, double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("k_interactions"),
                    py::arg("k_frame"), py::arg("dt"), FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
        }
        {
            py::class_<SM::System_Cuspy_Quartic, System> cls(sm, "System_Cuspy_Quartic");
            cls.def(py::init<double, double, double, double, double, double, double, S1, uint64_t,
                             Str, Par, double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("a1"), py::arg("a2"),
                    py::arg("k_frame"), py::arg("dt"), FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
        }
        {
            py::class_<SM::System_Cuspy_QuarticGradient, System> cls(
                sm, "System_Cuspy_QuarticGradient");
            cls.def(py::init<double, double, double, double, double, double, double, S1, uint64_t,
                             Str, Par, double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("k2"), py::arg("k4"),
                    py::arg("k_frame"), py::arg("dt"), FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
        }
        using Inc = const std::vector<int64_t>&;
        { // main.cpp:570-622
            using S = SM::System_Cuspy_Laplace_RandomForcing;
            py::class_<S, System> cls(sm, "System_Cuspy_Laplace_RandomForcing");
            cls.def(py::init<double, double, double, double, double, double, double, double,
                             uint64_t, Inc, Inc, S1, uint64_t, Str, Par, double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("k_interactions"),
                    py::arg("k_frame"), py::arg("dt"), py::arg("mean"), py::arg("stddev"),
                    py::arg("seed_forcing"), py::arg("dinc_init"), py::arg("dinc"),
                    FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
            mySystemNdExternal<decltype(cls), S>(cls);
        }
        { // main.cpp:623-677
            using S = SM::System_Cuspy_Quartic_RandomForcing;
            py::class_<S, System> cls(sm, "System_Cuspy_Quartic_RandomForcing");
            cls.def(py::init<double, double, double, double, double, double, double, double,
                             double, uint64_t, Inc, Inc, S1, uint64_t, Str, Par, double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("a1"), py::arg("a2"),
                    py::arg("k_frame"), py::arg("dt"), py::arg("mean"), py::arg("stddev"),
                    py::arg("seed_forcing"), py::arg("dinc_init"), py::arg("dinc"),
                    FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
            mySystemNdExternal<decltype(cls), S>(cls);
        }
        {
            py::class_<SM::System_Cuspy_LongRange, System> cls(sm, "System_Cuspy_LongRange");
            cls.def(py::init<double, double, double, double, double, double, double, S1, uint64_t,
                             Str, Par, double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("k_interactions"),
                    py::arg("alpha"), py::arg("k_frame"), py::arg("dt"), FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
        }
    }
    { // main.cpp:279-470
        py::module sm = m.def_submodule("Particles", "Particles");
        namespace SM = M::Particles;
        using S1 = const std::array<size_t, 1>&;
        using Str = const std::string&;
        using Par = const std::vector<double>&;
        using Inc = const std::vector<int64_t>&;
        {
            py::class_<SM::System_Cuspy, System> cls(sm, "System_Cuspy");
            cls.def(py::init<double, double, double, double, double, S1, uint64_t, Str, Par,
                             double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("k_frame"),
                    py::arg("dt"), FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
        }
        {
            using S = SM::System_Cuspy_RandomForcing;
            py::class_<S, System> cls(sm, "System_Cuspy_RandomForcing");
            cls.def(py::init<double, double, double, double, double, double, double, uint64_t,
                             Inc, Inc, S1, uint64_t, Str, Par, double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("k_frame"),
                    py::arg("dt"), py::arg("mean"), py::arg("stddev"), py::arg("seed_forcing"),
                    py::arg("dinc_init"), py::arg("dinc"), FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
            mySystemNdExternal<decltype(cls), S>(cls);
        }
        {
            py::class_<SM::System_SemiSmooth, System> cls(sm, "System_SemiSmooth");
            cls.def(py::init<double, double, double, double, double, double, S1, uint64_t, Str,
                             Par, double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("kappa"),
                    py::arg("k_frame"), py::arg("dt"), FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
        }
        {
            py::class_<SM::System_Smooth, System> cls(sm, "System_Smooth");
            cls.def(py::init<double, double, double, double, double, S1, uint64_t, Str, Par,
                             double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("k_frame"),
                    py::arg("dt"), FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
        }
    }
    {
        py::module sm = m.def_submodule("Line2d", "Line2d");
        namespace SM = M::Line2d;
        using S2 = const std::array<size_t, 2>&;
        using Str = const std::string&;
        using Par = const std::vector<double>&;
        {
            py::class_<SM::System_Cuspy_Laplace, System> cls(sm, "System_Cuspy_Laplace");
            cls.def(py::init<double, double, double, double, double, double, S2, uint64_t, Str,
                             Par, double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("k_interactions"),
                    py::arg("k_frame"), py::arg("dt"), FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
        }
        {
            py::class_<SM::System_Cuspy_QuarticGradient, System> cls(
                sm, "System_Cuspy_QuarticGradient");
            cls.def(py::init<double, double, double, double, double, double, double, S2, uint64_t,
                             Str, Par, double, size_t>(),
                    py::arg("m"), py::arg("eta"), py::arg("mu"), py::arg("k2"), py::arg("k4"),
                    py::arg("k_frame"), py::arg("dt"), FQSB_COMMON_ARGS);
            mySystemNdDynamics(cls);
        }
        { // 2-D generalisation of Line1d.System_Cuspy_Laplace_Nopassing (new)
            py::class_<SM::System_Cuspy_Laplace_Nopassing, System> cls(
                sm, "System_Cuspy_Laplace_Nopassing");
            cls.def(py::init<double, double, double, S2, uint64_t, Str, Par, double, size_t,
                             double, double>(),
                    py::arg("mu"), py::arg("k_interactions"), py::arg("k_frame"),
                    FQSB_COMMON_ARGS, py::arg("eta") = 0.0, py::arg("dt") = 0.0);
        }
    }
}
