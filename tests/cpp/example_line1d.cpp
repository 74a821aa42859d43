// C++ rendering of /root/reference/examples/Line1d_Cuspy_Laplace.py:17-61 against include/fqsb.hpp:
// the quasistatic event-driven protocol with the reference's class names and signatures.
// Prints "step u_frame mean(f_frame) S" per step; tests/test_cpp_host.py compares with the golden.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <array>
#include <memory>
#include <numeric>
#include <vector>

#include "fqsb.hpp"

int main(int argc, char** argv)
{
    namespace model = FrictionQPotSpringBlock::Line1d;
    const size_t N = 1000;
    const double xdelta = 1e-3;
    const int nstep = argc > 1 ? std::atoi(argv[1]) : 40;
    try {
        model::System_Cuspy_Laplace system(
            1.0, 2.0 * std::sqrt(3.0) / 10.0, 1.0, 1.0, 1.0 / static_cast<double>(N), 0.1, {N}, 0,
            "random", {2.0}, -50.0);
        for (int step = 0; step < nstep; ++step) {
            std::vector<int64_t> i_n = system.index_at_align();
            if (step == 0) {
                system.set_u_frame(0.0);
            }
            else {
                system.eventDrivenStep(xdelta, step % 2 == 0);
            }
            if (step % 2 == 0) {
                if (system.minimise() != 0) {
                    return 2;
                }
            }
            const auto& ff = system.f_frame();
            double mean = std::accumulate(ff.begin(), ff.end(), 0.0) / static_cast<double>(N);
            const auto& i = system.index_at_align();
            long long S = 0;
            for (size_t p = 0; p < N; ++p) {
                S += i[p] - i_n[p];
            }
            std::printf("%d %.17g %.17g %lld\n", step, system.u_frame(), mean, S);
        }
        // system.chunk() (detail.h:1146-1149): the reference's generator object
        {
            const auto& ch = system.chunk();
            const std::vector<int64_t> i = ch.index_at_align();
            const std::vector<double> yl = ch.left_of_align(), yr = ch.right_of_align();
            const std::vector<int64_t> start = ch.start();
            const auto& data = ch.data(); // [N][chunk_size]
            const auto& u = system.u();
            bool ok = true;
            for (size_t p = 0; p < N; ++p) {
                const size_t c = static_cast<size_t>(i[p] - start[p]);
                ok = ok && yl[p] < u[p] && u[p] <= yr[p];
                ok = ok && data[p * ch.chunk_size() + c] == yl[p];
                ok = ok && data[p * ch.chunk_size() + c + 1] == yr[p];
            }
            std::printf("chunk %s\n", ok ? "ok" : "MISMATCH");
        }
        // Ensemble<S>: 3 realisations in one handle; realisation r is the system seeded seed + r*N
        {
            namespace F = FrictionQPotSpringBlock;
            F::detail::Options opt;
            opt.nrealisations = 3;
            F::Ensemble<model::System_Cuspy_Laplace> ens(
                opt, 1.0, 2.0 * std::sqrt(3.0) / 10.0, 1.0, 1.0, 1.0 / static_cast<double>(N), 0.1,
                std::array<size_t, 1>{N}, 0, "random", std::vector<double>{2.0}, -50.0);
            model::System_Cuspy_Laplace single(
                1.0, 2.0 * std::sqrt(3.0) / 10.0, 1.0, 1.0, 1.0 / static_cast<double>(N), 0.1, {N},
                2 * N, "random", {2.0}, -50.0);
            ens.mark_indices();
            bool ok = true;
            for (auto r : ens.minimise_all()) {
                ok = ok && r == 0;
            }
            ok = ok && single.minimise() == 0;
            ens.eventDrivenStep_all(xdelta, false);
            ens.eventDrivenStep_all(xdelta, true);
            single.eventDrivenStep(xdelta, false);
            single.eventDrivenStep(xdelta, true);
            ens.minimise_all();
            single.minimise();
            std::vector<int64_t> S, A;
            ens.avalanche_since_mark(S, A);
            const auto& ue = ens.u();
            const auto& us = single.u();
            for (size_t p = 0; p < N; ++p) {
                ok = ok && ue[2 * N + p] == us[p];
            }
            ok = ok && ens.inc_all()[2] == single.inc() && ens.u_frame_all()[2] == single.u_frame();
            std::printf("ensemble %s S=%lld,%lld,%lld\n", ok ? "ok" : "MISMATCH",
                        static_cast<long long>(S[0]), static_cast<long long>(S[1]),
                        static_cast<long long>(S[2]));
        }
        // Slab<S, 1>: the same line over two members of this process (both on device 0 here)
        {
            namespace F = FrictionQPotSpringBlock;
            auto make = [&](const std::array<size_t, 1>& local) {
                return std::make_unique<model::System_Cuspy_Laplace>(
                    1.0, 2.0 * std::sqrt(3.0) / 10.0, 1.0, 1.0, 1.0 / static_cast<double>(N), 0.1,
                    local, 0, "random", std::vector<double>{2.0}, -50.0);
            };
            F::Slab<model::System_Cuspy_Laplace, 1> slab({0, 0}, 16, {N}, make);
            model::System_Cuspy_Laplace one(1.0, 2.0 * std::sqrt(3.0) / 10.0, 1.0, 1.0,
                                            1.0 / static_cast<double>(N), 0.1, {N}, 0, "random",
                                            {2.0}, -50.0);
            bool ok = slab.minimise() == 0 && one.minimise() == 0;
            slab.mark_indices();
            const std::vector<int64_t> i_n = one.index_at_align();
            slab.eventDrivenStep(xdelta, false);
            slab.eventDrivenStep(xdelta, true);
            one.eventDrivenStep(xdelta, false);
            one.eventDrivenStep(xdelta, true);
            ok = ok && slab.minimise() == 0 && one.minimise() == 0;
            int64_t S = 0, A = 0;
            slab.avalanche_since_mark(S, A);
            long long S1 = 0;
            const auto& i = one.index_at_align();
            for (size_t p = 0; p < N; ++p) {
                S1 += i[p] - i_n[p];
            }
            ok = ok && S == S1 && std::fabs(slab.u_frame() - one.u_frame()) <= 1e-12 * one.u_frame();
            std::printf("slab %s S=%lld\n", ok ? "ok" : "MISMATCH", static_cast<long long>(S));
        }
        // error convention: std::runtime_error with the reference's text
        try {
            system.minimise(2.0);
            return 3;
        }
        catch (const std::runtime_error& e) {
            std::printf("error %s\n", e.what());
        }
    }
    catch (const std::runtime_error& e) {
        std::fprintf(stderr, "runtime_error: %s\n", e.what());
        return 1;
    }
    return 0;
}
