// C++ rendering of /root/reference/examples/Line1d_Cuspy_Laplace.py:17-61 against include/fqsb.hpp:
// the quasistatic event-driven protocol with the reference's class names and signatures.
// Prints "step u_frame mean(f_frame) S" per step; tests/test_cpp_host.py compares with the golden.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>

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
