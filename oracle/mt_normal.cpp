// TEST INFRASTRUCTURE — not product code.
// Exposes the libstdc++ std::mt19937 + std::normal_distribution<double> stream
// that the reference CPU backend uses for its start vectors
// (/root/reference/Impl/chase_cpu/chase_cpu.hpp:296-309) to numpy.
#include <random>
#include <cstddef>
extern "C" void mt_normal_fill(unsigned seed, std::size_t n, double* out)
{
    std::mt19937 gen(seed);
    std::normal_distribution<> d;
    for (std::size_t i = 0; i < n; ++i)
        out[i] = d(gen);
}
// continue an existing stream: skip `skip` draws first
extern "C" void mt_normal_fill_skip(unsigned seed, std::size_t skip, std::size_t n, double* out)
{
    std::mt19937 gen(seed);
    std::normal_distribution<> d;
    for (std::size_t i = 0; i < skip; ++i)
        (void)d(gen);
    for (std::size_t i = 0; i < n; ++i)
        out[i] = d(gen);
}
