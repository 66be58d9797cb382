// Host-logic check of cpp/phastft.hpp (CPU-only part runs without a GPU; with a GPU it also runs the
// reference's doctest lib.rs:171-178).  Built and run by tests/test_cpu_cpp_mirror.py.
#include <cmath>
#include <cstdio>
#include <cstring>
#include "phastft.hpp"

template <typename F>
static bool panics_with(F f, const char* msg) {
    try { f(); } catch (const phastft::Panic& e) { return std::strstr(e.what(), msg) != nullptr; }
    return false;
}

int main() {
    using namespace phastft;
    int fails = 0;
    // planner.rs:66 / planner.rs:195 are checked before any device is touched
    fails += !panics_with([] { PlannerDit64 p(5); }, "power of two");
    fails += !panics_with([] { PlannerDit32 p(0); }, "power of two");
    fails += !panics_with([] { PlannerR2c64 p(6); }, "n must be a power of 2 >= 4");
    fails += !panics_with([] { std::vector<float> x(2), a(2), b(2); r2c_fft_f32(x, a, b); }, "n must be a power of 2 >= 4");
    Options o = Options::guess_options(1 << 16);
    fails += !(o.multithreaded_bit_reversal && o.smallest_parallel_chunk_size == 16384);
    int ndev = 0;
    phastft_device_count(&ndev);
    if (ndev == 0) {
        fails += !panics_with([] { PlannerDit64 p(1024); }, "no CUDA device");
        std::printf("cpu-only checks: %d failures\n", fails);
        return fails;
    }
    std::vector<double> re{1, 0, 0, 0}, im(4, 0.0);
    fft_64_dit(re, im, Direction::Forward);
    for (double v : re) fails += std::fabs(v - 1.0) > 1e-12;
    fails += !panics_with([] { std::vector<double> a(16), b(8); PlannerDit64 p(16); fft_64_dit_with_planner(a, b, Direction::Forward, p); }, "reals.len() == imags.len()");
    {   // additive surface: describe / reserve / batch == loop over transforms (bit-exact below the batched-kernel threshold)
        const std::size_t n = 256, batch = 5, stride = n + 8;
        PlannerDit32 p(n);
        fails += p.num_points() != n || p.describe().find("n=2^8") == std::string::npos;
        p.reserve(batch);
        std::vector<float> bre(batch * stride), bim(batch * stride);
        for (std::size_t i = 0; i < bre.size(); ++i) { bre[i] = std::sin(0.37f * i); bim[i] = std::cos(0.11f * i); }
        std::vector<float> lre = bre, lim = bim;
        fft_32_dit_batch(bre, bim, Direction::Forward, p, batch, stride);
        for (std::size_t b = 0; b < batch; ++b) {
            std::vector<float> a(lre.begin() + b * stride, lre.begin() + b * stride + n), c(lim.begin() + b * stride, lim.begin() + b * stride + n);
            fft_32_dit_with_planner(a, c, Direction::Forward, p);
            for (std::size_t i = 0; i < n; ++i) fails += a[i] != bre[b * stride + i] || c[i] != bim[b * stride + i];
        }
    }
    std::printf("gpu checks: %d failures\n", fails);
    return fails;
}
