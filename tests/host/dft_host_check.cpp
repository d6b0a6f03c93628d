// Host check of the in-register DFT templates (csrc/sml_dft.cuh) against a float64 naive DFT.
// Built and run by tests/test_dft_host.py with g++ -std=c++17 (no GPU needed).
#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include "sml_dft.cuh"

template <int N, int DIR>
double check(std::mt19937& rng) {
    std::normal_distribution<float> nd(0.f, 1.f);
    sml::cf v[N];
    std::complex<double> in[N];
    for (int i = 0; i < N; ++i) {
        v[i].re = nd(rng);
        v[i].im = nd(rng);
        in[i] = {v[i].re, v[i].im};
    }
    sml::Dft<N, DIR>::run(v);
    double num = 0, den = 0;
    for (int k = 0; k < N; ++k) {
        std::complex<double> s = 0;
        for (int n = 0; n < N; ++n) s += in[n] * std::polar(1.0, DIR * 2.0 * M_PI * n * k / N);
        std::complex<double> d = s - std::complex<double>(v[k].re, v[k].im);
        num += std::norm(d);
        den += std::norm(s);
    }
    return std::sqrt(num / den);
}

int main() {
    std::mt19937 rng(123);
    double worst = 0;
#define RUN(N)                                                         \
    {                                                                  \
        double a = check<N, -1>(rng), b = check<N, +1>(rng);           \
        std::printf("N=%d fwd=%.3e inv=%.3e\n", N, a, b);              \
        worst = std::fmax(worst, std::fmax(a, b));                     \
    }
    RUN(2) RUN(4) RUN(8) RUN(16) RUN(32) RUN(64)
    std::printf("WORST %.3e\n", worst);
    return worst < 5e-7 ? 0 : 1;
}
