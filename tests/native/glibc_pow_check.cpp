// Host instantiation of brdf_b200/csrc/glibc_pow.cuh against libm's pow() itself, bit for bit (tests/test_glibc_pow.py).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#define __device__
#include "../../brdf_b200/csrc/glibc_pow.cuh"

static bool same(double a, double b) {
    if (a != a && b != b) return true;  // any NaN equals any NaN (payloads are not part of the contract)
    unsigned long long x, y;
    memcpy(&x, &a, 8);
    memcpy(&y, &b, 8);
    return x == y;
}

int main(int argc, char** argv) {
    const long scale = argc > 1 ? atol(argv[1]) : 1;
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0, 1);
    long bad = 0, total = 0;
    auto check = [&](double x, double y) {
        const double a = brdfgpu::glibc_pow(x, y), b = pow(x, y);
        ++total;
        if (!same(a, b)) {
            if (bad < 20) printf("x=%a y=%a port=%a libm=%a\n", x, y, a, b);
            ++bad;
        }
    };
    for (long i = 0; i < 2000000 * scale; i++) check(U(rng), U(rng) * 100.0);                                  // the BRDF range: cosine ** n
    for (long i = 0; i < 500000 * scale; i++) check(std::ldexp(U(rng), -(int)(rng() % 1100)), U(rng) * 100.0);  // tiny x: underflow, subnormal results
    for (long i = 0; i < 200000 * scale; i++) { check(-U(rng), (double)(rng() % 50)); check(-U(rng), U(rng) * 10.0); }  // x < 0
    for (long i = 0; i < 200000 * scale; i++) check(U(rng) * 1000.0, (U(rng) - 0.5) * 400.0);                 // general, overflow
    for (long i = 0; i < 200000 * scale; i++) check(U(rng), std::ldexp(U(rng), -(int)(rng() % 120)));          // tiny y
    const double sp[] = {0.0, -0.0, 1.0, -1.0, 2.0, 0.5, INFINITY, -INFINITY, NAN, 1e-310, -1e-310, 5e-324, 1e308, -3.0, 3.0, 1e-20, 1e20,
                         0.9999999999999999, 1.0000000000000002, 1.0001, 1e-6};
    for (double x : sp)
        for (double y : sp) check(x, y);
    printf("total %ld bad %ld\n", total, bad);
    return bad != 0;
}
