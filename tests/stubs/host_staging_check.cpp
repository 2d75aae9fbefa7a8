// host_staging_check.cpp -- CPU self-check of ldpc_3gpp_matlab_b200/csrc/host_staging.cpp (built and run by
// tests/test_host_staging.py): the pool covers every index exactly once for many (n, grain, threads), the float64 ->
// float32 narrowing equals a plain cast (round to nearest even, +-inf, NaN, overflow to inf, subnormals), and it prints
// the narrowing bandwidth of this host.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "host_staging.h"

int main(int argc, char **argv) {
    using namespace nrldpc;
    int bad = 0;
    for (int threads : {1, 2, 5, 16}) {
        HostPool pool(threads);
        for (size_t n : {size_t(0), size_t(1), size_t(17), size_t(4096), size_t(100003), size_t(1) << 20}) {
            for (size_t grain : {size_t(1), size_t(64), size_t(1) << 16}) {
                std::vector<unsigned char> hit(n, 0);
                for (int rep = 0; rep < 3; ++rep)
                    pool.parallel_for(n, grain, [&](size_t b, size_t e) { for (size_t i = b; i < e; ++i) ++hit[i]; });
                for (size_t i = 0; i < n; ++i) if (hit[i] != 3) { ++bad; break; }
            }
        }
    }
    const size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : (size_t(1) << 24);
    std::vector<double> in(n);
    unsigned long long x = 88172645463325252ull;
    for (size_t i = 0; i < n; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        double v;
        memcpy(&v, &x, 8);                       // arbitrary bit patterns: all exponents, NaNs, infinities
        in[i] = (i & 3) ? (double)(long long)(x >> 20) * 1e-9 - 4000.0 : v;
    }
    in[0] = std::numeric_limits<double>::infinity(); in[1] = -in[0]; in[2] = std::nan(""); in[3] = 1e300; in[4] = -0.0;
    in[5] = 1.0 + std::ldexp(1.0, -24);          // exactly half way between two floats: ties to even
    in[6] = 1e-45;
    float *out = static_cast<float *>(aligned_alloc(64, n * sizeof(float)));
    HostPool pool(default_host_threads());
    auto t0 = std::chrono::steady_clock::now();
    const int reps = 5;
    for (int r = 0; r < reps; ++r)
        pool.parallel_for(n, size_t(1) << 16, [&](size_t b, size_t e) { narrow_f64_to_f32(in.data() + b, out + b, e - b); });
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
    for (size_t i = 0; i < n; ++i) {
        const float ref = (float)in[i];
        if (memcmp(&ref, &out[i], 4) != 0 && !(ref != ref && out[i] != out[i])) { ++bad; break; }
    }
    std::vector<unsigned char> src(n), dst(n + 64);
    for (size_t i = 0; i < n; ++i) src[i] = (unsigned char)(i * 131u);
    unsigned char *d = dst.data() + (32 - (reinterpret_cast<uintptr_t>(dst.data()) & 31)) % 32;
    pool.parallel_for(n, size_t(1) << 18, [&](size_t b, size_t e) { copy_stream(src.data() + b, d + b, e - b); });
    if (memcmp(src.data(), d, n) != 0) ++bad;
    printf("{\"threads\": %d, \"narrow_GBps_read\": %.2f, \"bad\": %d}\n", pool.size(), n * 8 / dt / 1e9, bad);
    free(out);
    return bad ? 1 : 0;
}
