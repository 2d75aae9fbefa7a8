// host_staging.cpp -- see host_staging.h.  Plain C++ (g++), no CUDA.
#include "host_staging.h"

#include <sched.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace nrldpc {

struct HostPool::Impl {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    const std::function<void(size_t, size_t)> *fn = nullptr;
    size_t n = 0, piece = 0, next = 0;   // current job: [0, n) in pieces of `piece`, next unclaimed piece start
    int active = 0;                      // workers inside the current job
    uint64_t generation = 0;
    bool stop = false;

    void run_pieces(std::unique_lock<std::mutex> &lk) {
        while (next < n) {
            const size_t b = next, e = std::min(n, b + piece);
            next = e;
            lk.unlock();
            (*fn)(b, e);
            lk.lock();
        }
    }
    void worker() {
        std::unique_lock<std::mutex> lk(mu);
        uint64_t seen = 0;
        while (true) {
            cv_work.wait(lk, [&] { return stop || (generation != seen && next < n); });
            if (stop) return;
            seen = generation;
            ++active;
            run_pieces(lk);
            if (--active == 0) cv_done.notify_all();
        }
    }
};

HostPool::HostPool(int threads) : p_(new Impl) {
    for (int i = 1; i < threads; ++i) p_->workers.emplace_back([this] { p_->worker(); });
}
HostPool::~HostPool() {
    {
        std::lock_guard<std::mutex> lk(p_->mu);
        p_->stop = true;
    }
    p_->cv_work.notify_all();
    for (auto &t : p_->workers) t.join();
    delete p_;
}
int HostPool::size() const { return (int)p_->workers.size() + 1; }

void HostPool::parallel_for(size_t n, size_t grain, const std::function<void(size_t, size_t)> &fn) {
    if (n == 0) return;
    const size_t T = (size_t)size();
    if (T == 1 || n <= grain) { fn(0, n); return; }
    // about four pieces per thread (load balance against threads that wake up late), never below the grain
    size_t piece = std::max(grain, (n + 4 * T - 1) / (4 * T));
    piece = (piece + 15) & ~(size_t)15;      // keep pieces 64-byte aligned in float units
    std::unique_lock<std::mutex> lk(p_->mu);
    p_->fn = &fn; p_->n = n; p_->piece = piece; p_->next = 0;
    ++p_->generation;
    p_->cv_work.notify_all();
    ++p_->active;
    p_->run_pieces(lk);
    --p_->active;
    p_->cv_done.wait(lk, [&] { return p_->active == 0; });
    p_->fn = nullptr; p_->n = 0; p_->next = 0;
}

int default_host_threads() {
    if (const char *v = getenv("NRLDPC_HOST_THREADS")) return std::max(1, std::min(256, atoi(v)));
    int n = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
    return std::max(1, std::min(32, n));
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void narrow_avx2(const double *in, float *out, size_t n) {
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        const __m128 lo = _mm256_cvtpd_ps(_mm256_loadu_pd(in + i));       // round to nearest even (MXCSR default)
        const __m128 hi = _mm256_cvtpd_ps(_mm256_loadu_pd(in + i + 4));
        _mm256_stream_ps(out + i, _mm256_set_m128(hi, lo));
    }
    for (; i < n; ++i) out[i] = (float)in[i];
    _mm_sfence();
}
__attribute__((target("avx2"))) static void copy_avx2(const unsigned char *in, unsigned char *out, size_t n) {
    size_t i = 0;
    for (; i + 32 <= n; i += 32)
        _mm256_stream_si256(reinterpret_cast<__m256i *>(out + i), _mm256_loadu_si256(reinterpret_cast<const __m256i *>(in + i)));
    if (i < n) memcpy(out + i, in + i, n - i);
    _mm_sfence();
}
static bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}
#endif

void narrow_f64_to_f32(const double *in, float *out, size_t n) {
#if defined(__x86_64__)
    if (have_avx2() && (reinterpret_cast<uintptr_t>(out) & 31) == 0) { narrow_avx2(in, out, n); return; }
#endif
    for (size_t i = 0; i < n; ++i) out[i] = (float)in[i];
}

void copy_stream(const void *in, void *out, size_t bytes) {
#if defined(__x86_64__)
    if (have_avx2() && (reinterpret_cast<uintptr_t>(out) & 31) == 0 && bytes >= 4096) {
        copy_avx2(static_cast<const unsigned char *>(in), static_cast<unsigned char *>(out), bytes);
        return;
    }
#endif
    memcpy(out, in, bytes);
}

}  // namespace nrldpc
