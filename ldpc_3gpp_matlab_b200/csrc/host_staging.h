// host_staging.h -- host-side staging for NRLDPC_MEM_HOST calls whose buffers the GPU cannot read at full speed:
// ordinary pageable memory (what a MEX gateway hands over: mxGetPr of a MATLAB matrix) and float64 LLRs (the
// reference's own type, NRLDPCDecoder.m:262).  Worker threads narrow / copy chunk i+1 into a pinned ring while the
// copy engine and the decode kernel work on chunk i, so PCIe carries 4 bytes per LLR from pinned memory and the
// caller's thread never runs a conversion loop.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>

namespace nrldpc {

class HostPool {
public:
    explicit HostPool(int threads);
    ~HostPool();
    HostPool(const HostPool &) = delete;
    HostPool &operator=(const HostPool &) = delete;
    int size() const;
    // fn(begin, end) over [0, n) split into contiguous pieces of at least `grain` items; the caller's thread takes a
    // piece too; returns when all pieces are done.  Small ranges run inline (no wake-up cost for one-codeword calls).
    void parallel_for(size_t n, size_t grain, const std::function<void(size_t, size_t)> &fn);

private:
    struct Impl;
    Impl *p_;
};

// NRLDPC_HOST_THREADS, else min(32, CPUs this process may run on)
int default_host_threads();

// out[i] = (float)in[i], round to nearest even (+inf stays +inf, NaN stays NaN): the same rounding as the device's
// cvt.rn.f32.f64.  Streaming stores when the destination is 32-byte aligned (the pinned ring is).
void narrow_f64_to_f32(const double *in, float *out, size_t n);
void copy_stream(const void *in, void *out, size_t bytes);

}  // namespace nrldpc
