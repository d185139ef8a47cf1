// devbuf.h — growable device array (cudaMalloc; contents are NOT preserved when it grows)
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

namespace shgpu {

// allocation statistics (process-wide): cudaMalloc / cudaFree synchronise the device and can take milliseconds, so the
// step driver reports when a step had to grow a buffer
struct DevBufStats { long long allocs = 0, bytes = 0; };
inline DevBufStats &devbuf_stats() { static DevBufStats s; return s; }

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  void ensure(size_t n) {
    if (n <= cap) return;
    if (p) cudaFree(p);
    // a buffer that has to grow grows by at least 50 % (first allocation: 12.5 % headroom), so that a slowly growing
    // demand (pair counts, ghost counts) does not reallocate every few rebuilds
    size_t want = std::max<size_t>(n + n / 8, cap + cap / 2);
    if (cudaMalloc(&p, want * sizeof(T)) != cudaSuccess) { p = nullptr; cap = 0; throw std::string("cudaMalloc failed"); }
    cap = want;
    devbuf_stats().allocs++; devbuf_stats().bytes += (long long)(want * sizeof(T));
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace shgpu
