// devbuf.h — growable device array (cudaMalloc; contents are NOT preserved when it grows)
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

namespace shgpu {

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  void ensure(size_t n) {
    if (n <= cap) return;
    if (p) cudaFree(p);
    size_t want = std::max<size_t>(n, cap + cap / 2);
    if (cudaMalloc(&p, want * sizeof(T)) != cudaSuccess) { p = nullptr; cap = 0; throw std::string("cudaMalloc failed"); }
    cap = want;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace shgpu
