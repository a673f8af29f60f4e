// Host-side helpers shared by the engines: RAII device buffer.
#pragma once
#include <cuda_runtime.h>
#include <vector>

namespace arp {

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  cudaError_t alloc(size_t bytes) {
    if (p) { cudaFree(p); p = nullptr; }
    return cudaMalloc(&p, bytes ? bytes : 1);
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

template <typename T>
static inline cudaError_t upload(DevBuf& buf, const std::vector<T>& h) {
  cudaError_t e = buf.alloc(h.size() * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(buf.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

}  // namespace arp
