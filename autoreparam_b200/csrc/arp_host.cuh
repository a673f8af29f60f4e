// Host-side helpers shared by the engines: RAII device buffer backed by a process-wide caching pool.
//
// Every entry point of the C ABI needs scratch (workspace, staged parameters, transposed samples).
// cudaMalloc / cudaFree per call put driver-lock latency (milliseconds to hundreds of milliseconds on a
// shared host) inside every timed call, so freed blocks go to a pool and are reused by later calls.
// All entry points synchronise their stream before returning, so a pooled block is never in flight.
// arp_release_cached_memory() returns the pool to the driver.
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <mutex>
#include <vector>

namespace arp {

class DevPool {
 public:
  static DevPool& get() { static DevPool p; return p; }
  cudaError_t take(size_t bytes, void** out, size_t* cap) {
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = free_.lower_bound(bytes);
      if (it != free_.end() && it->first <= 2 * bytes + (1u << 20)) {   // best fit, at most 2x oversize
        *out = it->second; *cap = it->first;
        free_.erase(it);
        return cudaSuccess;
      }
    }
    const size_t rounded = bytes < (1u << 20) ? ((bytes + 511) / 512) * 512 : ((bytes + (1u << 20) - 1) >> 20) << 20;
    cudaError_t e = cudaMalloc(out, rounded);
    if (e != cudaSuccess) {            // out of memory: drop the cache and retry once
      cudaGetLastError();
      release();
      e = cudaMalloc(out, rounded);
    }
    *cap = rounded;
    return e;
  }
  void give(void* p, size_t cap) {
    std::lock_guard<std::mutex> g(mu_);
    free_.emplace(cap, p);
  }
  void release() {
    std::lock_guard<std::mutex> g(mu_);
    for (auto& kv : free_) cudaFree(kv.second);
    free_.clear();
  }
 private:
  std::mutex mu_;
  std::multimap<size_t, void*> free_;
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  ~DevBuf() { reset(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  void reset() {
    if (p) DevPool::get().give(p, cap);
    p = nullptr; cap = 0;
  }
  cudaError_t alloc(size_t bytes) {
    reset();
    return DevPool::get().take(bytes ? bytes : 1, &p, &cap);
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
