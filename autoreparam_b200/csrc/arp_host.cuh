// Host-side helpers shared by the engines: RAII device buffer backed by a per-device caching pool.
//
// Every entry point of the C ABI needs scratch (workspace, staged parameters, transposed samples).
// cudaMalloc / cudaFree per call put driver-lock latency (milliseconds to hundreds of milliseconds on a
// shared host) inside every timed call, so freed blocks go to a pool (one free list per device) and are reused.
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
  // blocks are pooled PER DEVICE (the current device of the calling thread): two devices driven from one
  // process never exchange pointers
  cudaError_t take(size_t bytes, void** out, size_t* cap, int* dev) {
    cudaError_t e = cudaGetDevice(dev);
    if (e != cudaSuccess) return e;
    {
      std::lock_guard<std::mutex> g(mu_);
      auto& fl = free_[*dev];
      auto it = fl.lower_bound(bytes);
      if (it != fl.end() && it->first <= 2 * bytes + (1u << 20)) {   // best fit, at most 2x oversize
        *out = it->second; *cap = it->first;
        fl.erase(it);
        return cudaSuccess;
      }
    }
    const size_t rounded = bytes < (1u << 20) ? ((bytes + 511) / 512) * 512 : ((bytes + (1u << 20) - 1) >> 20) << 20;
    e = cudaMalloc(out, rounded);
    if (e != cudaSuccess) {            // out of memory: drop this device's cache and retry once
      cudaGetLastError();
      release(*dev);
      e = cudaMalloc(out, rounded);
    }
    *cap = rounded;
    return e;
  }
  void give(void* p, size_t cap, int dev) {
    std::lock_guard<std::mutex> g(mu_);
    free_[dev].emplace(cap, p);
  }
  // dev < 0: every device
  void release(int dev = -1) {
    std::lock_guard<std::mutex> g(mu_);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& dv : free_) {
      if (dev >= 0 && dv.first != dev) continue;
      cudaSetDevice(dv.first);
      for (auto& kv : dv.second) cudaFree(kv.second);
      dv.second.clear();
    }
    cudaSetDevice(cur);
  }
 private:
  std::mutex mu_;
  std::map<int, std::multimap<size_t, void*>> free_;
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int dev = 0;
  ~DevBuf() { reset(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  void reset() {
    if (p) DevPool::get().give(p, cap, dev);
    p = nullptr; cap = 0;
  }
  cudaError_t alloc(size_t bytes) {
    reset();
    return DevPool::get().take(bytes ? bytes : 1, &p, &cap, &dev);
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
