// tcgen05 engine for German credit (placeholder until the tensor-core kernel lands).
#pragma once
#include <atomic>
#include <string>
#include "arp_host.cuh"
#include "arp_hmc.cuh"

namespace arp {
struct GermanTc {
  bool build(const float*, const float*, int, int, std::string*) { return true; }
  bool ready() const { return false; }
};
static inline bool german_tc_auto(long long) { return false; }
static inline int german_tc_hmc(GermanTc&, const HmcArgs&, const real*, cudaStream_t, DevBuf*, DevBuf*, DevBuf*,
                                std::atomic<long long>*, std::string* err) {
  *err = "tcgen05 engine not built";
  return 1;
}
}  // namespace arp
