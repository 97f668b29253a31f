// Library-level state: error string, launch counter, device queries.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace dusty {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
}  // namespace dusty

extern "C" {
int dusty_abi_version(void) { return DUSTY_ABI_VERSION; }
const char *dusty_last_error(void) { return dusty::g_err; }
int64_t dusty_launch_count(void) { return dusty::g_launches.load(); }
int dusty_query_sm(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    dusty::set_error("dusty_query_sm: no CUDA device");
    return DUSTY_ECUDA;
  }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}
}
