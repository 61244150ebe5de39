// Library-wide state of libmmbs: last-error string, launch counter, device check.
#include <cstdlib>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace mmbs {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool debug_sync() {
  static const bool on = []() {
    const char* e = getenv("MMBS_DEBUG_SYNC");
    return e && e[0] == '1';
  }();
  return on;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

}  // namespace mmbs

extern "C" const char* mmbs_last_error(void) { return mmbs::g_error; }
extern "C" int mmbs_version(void) { return 100; }
extern "C" int64_t mmbs_launch_count(void) { return mmbs::g_launches.load(); }

extern "C" int mmbs_device_check(void) {
  static thread_local int ok_dev = -1;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    mmbs::set_error("no CUDA device: %s (libmmbs has no CPU fallback)", cudaGetErrorString(e));
    return MMBS_ERR_DEVICE;
  }
  if (dev == ok_dev) return MMBS_OK;
  int major = 0, minor = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
    cudaGetLastError();
    mmbs::set_error("cannot query device %d", dev);
    return MMBS_ERR_DEVICE;
  }
  if (major != 10) {
    mmbs::set_error("device %d is sm_%d%d; libmmbs is built for sm_100a only", dev, major, minor);
    return MMBS_ERR_DEVICE;
  }
  ok_dev = dev;
  return MMBS_OK;
}
