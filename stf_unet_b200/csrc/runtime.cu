// Library-wide state: last error, device probe, launch counter.
#include "common.cuh"
#include <stdlib.h>
#include <atomic>
#include <string.h>

namespace stfb {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
static int g_dev_status = 1;  // 1 = unknown
static int g_num_sms = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_device() {
  if (g_dev_status != 1) {
    if (g_dev_status != STFB_OK) set_error("no sm_100 CUDA device available (libstfb200 has no CPU fallback)");
    return g_dev_status;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    g_dev_status = STFB_ENODEV;
    set_error("no CUDA device: %s (libstfb200 has no CPU fallback)", cudaGetErrorString(e));
    return g_dev_status;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  if (prop.major != 10) {
    g_dev_status = STFB_ENODEV;
    set_error("device %s is sm_%d%d; libstfb200 is built for sm_100a only", prop.name, prop.major, prop.minor);
    return g_dev_status;
  }
  g_num_sms = prop.multiProcessorCount;
  g_dev_status = STFB_OK;
  return STFB_OK;
}

int num_sms() { return g_num_sms > 0 ? g_num_sms : 148; }

// STFB_PDL: bit mask of the kernel families launch_ex() marks for programmatic dependent launch (read per call: tests switch it)
bool pdl_enabled(int family) {
  const char* e = getenv("STFB_PDL");
  return e != nullptr && (atoi(e) & family) != 0;
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int post_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return STFB_ECUDA;
  }
  count_launch(1);
  return STFB_OK;
}

}  // namespace stfb

extern "C" int stfb_version(void) { return 100; }
extern "C" const char* stfb_last_error(void) { return stfb::g_err; }
extern "C" unsigned long long stfb_launch_count(void) { return stfb::g_launches.load(); }
