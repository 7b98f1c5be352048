// fiber_b200 — extern "C" surface (see include/fiber_b200.h) and process-wide plumbing.
#include "common.cuh"
#include "../../include/fiber_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace fiber {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
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

int gemm_dispatch(const fiber_gemm_args* a, cudaStream_t stream);

}  // namespace fiber

extern "C" {

const char* fiber_last_error(void) { return fiber::g_err; }
int fiber_version(void) { return 100; }
int64_t fiber_launch_count(void) { return fiber::g_launches.load(); }

int fiber_init(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    fiber::set_last_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
    return -2;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    fiber::set_last_error("cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
    return -2;
  }
  if (prop.major != 10) {
    fiber::set_last_error("fiber_b200 requires an sm_100a device (found sm_%d%d)", prop.major,
                          prop.minor);
    return -3;
  }
  return 0;
}

int fiber_gemm(const fiber_gemm_args* args, fiber_stream_t stream) {
  return fiber::gemm_dispatch(args, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
