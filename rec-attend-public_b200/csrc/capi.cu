// Library-level entry points and error plumbing of librecattend_b200.so.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ra {

static thread_local char g_last_error[256] = "";
static unsigned long long g_launches = 0;  // kernels launched by this library (bench.py's gpu_launches)

void set_last_error(const char *what, cudaError_t e) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
}

bool pdl_enabled() {
  static const bool on = []() {
    const char *e = getenv("RA_PDL");
    return e == nullptr || atoi(e) != 0;
  }();
  return on;
}

int finish_launch(const char *what) {
  __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error(what, e);
    return RA_ERR_CUDA;
  }
  return RA_OK;
}

}  // namespace ra

extern "C" int ra_version(void) { return 10000 * 0 + 100 * 1 + 0; }

extern "C" int ra_device_count(void) {
  int n = 0;
  const cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    ra::set_last_error("cudaGetDeviceCount", e);
    return 0;
  }
  return n;
}

extern "C" const char *ra_last_error(void) { return ra::g_last_error; }

extern "C" unsigned long long ra_launch_count(void) { return __atomic_load_n(&ra::g_launches, __ATOMIC_RELAXED); }
