#include "common.cuh"

#include <cstring>

namespace cherry {

std::atomic<long long> g_launches{0};

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return fail(CHERRY_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

int sm_count() {
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

}  // namespace cherry

extern "C" {

const char* cherry_last_error(void) { return cherry::err_buf(); }

const char* cherry_version(void) { return "cherryml_b200 0.1 (sm_100a)"; }

int64_t cherry_launch_count(void) { return cherry::g_launches.load(); }

void cherry_reset_launch_count(void) { cherry::g_launches.store(0); }

}  // extern "C"
