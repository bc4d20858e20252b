#include "common.cuh"

#include <malloc.h>

#include <cstring>
#include <mutex>
#include <vector>

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

void keep_large_buffers_on_heap() {
  static std::once_flag once;
  std::call_once(once, [] {
    mallopt(M_MMAP_THRESHOLD, 256 << 20);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
  });
}

namespace {
struct PinnedEntry {
  void* p;
  size_t cap;
  bool in_use;
};
std::mutex g_pinned_mutex;
std::vector<PinnedEntry> g_pinned;
}  // namespace

void* pinned_alloc(size_t bytes) {
  std::lock_guard<std::mutex> lock(g_pinned_mutex);
  int best = -1;
  for (size_t i = 0; i < g_pinned.size(); ++i)
    if (!g_pinned[i].in_use && g_pinned[i].cap >= bytes && (best < 0 || g_pinned[i].cap < g_pinned[(size_t)best].cap))
      best = (int)i;
  if (best >= 0) {
    g_pinned[(size_t)best].in_use = true;
    return g_pinned[(size_t)best].p;
  }
  void* p = nullptr;
  const size_t cap = bytes + bytes / 8 + 4096;  // a little slack: batches of one job vary slightly
  if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  g_pinned.push_back(PinnedEntry{p, cap, true});
  return p;
}

void pinned_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lock(g_pinned_mutex);
  int n_free = 0;
  for (PinnedEntry& e : g_pinned) {
    if (e.p == p) e.in_use = false;
    if (!e.in_use) ++n_free;
  }
  while (n_free > 2) {  // drop the smallest idle buffer
    int victim = -1;
    for (size_t i = 0; i < g_pinned.size(); ++i)
      if (!g_pinned[i].in_use && (victim < 0 || g_pinned[i].cap < g_pinned[(size_t)victim].cap)) victim = (int)i;
    cudaFreeHost(g_pinned[(size_t)victim].p);
    g_pinned.erase(g_pinned.begin() + victim);
    --n_free;
  }
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
