// Shared helpers for the cherryml_b200 CUDA library: error reporting behind the C ABI
// and the kernel-launch counter bench.py reports as `gpu_launches`.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/cherryml_b200.h"

namespace cherry {

char* err_buf();                 // thread-local message buffer (512 bytes)
int fail(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
extern std::atomic<long long> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define CHERRY_CUDA(expr)                                      \
  do {                                                         \
    int _rc = ::cherry::check_cuda((expr), #expr);             \
    if (_rc != 0) return _rc;                                  \
  } while (0)

#define CHERRY_LAUNCH_CHECK(name)                              \
  do {                                                         \
    ::cherry::count_launch();                                  \
    int _rc = ::cherry::check_cuda(cudaGetLastError(), name);  \
    if (_rc != 0) return _rc;                                  \
  } while (0)

int sm_count();  // SMs of the current device (cached per device)

// The host-side text readers/writers allocate a few buffers of 0.1-1 MB per family (file text,
// encoded rows).  glibc serves allocations of that size by mmap/munmap, which serialises the
// worker threads on the address-space lock and page-faults every buffer in afresh (measured:
// 1.6x of the whole ingest).  Called once by the multithreaded entry points: raises glibc's
// mmap threshold so that those buffers live on the per-thread heaps.
void keep_large_buffers_on_heap();

// Page-locked host buffers for the encoded batches.  Pinning costs ~0.2 ms per MB, more than
// parsing the batch that goes into it, so released buffers are kept (at most two) and handed
// out again to the next batch that fits.  nullptr if pinned memory is unavailable.
void* pinned_alloc(size_t bytes);
void pinned_free(void* p);

}  // namespace cherry
