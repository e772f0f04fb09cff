// Library-level bookkeeping: version string, last-error text, launch counter.
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace gr {
static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

void set_last_error(const char* what, cudaError_t e) {
  g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GAUSSREG_PDL"); v = e ? atoi(e) : 1; }
  return v != 0;
}

cudaError_t ensure_smem_attr(const void* kernel, int bytes) {
  struct Key { const void* k; int dev; int bytes; };
  static std::mutex mu;
  static std::vector<Key> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  for (const Key& k : done)
    if (k.k == kernel && k.dev == dev && k.bytes >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.push_back(Key{kernel, dev, bytes});
  return e;
}
}  // namespace gr

extern "C" const char* gr_version(void) { return "gaussreg_b200 0.1 (sm_100a, CUDA " GR_STR(__CUDACC_VER_MAJOR__) "." GR_STR(__CUDACC_VER_MINOR__) ")"; }
extern "C" const char* gr_last_error(void) { return gr::g_last_error.c_str(); }
extern "C" int64_t gr_launch_count(void) { return gr::g_launches.load(std::memory_order_relaxed); }
