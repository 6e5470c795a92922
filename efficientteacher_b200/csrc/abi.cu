// abi.cu -- error plumbing + version for the C ABI (include/etb200.h).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <atomic>

static thread_local char g_err[512] = "";

void etb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* etb_last_error(void) { return g_err; }
extern "C" int etb_version(void) { return 100; }

static std::atomic<long long> g_launches{0};
void etb_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
// number of kernels this library has launched in this process (bench.py reports the per-step delta as gpu_launches)
extern "C" long long etb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
