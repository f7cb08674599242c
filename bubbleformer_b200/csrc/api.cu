// Library-level plumbing of the C ABI: error string, launch counter, device properties.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace bf {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return BF_OK;
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return BF_ERR_CUDA;
}

static std::atomic<int> g_reserved_sms{0};

// SMs the kernels size their one-wave / persistent grids for: the device's SM count minus the SMs set aside with
// bf_set_reserved_sms (for the NCCL kernels of the gradient all-reduce that run next to the backward pass).
int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  const int r = g_reserved_sms.load(std::memory_order_relaxed);
  return n - r > 0 ? n - r : 1;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("BF_PDL");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::atomic<int> g_gelu_exact{-1};
bool gelu_exact() {
  int v = g_gelu_exact.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("BF_GELU_ERF");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
    g_gelu_exact.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

}  // namespace bf

extern "C" const char* bf_last_error(void) { return bf::g_err; }
extern "C" int bf_version(void) { return 100; }
extern "C" int64_t bf_launch_count(void) { return bf::g_launches.load(std::memory_order_relaxed); }
extern "C" int bf_set_gelu_mode(int exact_erf) {
  bf::g_gelu_exact.store(exact_erf ? 1 : 0, std::memory_order_relaxed);
  return BF_OK;
}
extern "C" int bf_get_gelu_mode(void) { return bf::gelu_exact() ? 1 : 0; }
extern "C" int bf_set_reserved_sms(int n) {
  if (n < 0 || n > 64) { bf::set_error("bf_set_reserved_sms: %d is outside [0, 64]", n); return BF_ERR_INVALID; }
  bf::g_reserved_sms.store(n, std::memory_order_relaxed);
  return BF_OK;
}
