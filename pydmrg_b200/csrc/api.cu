// Library-level state of libpydmrg_b200: error string, launch counter, device query.
#include "common.cuh"

#include <cstdarg>
#include <mutex>

namespace pdmrg {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace pdmrg

extern "C" const char* pdmrg_last_error(void) { return pdmrg::g_err; }
extern "C" int pdmrg_version(void) { return 100; }
extern "C" long long pdmrg_launch_count(void) { return pdmrg::g_launches.load(); }
