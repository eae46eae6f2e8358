// Library-level entry points: version, launch accounting, last error text.
#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace gs {
std::atomic<int64_t> g_launches{0};
static std::mutex g_err_mu;
static char g_err[512] = "";

void set_error(const char* what, cudaError_t e) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  std::snprintf(g_err, sizeof(g_err), "%s: %s (%d)", what, cudaGetErrorString(e), (int)e);
}
void set_error_msg(const char* what) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  std::snprintf(g_err, sizeof(g_err), "%s", what);
}
}  // namespace gs

extern "C" {
int gs_version(void) { return 100; }
int64_t gs_launch_count(void) { return gs::g_launches.load(); }
void gs_launch_count_reset(void) { gs::g_launches.store(0); }
const char* gs_last_error(void) { return gs::g_err; }
}
