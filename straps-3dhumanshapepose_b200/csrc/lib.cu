// lib.cu -- error string, launch counter, ABI version.
#include "common.cuh"
#include "../../include/straps_b200.h"

namespace straps {
static thread_local char g_err[1024] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace straps

extern "C" const char* straps_last_error(void) { return straps::g_err; }
extern "C" int straps_abi_version(void) { return 5; }
extern "C" unsigned long long straps_launch_count(void) { return straps::g_launches.load(); }
