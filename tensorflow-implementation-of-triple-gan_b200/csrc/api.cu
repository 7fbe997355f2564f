// api.cu -- process-wide pieces of the C ABI: error string, launch counter, version.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

#include "common.cuh"

namespace tgan {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
static int pdl_default() {
  const char* e = getenv("TGAN_NO_PDL");
  return (e && atoi(e)) ? 0 : 1;
}
int g_use_pdl = pdl_default();
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace tgan

extern "C" const char* tgan_last_error(void) { return tgan::g_err; }
extern "C" int tgan_version(void) { return 100; }
extern "C" int64_t tgan_launch_count(void) { return tgan::g_launches.load(); }
extern "C" void tgan_launch_count_reset(void) { tgan::g_launches.store(0); }
