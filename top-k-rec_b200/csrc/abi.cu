// ABI plumbing: version, thread-local error string, launch counter.
#include "common.cuh"
#include <string.h>

namespace tkr {
static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }
}  // namespace tkr

extern "C" int tkr_version(void) { return TKR_VERSION; }
extern "C" const char* tkr_last_error(void) { return tkr::g_err; }
extern "C" int64_t tkr_launch_count(void) { return tkr::g_launches; }
extern "C" void tkr_reset_launch_count(void) { tkr::g_launches = 0; }
