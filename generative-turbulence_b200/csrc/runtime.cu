// Error reporting, launch accounting and version entry points of libturbdiff_b200.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace

namespace tdb {
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace tdb

extern "C" {
const char* tdb_last_error(void) { return g_err; }
int tdb_version(void) { return 100; }
int64_t tdb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
}
