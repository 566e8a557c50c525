// api.cpp -- error reporting, version and launch accounting of libpcgc.
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/pcgc.h"

namespace pcgc {

static thread_local char g_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

}  // namespace pcgc

extern "C" {

int pcgc_version(void) { return 100; }
const char *pcgc_last_error(void) { return pcgc::g_error; }
uint64_t pcgc_launch_count(void) { return pcgc::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
