#include "common.cuh"
#include <string.h>

namespace pcfb {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }
}  // namespace pcfb

extern "C" const char *pcfb_last_error(void) { return pcfb::get_error(); }
extern "C" const char *pcfb_version(void) { return "pcf_b200 abi1 sm_100a"; }
extern "C" uint64_t pcfb_launch_count(void) { return pcfb::g_launches.load(); }
