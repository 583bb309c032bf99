#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace pcfb {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
void set_pdl(int on);

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

static int g_pdl = -1;
bool pdl_enabled() {
    if (g_pdl < 0) {
        const char *e = getenv("PCFB_PDL");
        g_pdl = (e && e[0] == '0') ? 0 : 1;
    }
    return g_pdl != 0;
}
void set_pdl(int on) { g_pdl = on ? 1 : 0; }
}  // namespace pcfb

extern "C" const char *pcfb_last_error(void) { return pcfb::get_error(); }
extern "C" const char *pcfb_version(void) { return "pcf_b200 abi1 sm_100a"; }
extern "C" uint64_t pcfb_launch_count(void) { return pcfb::g_launches.load(); }
extern "C" int pcfb_set_pdl(int on) { const int old = pcfb::pdl_enabled() ? 1 : 0; pcfb::set_pdl(on); return old; }
