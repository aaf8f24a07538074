// Error reporting, ABI version and launch accounting for libplsb200.
#include "common.cuh"

namespace plsb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

char* err_buf() { return g_err; }

void set_err(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace plsb

extern "C" int plsb200_abi_version(void) { return PLSB200_ABI_VERSION; }
extern "C" const char* plsb200_last_error(void) { return plsb::err_buf(); }
extern "C" int64_t plsb200_launch_count(void) { return (int64_t)plsb::g_launches.load(std::memory_order_relaxed); }
