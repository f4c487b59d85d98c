// Error plumbing + small utility entry points of the C ABI.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace cartnet {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace cartnet

extern "C" {

int cartnet_version(void) { return 100; }

int64_t cartnet_launch_count(void) { return (int64_t)cartnet::g_launches.load(); }

const char* cartnet_last_error(void) { return cartnet::g_err; }

int cartnet_device_ok(int device) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return 0;
    return (p.major == 10) ? 1 : 0;
}

}  // extern "C"
