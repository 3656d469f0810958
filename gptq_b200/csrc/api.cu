// Library-level plumbing of the C ABI: error strings, launch counter, device properties.
#include "common.cuh"
#include <stdarg.h>

namespace gq {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cur_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    return dev;
}

int num_sms() {
    static std::atomic<int> sms[256] = {};
    const int dev = cur_device();
    const int slot = (dev >= 0 && dev < 256) ? dev : 0;
    int v = sms[slot].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms[slot].store(v, std::memory_order_relaxed);
    }
    return v;
}

}  // namespace gq

extern "C" int gptq_b200_abi_version(void) { return GPTQ_B200_ABI_VERSION; }
extern "C" const char* gptq_b200_last_error(void) { return gq::g_err; }
extern "C" int64_t gptq_b200_launch_count(void) { return gq::g_launches.load(); }
