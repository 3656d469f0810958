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

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

}  // namespace gq

extern "C" int gptq_b200_abi_version(void) { return GPTQ_B200_ABI_VERSION; }
extern "C" const char* gptq_b200_last_error(void) { return gq::g_err; }
extern "C" int64_t gptq_b200_launch_count(void) { return gq::g_launches.load(); }
