// Error channel and device queries of the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace mvn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        else return 148;
    }
    return cached;
}
}  // namespace mvn

extern "C" const char* mvn_last_error(void) { return mvn::g_err; }
extern "C" int mvn_abi_version(void) { return 1; }
extern "C" int mvn_num_sms(void) { return mvn::num_sms(); }
