// errors.cu -- version, thread-local error string, launch checking.
#include <stdarg.h>
#include <string.h>

#include "phc_common.cuh"

namespace phc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return PHC_OK;
}

int sm_count() {      // cached per DEVICE (a process may drive several GPUs from one thread)
    static int cached[64] = {0};
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;   // B200
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
}

}  // namespace phc

extern "C" int phc_version(void) { return PHC_B200_VERSION; }
extern "C" const char* phc_last_error(void) { return phc::g_err; }
