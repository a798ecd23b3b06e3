// Library-wide state of libver_b200.so: error string, launch counter, device queries.
#include <stdarg.h>

#include "common.cuh"

std::atomic<int64_t> g_ver_launches{0};

static thread_local char t_error[512] = "";

void ver_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

int ver_device_sm_count() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

int ver_device_max_smem_optin() {
    int dev = 0, n = 227 * 1024;
    if (cudaGetDevice(&dev) == cudaSuccess)
        cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return n;
}

extern "C" int ver_abi_version(void) { return 7; }
extern "C" const char* ver_last_error(void) { return t_error; }
extern "C" int64_t ver_launch_count(void) { return g_ver_launches.load(); }
