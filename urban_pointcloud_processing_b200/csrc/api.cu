// Status / error plumbing and process-wide counters of the upcp_cuda C ABI.
#include <atomic>
#include <cstdio>
#include <cstring>
#include "common.cuh"

namespace upcp {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* msg, const char* file, int line) {
    const char* base = strrchr(file, '/');
    snprintf(g_err, sizeof(g_err), "%s (%s:%d)", msg, base ? base + 1 : file, line);
    return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            cached_sms = sms;
        cached_dev = dev;
    }
    return cached_sms;
}

}  // namespace upcp

extern "C" {

int upcp_abi_version(void) { return UPCP_ABI_VERSION; }

const char* upcp_last_error(void) { return upcp::g_err; }

int upcp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();  // clear sticky "no device" state
        return 0;
    }
    return n;
}

int64_t upcp_launch_count(void) { return upcp::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
