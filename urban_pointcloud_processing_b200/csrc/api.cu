// Status / error plumbing and process-wide counters of the upcp_cuda C ABI.
#include <atomic>
#include <mutex>
#include <vector>
#include <cstdio>
#include <cstring>
#include <nvtx3/nvToolsExt.h>          // header-only NVTX v3: ranges cost nothing unless a tool is attached
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

// --------------------------------------------------------------- tuning -------
static std::mutex g_tune_mu;
static upcp_tuning g_tune;
static bool g_tune_init = false;

static void tuning_defaults(upcp_tuning* t) {
    memset(t, 0, sizeof(*t));
    t->knn_tiers = 2;
    t->knn_tier1_max_block = 60000;
    t->knn_tier2_max_block = 600;
    t->knn_coarse = 2;
    t->grow_local_chain = 2;
    t->grow_ctas_per_sm = 1;
    t->knn_guess_tenths = 30;
    t->knn_guess_max_pct = 25;
}

upcp_tuning tuning() {
    std::lock_guard<std::mutex> lk(g_tune_mu);
    if (!g_tune_init) { tuning_defaults(&g_tune); g_tune_init = true; }
    return g_tune;
}

// ------------------------------------------------------------ profiling -------
struct ProfBracket { cudaEvent_t a, b; int64_t units; };
static std::mutex g_prof_mu;
static std::atomic<int> g_prof_on{0};
static std::vector<ProfBracket> g_prof[UPCP_PROF_SLOTS];
static thread_local cudaEvent_t g_prof_open[UPCP_PROF_SLOTS];   // per host thread: begin / end pair up on the thread's own stream

static const char* const kSlotNames[UPCP_PROF_SLOTS] = {
    "raster_kernel", "pip_query_kernel", "radix_sort (hist+scan+scatter)", "cc_link_kernel",
    "knn_stage (build+search+epilogue)", "grow_kernel", "knn_build", "knn_search (block tiers + per-query tier)",
    "knn_epilogue_kernel"};

void prof_begin(int slot, cudaStream_t st) {
    nvtxRangePushA(kSlotNames[slot]);          // NVTX range named like the bracket (nsys / ncu --nvtx)
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    g_prof_open[slot] = e;
}

void prof_end(int slot, cudaStream_t st, int64_t units) {
    nvtxRangePop();
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_open[slot]) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    g_prof[slot].push_back({g_prof_open[slot], e, units});
    g_prof_open[slot] = nullptr;
}

}  // namespace upcp

extern "C" {

int upcp_prof_enable(int on) { upcp::g_prof_on.store(on ? 1 : 0); return UPCP_OK; }

int upcp_prof_reset(void) {
    std::lock_guard<std::mutex> lk(upcp::g_prof_mu);
    for (int s = 0; s < UPCP_PROF_SLOTS; ++s) {
        for (auto& b : upcp::g_prof[s]) { cudaEventDestroy(b.a); cudaEventDestroy(b.b); }
        upcp::g_prof[s].clear();
    }
    return UPCP_OK;
}

int upcp_prof_read(int slot, double* total_ms, int64_t* launches, int64_t* units) {
    if (slot < 0 || slot >= UPCP_PROF_SLOTS) UPCP_FAIL(UPCP_E_INVALID, "prof_read: bad slot");
    UPCP_CUDA_OK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(upcp::g_prof_mu);
    double ms = 0.0; int64_t u = 0;
    for (auto& b : upcp::g_prof[slot]) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, b.a, b.b) == cudaSuccess) ms += t;
        u += b.units;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = (int64_t)upcp::g_prof[slot].size();
    if (units) *units = u;
    return UPCP_OK;
}

const char* upcp_prof_slot_name(int slot) {
    const char* const* names = upcp::kSlotNames;
    return (slot >= 0 && slot < UPCP_PROF_SLOTS) ? names[slot] : "";
}


void upcp_tuning_default(upcp_tuning* out) { if (out) upcp::tuning_defaults(out); }

int upcp_tuning_get(upcp_tuning* out) {
    if (!out) UPCP_FAIL(UPCP_E_INVALID, "tuning_get: NULL");
    *out = upcp::tuning();
    return UPCP_OK;
}

int upcp_tuning_set(const upcp_tuning* in) {
    if (!in) UPCP_FAIL(UPCP_E_INVALID, "tuning_set: NULL");
    if (in->knn_tiers < 1 || in->knn_tiers > 2 || in->knn_tier1_max_block < 1 || in->knn_tier1_max_block > 60000 ||
        in->knn_tier2_max_block < 1 || in->knn_tier2_max_block > 60000 || in->knn_coarse < 1 || in->knn_coarse > 16 ||
        in->grow_local_chain < 0 || in->grow_local_chain > 32 || in->grow_ctas_per_sm < 1 || in->grow_ctas_per_sm > 8 ||
        in->knn_guess_tenths < 5 || in->knn_guess_tenths > 1000 || in->knn_guess_max_pct < 0 || in->knn_guess_max_pct > 100)
        UPCP_FAIL(UPCP_E_INVALID, "tuning_set: value out of range");
    std::lock_guard<std::mutex> lk(upcp::g_tune_mu);
    upcp::g_tune = *in;
    upcp::g_tune_init = true;
    return UPCP_OK;
}

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
