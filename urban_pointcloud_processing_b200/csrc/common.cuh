// Shared helpers for the upcp_cuda translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/upcp_cuda.h"

namespace upcp {

// Thread-local error message + status plumbing (defined in api.cu).
int set_error(int code, const char* msg, const char* file, int line);
void count_launch(int n = 1);
int num_sms();
upcp_tuning tuning();        // copy of the process-wide tuning knobs (upcp_tuning_set)

#define UPCP_FAIL(code, msg) return ::upcp::set_error((code), (msg), __FILE__, __LINE__)

#define UPCP_CUDA_OK(expr)                                                         \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess)                                                     \
            return ::upcp::set_error(UPCP_E_CUDA, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define UPCP_LAUNCH_OK()                                                           \
    do {                                                                           \
        ::upcp::count_launch();                                                    \
        cudaError_t _e = cudaGetLastError();                                       \
        if (_e != cudaSuccess)                                                     \
            return ::upcp::set_error(UPCP_E_CUDA, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define UPCP_TRY(expr)                                                             \
    do {                                                                           \
        int _s = (expr);                                                           \
        if (_s != UPCP_OK) return _s;                                              \
    } while (0)

// Optional cudaEvent bracket around a kernel (see upcp_prof_* in the header).  `units` is
// the number of points / cells the launch processes (for the roofline arithmetic).
void prof_begin(int slot, cudaStream_t st);
void prof_end(int slot, cudaStream_t st, int64_t units);

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Bump allocator over the caller-provided workspace.
struct Workspace {
    char* base;
    size_t size;
    size_t off;
    bool ok;
    Workspace(void* p, size_t n) : base((char*)p), size(n), off(0), ok(true) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = align_up(count * sizeof(T));
        if (!base || off + bytes > size) { ok = false; off += bytes; return nullptr; }
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
};

// Grid size for a grid-stride kernel: enough CTAs to fill the 148 SMs a few
// times over, never more than the work needs.
static inline int grid_for(int64_t n, int block, int ctas_per_sm = 8) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)num_sms() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace upcp
