// S5 -- facade region growing as frontier expansion over the kNN graph, plus the
// compact <-> full index-space mask helpers.
//
// Restates RegionGrowing._region_growing(method='knn') (region_growing.py:75-141).
// Acceptance of a neighbour depends only on the two normals and on membership of the
// seed's kNN row; seed-hood of an accepted point depends only on its own neighbourhood
// (`d_can_seed`); `processed` is set only on acceptance.  The result is therefore the
// least fixpoint of
//     region  = S0  u  { p : exists seed s, p in knn(s)[1:], angle(n_s, n_p) < thr }
//     seeds   = S0  u  { p in region : can_seed[p] }
// and is independent of the expansion order, so the Python `while` loop over a growing
// list maps onto an asynchronous device-wide work queue:
//   * the initial seeds S0 (usually most of the cloud: every point the BGT footprint
//     labelled) are expanded data-parallel, one thread per (seed, neighbour) pair, in index
//     order -- coalesced, no queue traffic; the points they accept go to the work queue;
//   * one persistent kernel, grid sized to be fully co-resident, drains the queue: a warp
//     takes 32 tickets from an atomic head counter, its lanes test the k-1 neighbours,
//     claim accepted points with an atomic exchange and push new seeds (atomic tail
//     counter).  `pending` counts pushed-but-unfinished seeds; a waiting warp leaves when
//     it is 0.  No launch per BFS level, no host round trip until the fixpoint is reached;
//   * begin / resume / end: the caller may start growing with some seed flags still
//     undecided (value 2: the curvature test that needs LAPACK's eigenvalue order), resolve
//     them on the host while the device grows, and resume from the points that turned out to
//     be seeds -- growth is monotone, so the fixpoint is the same.
#include <math.h>
#include <stdlib.h>
#include "common.cuh"
#include "primitives.cuh"
#include "upcp_math.cuh"

namespace upcp {

constexpr int kBlock = 256;
constexpr int kGrowThreads = 256;

struct GrowCounters {      // device-resident, 64-bit aligned
    unsigned int head;
    unsigned int tail;
    int pending;
    unsigned int error;
    unsigned long long region_size;
};

__global__ void __launch_bounds__(kBlock)
grow_init_kernel(const uint8_t* __restrict__ seed, int64_t m, int32_t* __restrict__ state) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < m; i += stride) state[i] = seed[i] != 0 ? 1 : 0;
}

// warp-aggregated push of accepted points that are seeds themselves
__device__ __forceinline__ void grow_push(bool push, int32_t nb, volatile int32_t* vqueue, GrowCounters* ctr) {
    const int lane = threadIdx.x & 31;
    const uint32_t bal = __ballot_sync(0xffffffffu, push);
    if (bal) {
        const int leader = __ffs(bal) - 1;
        unsigned int base = 0;
        if (lane == leader) {
            atomicAdd(&ctr->pending, __popc(bal));           // before the slots become visible
            base = atomicAdd(&ctr->tail, (unsigned int)__popc(bal));
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        if (push) vqueue[base + __popc(bal & ((1u << lane) - 1u))] = nb;
    }
}

// Expansion of the initial seeds: PER threads per point (one per kNN slot), index order.
template <int PER>
__global__ void __launch_bounds__(kBlock)
grow_seed_pass_kernel(const int32_t* __restrict__ knn, int k, const double* __restrict__ normals,
                      const uint8_t* __restrict__ seed, const uint8_t* __restrict__ can_seed, int64_t m,
                      double threshold_deg, int32_t* __restrict__ state, int32_t* __restrict__ queue,
                      GrowCounters* __restrict__ ctr) {
    const int64_t total = (m * PER + 31) & ~(int64_t)31;          // whole warps
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x; t < total; t += stride) {
        const int64_t s = t / PER;
        const int j = (int)(t % PER);
        bool push = false;
        int32_t nb = -1;
        // neighbour_idx[1:k] (region_growing.py:112): the first returned index is dropped
        if (s < m && j >= 1 && j < k && seed[s] != 0) {
            nb = knn[(size_t)s * k + j];
            if (nb >= 0 && __ldcg(state + nb) == 0) {
                const Vec3 ns = {normals[(size_t)s * 3], normals[(size_t)s * 3 + 1], normals[(size_t)s * 3 + 2]};
                const Vec3 nn = {normals[(size_t)nb * 3], normals[(size_t)nb * 3 + 1], normals[(size_t)nb * 3 + 2]};
                if (vector_angle_deg(ns, nn) < threshold_deg) {
                    if (atomicExch(&state[nb], 1) == 0) push = can_seed[nb] == 1;
                }
            }
        }
        grow_push(push, nb, queue, ctr);
    }
}

// resume: the queue restarts at its current tail ...
__global__ void grow_rewind_kernel(GrowCounters* __restrict__ ctr) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { ctr->head = ctr->tail; ctr->pending = 0; }
}
// ... with the points of `extra` that are in the region, were not seeds so far and are seeds now
__global__ void __launch_bounds__(kBlock)
grow_requeue_kernel(const int64_t* __restrict__ extra, int64_t n_extra, const uint8_t* __restrict__ seed,
                    const uint8_t* __restrict__ can_seed, const int32_t* __restrict__ state,
                    int32_t* __restrict__ queue, GrowCounters* __restrict__ ctr) {
    const int64_t total = (n_extra + 31) & ~(int64_t)31;
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < total; i += stride) {
        bool push = false;
        int32_t p = -1;
        if (i < n_extra) {
            p = (int32_t)extra[i];
            push = state[p] != 0 && seed[p] == 0 && can_seed[p] == 1;
        }
        grow_push(push, p, queue, ctr);
    }
}

// Seeds a warp discovers stay with that warp (a small ring in shared memory) up to the tuning knob grow_local_chain (measured optimum 2; 1 .. 32 swept); only
// the overflow goes through the global queue.  The frontier is narrow and every hop through the
// global queue costs three dependent L2 round trips (tail / pending atomics, slot write, the
// consumer's poll) on top of the expansion itself; the local path removes them from the chain.

constexpr int kRing = 64;                // ring size (power of two, >= the largest local cap + 32)

__global__ void __launch_bounds__(kGrowThreads)
grow_kernel(const int32_t* __restrict__ knn, int k, const double* __restrict__ normals,
            const uint8_t* __restrict__ can_seed, int64_t m, double threshold_deg,
            int32_t* __restrict__ state, int32_t* __restrict__ queue, GrowCounters* __restrict__ ctr,
            int local_cap, int sleep_ns) {
    __shared__ int32_t s_ring[kGrowThreads / 32][kRing];
    const int lane = threadIdx.x & 31;
    int32_t* const ring = s_ring[threadIdx.x >> 5];
    unsigned int r_head = 0, r_tail = 0;                       // warp-uniform
    volatile int32_t* vqueue = queue;
    volatile int* vpending = &ctr->pending;
    // two seeds per step when a kNN row fits a half-warp, else one
    const int per_step = k <= 16 ? 2 : 1;
    const int sub = per_step == 2 ? (lane >> 4) : 0;          // which seed of the step this lane serves
    const int j = per_step == 2 ? (lane & 15) : lane;        // neighbour slot within the row

    // expand the seeds held by the lanes of `batch` (lane l holds seed `mine`)
    auto process = [&](int32_t mine, uint32_t batch) {
        uint32_t todo = batch;
        while (todo) {
            // pick the seed(s) of this step
            const int srcA = __ffs(todo) - 1;
            todo &= todo - 1;
            int srcB = -1;
            if (per_step == 2 && todo) { srcB = __ffs(todo) - 1; todo &= todo - 1; }
            const int src = sub == 0 ? srcA : srcB;
            const int32_t s = __shfl_sync(0xffffffffu, mine, src < 0 ? 0 : src);
            bool push = false;
            int32_t nb = -1;
            // neighbour_idx[1:k] (region_growing.py:112): the first returned index is dropped
            for (int jj = j; jj < k; jj += (per_step == 2 ? 16 : 32)) {
                // (k <= 32: at most one trip)
                if (src >= 0 && jj >= 1) {
                    // issue every load that only needs the indices at once (the seed's normal with
                    // the row, the neighbour's state / normal / flag together)
                    const Vec3 ns = {normals[(size_t)s * 3], normals[(size_t)s * 3 + 1], normals[(size_t)s * 3 + 2]};
                    nb = knn[(size_t)s * k + jj];
                    if (nb >= 0) {
                        const int32_t st = __ldcg(state + nb);
                        const Vec3 nn = {normals[(size_t)nb * 3], normals[(size_t)nb * 3 + 1], normals[(size_t)nb * 3 + 2]};
                        const uint8_t cs = can_seed[nb];
                        if (st == 0 && vector_angle_deg(ns, nn) < threshold_deg) {
                            if (atomicExch(&state[nb], 1) == 0) push = cs == 1;
                        }
                    }
                }
            }
            // new seeds: the first ones stay in this warp's ring, the rest go to the global queue
            const uint32_t bal = __ballot_sync(0xffffffffu, push);
            if (bal) {
                const int cnt = __popc(bal);
                const int room = local_cap - (int)(r_tail - r_head);
                const int n_loc = room > 0 ? (cnt < room ? cnt : room) : 0;
                const int my = __popc(bal & ((1u << lane) - 1u));
                const int leader = __ffs(bal) - 1;
                unsigned int base = 0;
                if (lane == leader) {
                    atomicAdd(&ctr->pending, cnt);               // before any of them becomes visible
                    if (cnt > n_loc) base = atomicAdd(&ctr->tail, (unsigned int)(cnt - n_loc));
                }
                base = __shfl_sync(0xffffffffu, base, leader);
                if (push) {
                    if (my < n_loc) ring[(r_tail + my) & (kRing - 1)] = nb;
                    else vqueue[base + (my - n_loc)] = nb;
                }
                r_tail += n_loc;
                __syncwarp();
            }
        }
    };
    // drain the local ring
    auto drain_local = [&]() {
        while (r_tail != r_head) {
            const int cnt = min((int)(r_tail - r_head), 32);
            const int32_t mine = lane < cnt ? ring[(r_head + lane) & (kRing - 1)] : -1;
            __syncwarp();
            r_head += cnt;
            process(mine, cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u));
            __syncwarp();
            if (lane == 0) { __threadfence(); atomicSub(&ctr->pending, cnt); }
        }
    };

    while (true) {
        // a warp takes 32 tickets at once (one atomic on the shared head counter per 32 seeds) ...
        unsigned int t0 = 0;
        if (lane == 0) t0 = atomicAdd(&ctr->head, 32u);
        t0 = __shfl_sync(0xffffffffu, t0, 0);
        if ((int64_t)t0 >= m) return;                 // at most m seeds can ever be queued
        // ... and serves them as they are published: waiting for all 32 first could deadlock,
        // because the seeds that fill the later tickets may come from the earlier ones
        int32_t mine = -1;
        uint32_t remaining = __ballot_sync(0xffffffffu, (int64_t)t0 + lane < m);
        unsigned int spins = 0;
        while (remaining) {
            if (((remaining >> lane) & 1u) && mine < 0) mine = vqueue[t0 + lane];
            const uint32_t ready = __ballot_sync(0xffffffffu, ((remaining >> lane) & 1u) && mine >= 0);
            if (ready == 0) {
                int pend = 1;
                if (lane == 0) pend = *vpending;
                pend = __shfl_sync(0xffffffffu, pend, 0);
                if (pend <= 0) return;                 // global quiescence: nothing left anywhere
                __nanosleep(sleep_ns);
                if (++spins > (1u << 26)) { if (lane == 0) atomicExch(&ctr->error, 1u); return; }   // safety valve
                continue;
            }
            process(mine, ready);
            __syncwarp();
            if (lane == 0) { __threadfence(); atomicSub(&ctr->pending, __popc(ready)); }
            remaining &= ~ready;
            spins = 0;
            drain_local();
        }
    }
}

__global__ void __launch_bounds__(kBlock)
grow_finish_kernel(const int32_t* __restrict__ state, int64_t m, uint8_t* __restrict__ region,
                   GrowCounters* __restrict__ ctr) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    unsigned long long cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < m; i += stride) {
        uint8_t r = state[i] != 0;
        region[i] = r;
        cnt += r;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&ctr->region_size, cnt);
}

__global__ void grow_info_kernel(const GrowCounters* __restrict__ ctr, long long* __restrict__ info) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        info[0] = (long long)ctr->region_size;
        info[1] = (long long)ctr->tail;
    }
}

// ------------------------------------------------------ compact <-> full masks --
struct ScatterFunctor {
    const uint8_t* compact; uint8_t* out;
    __device__ void operator()(int64_t i, bool selected, uint32_t rank) const {
        out[i] = selected ? compact[rank] : (uint8_t)0;
    }
};
struct GatherFunctor {
    const uint8_t* full; uint8_t* compact;
    __device__ void operator()(int64_t i, bool selected, uint32_t rank) const {
        if (selected) compact[rank] = full[i];
    }
};
// sorted-space <-> full-tile masks through the order array of upcp_knn_normals_sorted
__global__ void __launch_bounds__(kBlock)
mask_take_kernel(const uint8_t* __restrict__ full, const uint32_t* __restrict__ order, int64_t m,
                 uint8_t* __restrict__ out) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t s = (int64_t)blockIdx.x * kBlock + threadIdx.x; s < m; s += stride) out[s] = full[order[s]];
}
__global__ void __launch_bounds__(kBlock)
mask_put_kernel(const uint8_t* __restrict__ sorted, const uint32_t* __restrict__ order, int64_t m,
                uint8_t* __restrict__ full) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t s = (int64_t)blockIdx.x * kBlock + threadIdx.x; s < m; s += stride)
        if (sorted[s]) full[order[s]] = 1;
}
__global__ void widen_count_kernel(const uint32_t* __restrict__ c, long long* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = (long long)*c;
}

}  // namespace upcp

using namespace upcp;

extern "C" {

size_t upcp_region_grow_ws_bytes(int64_t m) {
    if (m < 1) m = 1;
    return align_up(256) + 2 * align_up((size_t)m * 4) + 4096;
}

namespace {
struct GrowWs { GrowCounters* ctr; int32_t* state; int32_t* queue; };
int grow_ws(void* d_ws, size_t ws_bytes, int64_t m, GrowWs* out) {
    Workspace ws(d_ws, ws_bytes);
    out->ctr = (GrowCounters*)ws.take<char>(256);
    out->state = ws.take<int32_t>((size_t)m);
    out->queue = ws.take<int32_t>((size_t)m);
    if (!ws.ok) UPCP_FAIL(UPCP_E_WORKSPACE, "region_grow: workspace too small");
    return UPCP_OK;
}
int grow_drain(const int32_t* d_knn, int knn, const double* d_normals, const uint8_t* d_can_seed, int64_t m,
               double threshold_angle_deg, const GrowWs& w, cudaStream_t st) {
    // persistent kernel: every CTA must be resident (waiting warps spin on the queue).  One CTA per
    // SM by default: more pollers only add contention on the queue counters (1 / 2 / 4 swept).
    const upcp_tuning tune = tuning();
    int per_sm = 0;
    UPCP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, grow_kernel, kGrowThreads, 0));
    if (per_sm < 1) UPCP_FAIL(UPCP_E_CUDA, "region_grow: kernel cannot be resident");
    if (per_sm > tune.grow_ctas_per_sm) per_sm = tune.grow_ctas_per_sm;
    int grid = num_sms() * per_sm;
    int64_t need = (m + (kGrowThreads / 32) - 1) / (kGrowThreads / 32);
    if (need < grid) grid = (int)(need < 1 ? 1 : need);
    const int local_cap = tune.grow_local_chain;
    const int sleep_ns = 64;
    grow_kernel<<<grid, kGrowThreads, 0, st>>>(d_knn, knn, d_normals, d_can_seed, m, threshold_angle_deg,
                                               w.state, w.queue, w.ctr, local_cap, sleep_ns);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}
}  // namespace

int upcp_region_grow_begin(const int32_t* d_knn, int knn, const double* d_normals,
                           const uint8_t* d_seed, const uint8_t* d_can_seed, int64_t m,
                           double threshold_angle_deg, void* d_ws, size_t ws_bytes, void* stream) {
    if (m < 0 || knn < 1 || knn > 32) UPCP_FAIL(UPCP_E_INVALID, "region_grow: bad argument");
    if (m >= ((int64_t)1 << 31)) UPCP_FAIL(UPCP_E_CAPACITY, "region_grow: m must be < 2^31");
    if (m == 0) return UPCP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    GrowWs w;
    UPCP_TRY(grow_ws(d_ws, ws_bytes, m, &w));
    UPCP_CUDA_OK(cudaMemsetAsync(w.ctr, 0, 256, st));
    UPCP_CUDA_OK(cudaMemsetAsync(w.queue, 0xff, (size_t)m * 4, st));
    grow_init_kernel<<<grid_for(m, kBlock), kBlock, 0, st>>>(d_seed, m, w.state);
    UPCP_LAUNCH_OK();
    prof_begin(UPCP_PROF_GROW, st);
    if (knn <= 16)
        grow_seed_pass_kernel<16><<<grid_for(m * 16, kBlock, 16), kBlock, 0, st>>>(
            d_knn, knn, d_normals, d_seed, d_can_seed, m, threshold_angle_deg, w.state, w.queue, w.ctr);
    else
        grow_seed_pass_kernel<32><<<grid_for(m * 32, kBlock, 16), kBlock, 0, st>>>(
            d_knn, knn, d_normals, d_seed, d_can_seed, m, threshold_angle_deg, w.state, w.queue, w.ctr);
    UPCP_LAUNCH_OK();
    int rc = grow_drain(d_knn, knn, d_normals, d_can_seed, m, threshold_angle_deg, w, st);
    prof_end(UPCP_PROF_GROW, st, m);
    return rc;
}

int upcp_region_grow_resume(const int32_t* d_knn, int knn, const double* d_normals,
                            const uint8_t* d_seed, const uint8_t* d_can_seed, int64_t m,
                            const int64_t* d_extra, int64_t n_extra,
                            double threshold_angle_deg, void* d_ws, size_t ws_bytes, void* stream) {
    if (m < 0 || knn < 1 || knn > 32 || n_extra < 0) UPCP_FAIL(UPCP_E_INVALID, "region_grow: bad argument");
    if (m == 0 || n_extra == 0) return UPCP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    GrowWs w;
    UPCP_TRY(grow_ws(d_ws, ws_bytes, m, &w));
    grow_rewind_kernel<<<1, 32, 0, st>>>(w.ctr);
    UPCP_LAUNCH_OK();
    grow_requeue_kernel<<<grid_for(n_extra, kBlock), kBlock, 0, st>>>(d_extra, n_extra, d_seed, d_can_seed, w.state,
                                                                      w.queue, w.ctr);
    UPCP_LAUNCH_OK();
    return grow_drain(d_knn, knn, d_normals, d_can_seed, m, threshold_angle_deg, w, st);
}

int upcp_region_grow_end(int64_t m, uint8_t* d_region, int64_t* d_info, void* d_ws, size_t ws_bytes, void* stream) {
    if (m < 0) UPCP_FAIL(UPCP_E_INVALID, "region_grow: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (d_info) UPCP_CUDA_OK(cudaMemsetAsync(d_info, 0, 2 * sizeof(int64_t), st));
    if (m == 0) return UPCP_OK;
    GrowWs w;
    UPCP_TRY(grow_ws(d_ws, ws_bytes, m, &w));
    grow_finish_kernel<<<grid_for(m, kBlock), kBlock, 0, st>>>(w.state, m, d_region, w.ctr);
    UPCP_LAUNCH_OK();
    if (d_info) {
        grow_info_kernel<<<1, 32, 0, st>>>(w.ctr, (long long*)d_info);
        UPCP_LAUNCH_OK();
    }
    // surface the safety valve (a spin that never resolved) as an error
    unsigned int h_err = 0;
    UPCP_CUDA_OK(cudaMemcpyAsync(&h_err, &w.ctr->error, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    UPCP_CUDA_OK(cudaStreamSynchronize(st));
    if (h_err) UPCP_FAIL(UPCP_E_INTERNAL, "region_grow: work queue did not quiesce");
    return UPCP_OK;
}

int upcp_region_grow(const int32_t* d_knn, int knn, const double* d_normals,
                     const uint8_t* d_seed, const uint8_t* d_can_seed, int64_t m,
                     double threshold_angle_deg, uint8_t* d_region, int64_t* d_info,
                     void* d_ws, size_t ws_bytes, void* stream) {
    UPCP_TRY(upcp_region_grow_begin(d_knn, knn, d_normals, d_seed, d_can_seed, m, threshold_angle_deg, d_ws, ws_bytes,
                                    stream));
    return upcp_region_grow_end(m, d_region, d_info, d_ws, ws_bytes, stream);
}

size_t upcp_compact_ws_bytes(int64_t n) { return (compact_ws_words(n < 1 ? 1 : n) + 64) * 4; }

int upcp_mask_scatter(const uint8_t* d_sel, int64_t n, const uint8_t* d_compact, uint8_t* d_out,
                      void* d_ws, size_t ws_bytes, void* stream) {
    if (n < 0 || !d_sel) UPCP_FAIL(UPCP_E_INVALID, "mask_scatter: bad argument");
    if (n == 0) return UPCP_OK;
    if (!d_ws || ws_bytes < upcp_compact_ws_bytes(n)) UPCP_FAIL(UPCP_E_WORKSPACE, "mask_scatter: workspace too small");
    uint32_t* w = (uint32_t*)d_ws;
    ScatterFunctor f{d_compact, d_out};
    return compact_apply(d_sel, n, w, w + 64, (cudaStream_t)stream, f);
}

int upcp_mask_take(const uint8_t* d_full, const uint32_t* d_order, int64_t m, uint8_t* d_out, void* stream) {
    if (m < 0 || (m > 0 && (!d_full || !d_order || !d_out))) UPCP_FAIL(UPCP_E_INVALID, "mask_take: bad argument");
    if (m == 0) return UPCP_OK;
    mask_take_kernel<<<grid_for(m, kBlock), kBlock, 0, (cudaStream_t)stream>>>(d_full, d_order, m, d_out);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

int upcp_mask_put(const uint8_t* d_sorted, const uint32_t* d_order, int64_t m, int64_t n, uint8_t* d_full,
                  void* stream) {
    if (m < 0 || n < 0 || m > n || (n > 0 && !d_full)) UPCP_FAIL(UPCP_E_INVALID, "mask_put: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (n > 0) UPCP_CUDA_OK(cudaMemsetAsync(d_full, 0, (size_t)n, st));
    if (m == 0) return UPCP_OK;
    if (!d_sorted || !d_order) UPCP_FAIL(UPCP_E_INVALID, "mask_put: bad argument");
    mask_put_kernel<<<grid_for(m, kBlock), kBlock, 0, st>>>(d_sorted, d_order, m, d_full);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

int upcp_mask_gather(const uint8_t* d_sel, int64_t n, const uint8_t* d_full, uint8_t* d_compact,
                     int64_t* d_count, void* d_ws, size_t ws_bytes, void* stream) {
    if (n < 0 || !d_sel) UPCP_FAIL(UPCP_E_INVALID, "mask_gather: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        if (d_count) UPCP_CUDA_OK(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
        return UPCP_OK;
    }
    if (!d_ws || ws_bytes < upcp_compact_ws_bytes(n)) UPCP_FAIL(UPCP_E_WORKSPACE, "mask_gather: workspace too small");
    uint32_t* w = (uint32_t*)d_ws;
    GatherFunctor f{d_full, d_compact};
    UPCP_TRY(compact_apply(d_sel, n, w, w + 64, st, f));
    if (d_count) {
        widen_count_kernel<<<1, 32, 0, st>>>(w, (long long*)d_count);
        UPCP_LAUNCH_OK();
    }
    return UPCP_OK;
}

}  // extern "C"
