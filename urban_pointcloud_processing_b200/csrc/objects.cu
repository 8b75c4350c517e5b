// Per-object reductions for the clients of LabelConnectedComp.get_components
// (SURVEY 8f row 2): CarFuser (reference fusion/car_fuser.py:45-89) and
// BGTStreetFurnitureFuser (fusion/street_furniture_fuser.py:46-89) loop over the clusters
// in Python and, for each one, rescan the whole masked cloud (`point_components == cc`,
// `ground_z[cc_mask]`, `np.amax(points[cc_mask][:, 2])`).  Here one pass groups the points
// by cluster (radix sort of (cluster id, point index)) and reduces every cluster at once:
//
//   component_group_kernel   flags the points that are in a cluster
//   (compaction + radix sort + boundary compaction from primitives.cuh)
//   component_stats_kernel   per cluster: #finite ground values, their sum, max z, xy bbox
//   component_decode_kernel  order-preserving integer keys -> fp64
//   gather_xy_kernel         the xy of one cluster, in original point order (qhull input)
//   component_mark_kernel    mask of the points whose cluster was accepted
//   prism_clip_kernel        clip_utils.poly_box_clip for a list of (ring, bottom, top)
//
// Determinism: the only floating-point reduction, the sum of the ground elevations, is done
// in 2^-24 m fixed point with integer atomics (associative, so independent of the order in
// which warps arrive).  AHN rasters are float16 (ahn_utils.py:136-137): every value is a
// multiple of 2^-24, the fixed-point sum is exact, and so is numpy's pairwise sum for any
// realistic cluster, hence the same mean.  Other readers' values are rounded to 6e-8 m.
#include "common.cuh"
#include "primitives.cuh"
#include "upcp_math.cuh"

namespace upcp {

namespace {
constexpr int kBlock = 256;
constexpr double kFixedScale = 16777216.0;           // 2^24
constexpr double kFixedLimit = 1048576.0;            // |gz| < 2^20 m

// order-preserving map fp64 -> u64 (NaN never passed in)
__device__ __forceinline__ unsigned long long ord_key(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_val(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(kBlock)
component_group_kernel(const int32_t* __restrict__ comp, int64_t n, uint8_t* __restrict__ flags) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) flags[i] = comp[i] >= 0;
}

struct PairFunctor {
    const int32_t* comp;
    uint32_t* keys;
    uint32_t* vals;
    __device__ void operator()(int64_t i, bool selected, uint32_t rank) const {
        if (selected) { keys[rank] = (uint32_t)comp[i]; vals[rank] = (uint32_t)i; }
    }
};

__global__ void __launch_bounds__(kBlock)
component_boundary_kernel(const uint32_t* __restrict__ keys, int64_t m, uint8_t* __restrict__ flags) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t s = (int64_t)blockIdx.x * kBlock + threadIdx.x; s < m; s += stride)
        flags[s] = (s == 0 || keys[s] != keys[s - 1]) ? 1 : 0;
}

struct SegmentFunctor {
    uint32_t* dense_of;            // [m] dense cluster id of sorted position s
    long long* seg_first;          // [cap + 1]
    int32_t cap;
    __device__ void operator()(int64_t s, bool selected, uint32_t rank) const {
        const uint32_t k = selected ? rank : rank - 1;
        dense_of[s] = k;
        if (selected && k < (uint32_t)cap) seg_first[k] = (long long)s;
    }
};

struct StatsOut {
    long long* count_finite;       // [cap]
    long long* sum_fixed;          // [cap]
    unsigned long long* max_z;     // [cap] ordered keys, decoded in place afterwards
    unsigned long long* bbox;      // [cap][4] ordered keys (min x, min y, max x, max y)
};

__global__ void component_init_kernel(StatsOut o, int32_t cap, long long* seg_first, const uint32_t* n_seg,
                                      long long m, long long* info) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < cap) {
        o.count_finite[k] = 0; o.sum_fixed[k] = 0; o.max_z[k] = 0ull;
        o.bbox[4 * k + 0] = ~0ull; o.bbox[4 * k + 1] = ~0ull; o.bbox[4 * k + 2] = 0ull; o.bbox[4 * k + 3] = 0ull;
    }
    if (k == 0) {
        const uint32_t K = *n_seg;
        info[0] = (long long)K; info[1] = m;
        if (K <= (uint32_t)cap) seg_first[K] = m;
    }
}

__global__ void __launch_bounds__(kBlock)
component_stats_kernel(const uint32_t* __restrict__ vals, const uint32_t* __restrict__ dense_of, int64_t m,
                       const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                       const double* __restrict__ gz, int32_t cap, int32_t* __restrict__ dense,
                       uint32_t* __restrict__ order, StatsOut o, long long* __restrict__ info) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    const int lane = threadIdx.x & 31;
    for (int64_t s0 = (int64_t)blockIdx.x * kBlock; s0 < m; s0 += stride) {
        const int64_t s = s0 + threadIdx.x;
        const bool valid = s < m;
        uint32_t k = 0xffffffffu;
        long long fin = 0, fx = 0;
        unsigned long long kz = 0ull, kx0 = ~0ull, ky0 = ~0ull, kx1 = 0ull, ky1 = 0ull;
        if (valid) {
            const uint32_t i = vals[s];
            k = dense_of[s];
            dense[i] = (int32_t)k;
            order[s] = i;
            const double px = x[i], py = y[i], pz = z ? z[i] : 0.0;
            if (!isnan(pz)) kz = ord_key(pz);
            if (!isnan(px)) { kx0 = ord_key(px); kx1 = kx0; }
            if (!isnan(py)) { ky0 = ord_key(py); ky1 = ky0; }
            if (gz) {
                const double g = gz[i];
                if (isfinite(g)) {                                   // car_fuser.py:62 np.isfinite
                    if (fabs(g) < kFixedLimit) { fin = 1; fx = __double2ll_rn(g * kFixedScale); }
                    else atomicAdd((unsigned long long*)&info[3], 1ull);
                }
            }
        }
        // sorted by cluster: a warp nearly always sits inside one cluster -> one atomic set per warp
        const uint32_t k0 = __shfl_sync(0xffffffffu, k, 0);
        if (__all_sync(0xffffffffu, k == k0)) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                fin += __shfl_xor_sync(0xffffffffu, fin, off);
                fx += __shfl_xor_sync(0xffffffffu, fx, off);
                unsigned long long t;
                t = __shfl_xor_sync(0xffffffffu, kz, off); kz = t > kz ? t : kz;
                t = __shfl_xor_sync(0xffffffffu, kx0, off); kx0 = t < kx0 ? t : kx0;
                t = __shfl_xor_sync(0xffffffffu, ky0, off); ky0 = t < ky0 ? t : ky0;
                t = __shfl_xor_sync(0xffffffffu, kx1, off); kx1 = t > kx1 ? t : kx1;
                t = __shfl_xor_sync(0xffffffffu, ky1, off); ky1 = t > ky1 ? t : ky1;
            }
            if (lane != 0) k = 0xffffffffu;                          // lane 0 carries the warp's totals
        }
        if (k < (uint32_t)cap) {
            if (fin) {
                atomicAdd((unsigned long long*)&o.count_finite[k], (unsigned long long)fin);
                atomicAdd((unsigned long long*)&o.sum_fixed[k], (unsigned long long)fx);
            }
            atomicMax(&o.max_z[k], kz);
            atomicMin(&o.bbox[4 * k + 0], kx0); atomicMin(&o.bbox[4 * k + 1], ky0);
            atomicMax(&o.bbox[4 * k + 2], kx1); atomicMax(&o.bbox[4 * k + 3], ky1);
        }
    }
}

__global__ void component_decode_kernel(StatsOut o, int32_t cap, const uint32_t* n_seg) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t K = *n_seg;
    if (k < cap && (uint32_t)k < K) {
        double* mz = (double*)o.max_z;
        double* bb = (double*)o.bbox;
        const unsigned long long kz = o.max_z[k];
        mz[k] = kz ? ord_val(kz) : __longlong_as_double(0x7ff8000000000000LL);
        for (int c = 0; c < 4; ++c) {
            const unsigned long long kk = o.bbox[4 * k + c];
            bb[4 * k + c] = (kk == 0ull || kk == ~0ull) ? __longlong_as_double(0x7ff8000000000000LL) : ord_val(kk);
        }
    }
}

__global__ void __launch_bounds__(kBlock)
gather_xy_kernel(const double* __restrict__ x, const double* __restrict__ y, const uint32_t* __restrict__ idx,
                 int64_t cnt, double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t s = (int64_t)blockIdx.x * kBlock + threadIdx.x; s < cnt; s += stride) {
        const uint32_t i = idx[s];
        out[2 * s] = x[i];
        out[2 * s + 1] = y[i];
    }
}

__global__ void __launch_bounds__(kBlock)
component_mark_kernel(const int32_t* __restrict__ dense, int64_t n, const uint8_t* __restrict__ accepted,
                      int32_t n_comp, uint8_t* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        const int32_t k = dense[i];
        out[i] = (k >= 0 && k < n_comp && accepted[k]) ? 1 : 0;
    }
}

// clip_utils.poly_box_clip (clip_utils.py:240-262) for P prisms at once: inclusive bounding-box
// prefilter (rectangle_clip, :22-40, on compute_bounding_box of the ring), the literal crossing-number
// test over the ring in vertex order (boundary counts as inside), then bottom <= z <= top.
constexpr int kPrismChunk = 512;

__global__ void __launch_bounds__(kBlock)
prism_clip_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                  int64_t n, const uint8_t* __restrict__ sel, const double* __restrict__ ring_xy,
                  const int32_t* __restrict__ ring_off, const double* __restrict__ zrange, int p0, int p1,
                  uint8_t* __restrict__ out) {
    __shared__ double s_box[kPrismChunk][6];             // x_min, y_min, x_max, y_max, bottom, top
    __shared__ int s_off[kPrismChunk + 1];
    const int np = p1 - p0;
    for (int p = threadIdx.x; p < np; p += kBlock) {
        const int a = ring_off[p0 + p], b = ring_off[p0 + p + 1];
        // np.min / np.max of the ring columns (math_utils.py:48-55)
        double x0 = ring_xy[2 * a], y0 = ring_xy[2 * a + 1], x1 = x0, y1 = y0;
        for (int v = a + 1; v < b; ++v) {
            const double vx = ring_xy[2 * v], vy = ring_xy[2 * v + 1];
            x0 = fmin(x0, vx); x1 = fmax(x1, vx); y0 = fmin(y0, vy); y1 = fmax(y1, vy);
        }
        s_box[p][0] = x0; s_box[p][1] = y0; s_box[p][2] = x1; s_box[p][3] = y1;
        s_box[p][4] = zrange[2 * (p0 + p)]; s_box[p][5] = zrange[2 * (p0 + p) + 1];
        s_off[p] = a;
        if (p == np - 1) s_off[np] = b;
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        if (out[i] || (sel && !sel[i])) continue;
        const double px = x[i], py = y[i], pz = z[i];
        bool in = false;
        for (int p = 0; p < np && !in; ++p) {
            if (px >= s_box[p][0] && px <= s_box[p][2] && py >= s_box[p][1] && py <= s_box[p][3] &&
                pz <= s_box[p][5] && pz >= s_box[p][4]) {
                const int a = s_off[p], b = s_off[p + 1];
                in = pip_ring(ring_xy + 2 * a, b - a, px, py) != 0;
            }
        }
        if (in) out[i] = 1;
    }
}

int bits_for(int64_t n) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n) ++b;
    return b;
}

}  // namespace

}  // namespace upcp

using namespace upcp;

extern "C" {

size_t upcp_component_stats_ws_bytes(int64_t n) {
    if (n < 1) n = 1;
    size_t bytes = 0;
    bytes += align_up(64 * 4);                                                          // counters
    bytes += align_up((size_t)n);                                                       // flags
    bytes += align_up((compact_ws_words(n) + sort_ws_words(n)) * 4);                    // primitives
    bytes += 4 * align_up((size_t)n * 4);                                               // keys / vals a, b
    bytes += align_up((size_t)n * 4);                                                   // dense_of
    return bytes + 4096;
}

int upcp_component_stats(const int32_t* d_comp, int64_t n, const double* d_x, const double* d_y,
                         const double* d_z, const double* d_gz, int32_t cap, int32_t* d_dense,
                         uint32_t* d_order, int64_t* d_seg_first, int64_t* d_count_finite,
                         int64_t* d_sum_fixed, double* d_max_z, double* d_bbox_xy, int64_t* d_info,
                         void* d_ws, size_t ws_bytes, void* stream) {
    if (n < 0 || cap < 1 || !d_comp || !d_x || !d_y || !d_dense || !d_order || !d_seg_first || !d_count_finite ||
        !d_sum_fixed || !d_max_z || !d_bbox_xy || !d_info)
        UPCP_FAIL(UPCP_E_INVALID, "component_stats: bad argument");
    if (n >= ((int64_t)1 << 31)) UPCP_FAIL(UPCP_E_CAPACITY, "component_stats: n must be < 2^31");
    cudaStream_t st = (cudaStream_t)stream;
    UPCP_CUDA_OK(cudaMemsetAsync(d_info, 0, 4 * sizeof(int64_t), st));
    if (n == 0) return UPCP_OK;
    Workspace ws(d_ws, ws_bytes);
    uint32_t* counters = ws.take<uint32_t>(64);           // [0] m, [1] K
    uint8_t* flags = ws.take<uint8_t>((size_t)n);
    uint32_t* prim_ws = ws.take<uint32_t>(compact_ws_words(n) + sort_ws_words(n));
    uint32_t* keys_a = ws.take<uint32_t>((size_t)n);
    uint32_t* keys_b = ws.take<uint32_t>((size_t)n);
    uint32_t* vals_a = ws.take<uint32_t>((size_t)n);
    uint32_t* vals_b = ws.take<uint32_t>((size_t)n);
    uint32_t* dense_of = ws.take<uint32_t>((size_t)n);
    if (!ws.ok) UPCP_FAIL(UPCP_E_WORKSPACE, "component_stats: workspace too small");

    UPCP_CUDA_OK(cudaMemsetAsync(d_dense, 0xff, (size_t)n * sizeof(int32_t), st));
    const int grid_n = grid_for(n, kBlock);
    component_group_kernel<<<grid_n, kBlock, 0, st>>>(d_comp, n, flags);
    UPCP_LAUNCH_OK();
    {
        PairFunctor f{d_comp, keys_a, vals_a};
        UPCP_TRY(compact_apply(flags, n, counters, prim_ws, st, f));
    }
    uint32_t hm = 0;
    UPCP_CUDA_OK(cudaMemcpyAsync(&hm, counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    UPCP_CUDA_OK(cudaStreamSynchronize(st));
    const int64_t m = hm;
    StatsOut o{(long long*)d_count_finite, (long long*)d_sum_fixed, (unsigned long long*)d_max_z,
               (unsigned long long*)d_bbox_xy};
    if (m == 0) {
        UPCP_CUDA_OK(cudaMemsetAsync(counters + 1, 0, sizeof(uint32_t), st));
        component_init_kernel<<<(cap + kBlock - 1) / kBlock, kBlock, 0, st>>>(o, cap, (long long*)d_seg_first,
                                                                              counters + 1, 0, (long long*)d_info);
        UPCP_LAUNCH_OK();
        return UPCP_OK;
    }
    // stable LSD sort on the cluster id: within a cluster the point indices stay ascending
    bool in_b = false;
    UPCP_TRY(radix_sort_pairs_u32(keys_a, vals_a, keys_b, vals_b, m, bits_for(n), prim_ws, st, &in_b));
    const uint32_t* keys = in_b ? keys_b : keys_a;
    const uint32_t* vals = in_b ? vals_b : vals_a;
    const int grid_m = grid_for(m, kBlock);
    component_boundary_kernel<<<grid_m, kBlock, 0, st>>>(keys, m, flags);
    UPCP_LAUNCH_OK();
    {
        SegmentFunctor f{dense_of, (long long*)d_seg_first, cap};
        UPCP_TRY(compact_apply(flags, m, counters + 1, prim_ws, st, f));
    }
    component_init_kernel<<<(cap + kBlock - 1) / kBlock, kBlock, 0, st>>>(o, cap, (long long*)d_seg_first,
                                                                          counters + 1, (long long)m, (long long*)d_info);
    UPCP_LAUNCH_OK();
    component_stats_kernel<<<grid_m, kBlock, 0, st>>>(vals, dense_of, m, d_x, d_y, d_z, d_gz, cap, d_dense, d_order,
                                                      o, (long long*)d_info);
    UPCP_LAUNCH_OK();
    component_decode_kernel<<<(cap + kBlock - 1) / kBlock, kBlock, 0, st>>>(o, cap, counters + 1);
    UPCP_LAUNCH_OK();
    uint32_t hk = 0;
    UPCP_CUDA_OK(cudaMemcpyAsync(&hk, counters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    UPCP_CUDA_OK(cudaStreamSynchronize(st));
    if (hk > (uint32_t)cap) UPCP_FAIL(UPCP_E_CAPACITY, "component_stats: more clusters than the output capacity");
    return UPCP_OK;
}

int upcp_gather_xy(const double* d_x, const double* d_y, const uint32_t* d_idx, int64_t cnt, double* d_out_xy,
                   void* stream) {
    if (cnt < 0 || (cnt > 0 && (!d_x || !d_y || !d_idx || !d_out_xy))) UPCP_FAIL(UPCP_E_INVALID, "gather_xy: bad argument");
    if (cnt == 0) return UPCP_OK;
    gather_xy_kernel<<<grid_for(cnt, kBlock), kBlock, 0, (cudaStream_t)stream>>>(d_x, d_y, d_idx, cnt, d_out_xy);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

int upcp_component_mark(const int32_t* d_dense, int64_t n, const uint8_t* d_accepted, int32_t n_comp,
                        uint8_t* d_out, void* stream) {
    if (n < 0 || n_comp < 0 || (n > 0 && (!d_dense || !d_out)) || (n_comp > 0 && !d_accepted))
        UPCP_FAIL(UPCP_E_INVALID, "component_mark: bad argument");
    if (n == 0) return UPCP_OK;
    component_mark_kernel<<<grid_for(n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(d_dense, n, d_accepted, n_comp, d_out);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

int upcp_prism_clip(const double* d_x, const double* d_y, const double* d_z, int64_t n, const uint8_t* d_sel,
                    const double* d_ring_xy, const int32_t* d_ring_off, const double* d_zrange, int32_t n_prisms,
                    uint8_t* d_out, void* stream) {
    if (n < 0 || n_prisms < 0 || (n > 0 && (!d_x || !d_y || !d_z || !d_out)) ||
        (n_prisms > 0 && (!d_ring_xy || !d_ring_off || !d_zrange)))
        UPCP_FAIL(UPCP_E_INVALID, "prism_clip: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return UPCP_OK;
    UPCP_CUDA_OK(cudaMemsetAsync(d_out, 0, (size_t)n, st));
    for (int p0 = 0; p0 < n_prisms; p0 += kPrismChunk) {
        const int p1 = p0 + kPrismChunk < n_prisms ? p0 + kPrismChunk : n_prisms;
        prism_clip_kernel<<<grid_for(n, kBlock), kBlock, 0, st>>>(d_x, d_y, d_z, n, d_sel, d_ring_xy, d_ring_off,
                                                                  d_zrange, p0, p1, d_out);
        UPCP_LAUNCH_OK();
    }
    return UPCP_OK;
}

}  // extern "C"
