// Exclusive scan, flag compaction and the voxel-key LSD radix sort.
//
// Radix sort (the "Morton/voxel-key radix sort" of the hot path): 8-bit digits, one
// histogram kernel + one scan + one scatter kernel per pass.  The scatter kernel ranks
// a 4096-key tile in shared memory (warp-striped, __match_any_sync multi-split,
// per-warp digit counters), reorders keys and payloads through shared memory so that
// the global writes are contiguous runs per digit, and is stable.  Per pass the HBM
// traffic is 4 B/key (histogram read) + 2 x (key + 4 B payload) (scatter read + write).
#include "primitives.cuh"

namespace upcp {

// ================================================================== scan ========
// Level kernel A: per-tile sums.
__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    const int64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * kScanTile;
        uint32_t sum = 0;
#pragma unroll 4
        for (int k = 0; k < kScanTile / kScanThreads; ++k) {
            int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
            if (i < n) sum += in[i];
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < kScanThreads / 32; ++w) t += s_warp[w];
            tile_sums[tile] = t;
        }
        __syncthreads();
    }
}

// Single-CTA exclusive scan of a (short) array, in place, with running carry.
__global__ void __launch_bounds__(kScanThreads)
scan_single_kernel(uint32_t* __restrict__ data, int64_t n, uint32_t* __restrict__ total) {
    __shared__ uint32_t s_warp[9];
    uint32_t carry = 0;
    for (int64_t base = 0; base < n; base += kScanThreads) {
        int64_t i = base + threadIdx.x;
        uint32_t v = i < n ? data[i] : 0;
        uint32_t tot;
        uint32_t ex = block_exclusive_scan(v, s_warp, &tot);
        if (i < n) data[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

// Level kernel C: per-tile exclusive scan + tile offset.
__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n,
                  const uint32_t* __restrict__ tile_offsets) {
    __shared__ uint32_t s_warp[9];
    const int64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint32_t carry = tile_offsets[tile];
        const int64_t base = tile * kScanTile;
        // 4 sub-tiles of 1024: each thread owns 4 consecutive elements
        for (int sub = 0; sub < kScanTile / (kScanThreads * 4); ++sub) {
            int64_t i0 = base + (int64_t)sub * kScanThreads * 4 + (int64_t)threadIdx.x * 4;
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? in[i0 + k] : 0;
            uint32_t s = v[0] + v[1] + v[2] + v[3];
            uint32_t tot;
            uint32_t ex = block_exclusive_scan(s, s_warp, &tot) + carry;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (i0 + k < n) out[i0 + k] = ex;
                ex += v[k];
            }
            carry += tot;
        }
    }
}

size_t scan_ws_words(int64_t n) {
    // level-1 tile sums, recursively (levels shrink by 4096x; two levels cover 2^36)
    size_t words = 0;
    int64_t m = n;
    while (m > kScanTile) {
        m = (m + kScanTile - 1) / kScanTile;
        words += align_up((size_t)m * 4) / 4;
    }
    return words + 64;
}

int exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out, int64_t n, uint32_t* d_total,
                       uint32_t* d_ws, cudaStream_t st) {
    if (n <= 0) {
        if (d_total) UPCP_CUDA_OK(cudaMemsetAsync(d_total, 0, sizeof(uint32_t), st));
        return UPCP_OK;
    }
    if (n <= kScanTile) {
        if (d_in != d_out)
            UPCP_CUDA_OK(cudaMemcpyAsync(d_out, d_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        scan_single_kernel<<<1, kScanThreads, 0, st>>>(d_out, n, d_total);
        UPCP_LAUNCH_OK();
        return UPCP_OK;
    }
    const int64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    uint32_t* tile_sums = d_ws;
    uint32_t* next_ws = d_ws + align_up((size_t)n_tiles * 4) / 4;
    int grid = (int)(n_tiles < (int64_t)num_sms() * 8 ? n_tiles : (int64_t)num_sms() * 8);
    scan_reduce_kernel<<<grid, kScanThreads, 0, st>>>(d_in, n, tile_sums);
    UPCP_LAUNCH_OK();
    UPCP_TRY(exclusive_scan_u32(tile_sums, tile_sums, n_tiles, d_total, next_ws, st));
    scan_apply_kernel<<<grid, kScanThreads, 0, st>>>(d_in, d_out, n, tile_sums);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

// ============================================================ compaction ========
size_t compact_ws_words(int64_t n) {
    int64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    if (n_tiles < 1) n_tiles = 1;
    return align_up((size_t)n_tiles * 4) / 4 + scan_ws_words(n_tiles) + 64;
}

__global__ void __launch_bounds__(kScanThreads)
compact_count_kernel(const uint8_t* __restrict__ flags, int64_t n, uint32_t* __restrict__ tile_counts) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    const int64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    const bool aligned = ((uintptr_t)flags & 15) == 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * kScanTile + (int64_t)threadIdx.x * 16;
        uint32_t cnt = 0;
        if (aligned && base + 16 <= n) {
            uint4 v = *reinterpret_cast<const uint4*>(flags + base);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 16; ++k) cnt += ((w[k >> 2] >> (8 * (k & 3))) & 0xffu) != 0;
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) cnt += (base + k < n) ? (flags[base + k] != 0) : 0;
        }
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < kScanThreads / 32; ++w) t += s_warp[w];
            tile_counts[tile] = t;
        }
        __syncthreads();
    }
}

// ============================================================ radix sort ========
constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;   // 4096 keys per CTA
constexpr int kRsBits = 8;
constexpr int kRsBins = 1 << kRsBits;

template <typename K>
__global__ void __launch_bounds__(kRsThreads)
rs_hist_kernel(const K* __restrict__ keys, int64_t n, int shift, uint32_t mask,
               uint32_t* __restrict__ hist /*[bins][n_tiles]*/, int n_tiles) {
    __shared__ uint32_t s_hist[kRsBins];
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int i = threadIdx.x; i < kRsBins; i += kRsThreads) s_hist[i] = 0;
        __syncthreads();
        const int64_t base = (int64_t)tile * kRsTile;
#pragma unroll 4
        for (int k = 0; k < kRsItems; ++k) {
            int64_t i = base + (int64_t)k * kRsThreads + threadIdx.x;
            if (i < n) atomicAdd(&s_hist[(uint32_t)(keys[i] >> shift) & mask], 1u);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kRsBins; i += kRsThreads)
            hist[(size_t)i * n_tiles + tile] = s_hist[i];
        __syncthreads();
    }
}

template <typename K>
__global__ void __launch_bounds__(kRsThreads)
rs_scatter_kernel(const K* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  K* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
                  int shift, uint32_t mask, const uint32_t* __restrict__ gbase, int n_tiles) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    K* s_keys = reinterpret_cast<K*>(rs_smem);
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys + kRsTile);
    uint32_t* s_cnt = s_vals + kRsTile;                  // [8 warps][256 digits]
    uint32_t* s_dstart = s_cnt + (kRsThreads / 32) * kRsBins;   // [256]
    uint32_t* s_gofs = s_dstart + kRsBins;               // [256]
    uint32_t* s_warp = s_gofs + kRsBins;                 // [9]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane_lt = (1u << lane) - 1u;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = (int64_t)tile * kRsTile;
        const int64_t rem = n - base;
        const int valid = rem < kRsTile ? (int)rem : kRsTile;

        K key[kRsItems];
        uint32_t val[kRsItems];
        uint32_t rank[kRsItems];
        // warp-striped load: warp w owns [w*512, (w+1)*512), round j covers 32 consecutive keys
#pragma unroll
        for (int j = 0; j < kRsItems; ++j) {
            int idx = warp * (32 * kRsItems) + j * 32 + lane;
            if (idx < valid) { key[j] = keys_in[base + idx]; val[j] = vals_in[base + idx]; }
            else { key[j] = ~(K)0; val[j] = 0; }
        }
        for (int i = threadIdx.x; i < (kRsThreads / 32) * kRsBins; i += kRsThreads) s_cnt[i] = 0;
        __syncthreads();

        uint32_t* my_cnt = s_cnt + warp * kRsBins;
#pragma unroll
        for (int j = 0; j < kRsItems; ++j) {
            uint32_t d = (uint32_t)(key[j] >> shift) & mask;
            uint32_t peers = __match_any_sync(0xffffffffu, d);
            int leader = __ffs(peers) - 1;
            uint32_t before = 0;
            if (lane == leader) { before = my_cnt[d]; my_cnt[d] = before + __popc(peers); }
            before = __shfl_sync(0xffffffffu, before, leader);
            rank[j] = before + __popc(peers & lane_lt);
            __syncwarp();
        }
        __syncthreads();

        {   // thread d: exclusive prefix over warps of digit d, then block scan over digits
            const int d = threadIdx.x;
            uint32_t sum = 0;
#pragma unroll
            for (int w = 0; w < kRsThreads / 32; ++w) {
                uint32_t c = s_cnt[w * kRsBins + d];
                s_cnt[w * kRsBins + d] = sum;
                sum += c;
            }
            uint32_t ex = block_exclusive_scan(sum, s_warp, nullptr);
            s_dstart[d] = ex;
            s_gofs[d] = gbase[(size_t)d * n_tiles + tile] - ex;
        }
        __syncthreads();

#pragma unroll
        for (int j = 0; j < kRsItems; ++j) {
            uint32_t d = (uint32_t)(key[j] >> shift) & mask;
            uint32_t pos = s_dstart[d] + my_cnt[d] + rank[j];
            s_keys[pos] = key[j];
            s_vals[pos] = val[j];
        }
        __syncthreads();

#pragma unroll 4
        for (int k = 0; k < kRsItems; ++k) {
            int p = k * kRsThreads + threadIdx.x;
            if (p < valid) {
                K kk = s_keys[p];
                uint32_t d = (uint32_t)(kk >> shift) & mask;
                uint32_t dst = s_gofs[d] + (uint32_t)p;
                keys_out[dst] = kk;
                vals_out[dst] = s_vals[p];
            }
        }
        __syncthreads();
    }
}

size_t sort_ws_words(int64_t n) {
    int64_t n_tiles = (n + kRsTile - 1) / kRsTile;
    if (n_tiles < 1) n_tiles = 1;
    size_t hist_words = align_up((size_t)kRsBins * n_tiles * 4) / 4;
    return hist_words + scan_ws_words((int64_t)kRsBins * n_tiles) + 64;
}

template <typename K>
static int radix_sort_impl(K* keys_a, uint32_t* vals_a, K* keys_b, uint32_t* vals_b, int64_t n,
                           int end_bit, uint32_t* d_ws, cudaStream_t st, bool* result_in_b) {
    *result_in_b = false;
    if (n <= 1 || end_bit <= 0) return UPCP_OK;
    if (n >= ((int64_t)1 << 32)) UPCP_FAIL(UPCP_E_CAPACITY, "radix sort: n must be < 2^32");
    if (end_bit > (int)sizeof(K) * 8) end_bit = (int)sizeof(K) * 8;
    const int n_tiles = (int)((n + kRsTile - 1) / kRsTile);
    uint32_t* hist = d_ws;
    uint32_t* scan_ws = d_ws + align_up((size_t)kRsBins * n_tiles * 4) / 4;
    const size_t smem = (size_t)kRsTile * (sizeof(K) + 4) + ((kRsThreads / 32) * kRsBins + 2 * kRsBins + 16) * 4;
    // (set on every call: the attribute is per device, and callers may drive several devices / threads)
    UPCP_CUDA_OK(cudaFuncSetAttribute(rs_scatter_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = n_tiles < num_sms() * 4 ? n_tiles : num_sms() * 4;
    K* kin = keys_a; uint32_t* vin = vals_a; K* kout = keys_b; uint32_t* vout = vals_b;
    bool in_b = false;
    prof_begin(UPCP_PROF_SORT, st);
    for (int shift = 0; shift < end_bit; shift += kRsBits) {
        int bits = end_bit - shift < kRsBits ? end_bit - shift : kRsBits;
        uint32_t mask = (1u << bits) - 1u;
        rs_hist_kernel<K><<<grid, kRsThreads, 0, st>>>(kin, n, shift, mask, hist, n_tiles);
        UPCP_LAUNCH_OK();
        UPCP_TRY(exclusive_scan_u32(hist, hist, (int64_t)kRsBins * n_tiles, nullptr, scan_ws, st));
        rs_scatter_kernel<K><<<grid, kRsThreads, smem, st>>>(kin, vin, kout, vout, n, shift, mask, hist, n_tiles);
        UPCP_LAUNCH_OK();
        K* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
        in_b = !in_b;
    }
    // units of this bracket = algorithmic BYTES of the sort: per pass the histogram reads the keys, the
    // scatter reads keys + payload and writes them: (3 sizeof(K) + 8) bytes per element and pass
    const int passes = (end_bit + kRsBits - 1) / kRsBits;
    prof_end(UPCP_PROF_SORT, st, n * (int64_t)passes * (int64_t)(3 * sizeof(K) + 8));
    *result_in_b = in_b;
    return UPCP_OK;
}

int radix_sort_pairs_u32(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                         int64_t n, int end_bit, uint32_t* d_ws, cudaStream_t st, bool* result_in_b) {
    return radix_sort_impl<uint32_t>(keys_a, vals_a, keys_b, vals_b, n, end_bit, d_ws, st, result_in_b);
}
int radix_sort_pairs_u64(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b,
                         int64_t n, int end_bit, uint32_t* d_ws, cudaStream_t st, bool* result_in_b) {
    return radix_sort_impl<uint64_t>(keys_a, vals_a, keys_b, vals_b, n, end_bit, d_ws, st, result_in_b);
}

}  // namespace upcp

using namespace upcp;

extern "C" {

size_t upcp_sort_ws_bytes(int64_t n) { return sort_ws_words(n < 1 ? 1 : n) * 4; }

#define UPCP_SORT_API(NAME, K, IMPL)                                                           \
    int NAME(K* d_keys_in, uint32_t* d_vals_in, K* d_keys_out, uint32_t* d_vals_out, int64_t n, \
             int end_bit, void* d_ws, size_t ws_bytes, void* stream) {                         \
        if (n < 0 || end_bit < 0) UPCP_FAIL(UPCP_E_INVALID, #NAME ": bad argument");           \
        if (n == 0) return UPCP_OK;                                                            \
        if (!d_ws || ws_bytes < upcp_sort_ws_bytes(n)) UPCP_FAIL(UPCP_E_WORKSPACE, #NAME ": workspace too small"); \
        cudaStream_t st = (cudaStream_t)stream;                                                \
        bool in_b = false;                                                                     \
        UPCP_TRY(IMPL(d_keys_in, d_vals_in, d_keys_out, d_vals_out, n, end_bit, (uint32_t*)d_ws, st, &in_b)); \
        if (!in_b) {                                                                           \
            UPCP_CUDA_OK(cudaMemcpyAsync(d_keys_out, d_keys_in, (size_t)n * sizeof(K), cudaMemcpyDeviceToDevice, st)); \
            UPCP_CUDA_OK(cudaMemcpyAsync(d_vals_out, d_vals_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, st)); \
        }                                                                                      \
        return UPCP_OK;                                                                        \
    }
UPCP_SORT_API(upcp_sort_pairs_u32, uint32_t, radix_sort_pairs_u32)
UPCP_SORT_API(upcp_sort_pairs_u64, uint64_t, radix_sort_pairs_u64)

}  // extern "C"
