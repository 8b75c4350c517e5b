// Device-wide primitives used by the cell-list builders: exclusive scan, flag
// compaction (with rank), LSD radix sort of (key, value) pairs.
//
// All are multi-kernel (reduce / scan-of-partials / apply) rather than single-pass
// chained scans: no inter-CTA spin waits, so no co-residency assumptions.
#pragma once
#include "common.cuh"

namespace upcp {

constexpr int kScanThreads = 256;
constexpr int kScanTile = 4096;          // elements per CTA tile (16 per thread)

// Workspace (in uint32 words) the scan needs for n elements.
size_t scan_ws_words(int64_t n);

// out[i] = sum_{j<i} in[i]; *total (device, optional) = sum of all.  in may alias out.
int exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out, int64_t n, uint32_t* d_total,
                       uint32_t* d_ws, cudaStream_t st);

// ---- compaction ---------------------------------------------------------------
// Functor-driven two-phase compaction over n flags (uint8, non-zero = selected).
//   phase 1: per-tile counts; scan; phase 2: f(i, selected, rank) for every i where
//   rank = number of selected j < i.
size_t compact_ws_words(int64_t n);

__global__ void compact_count_kernel(const uint8_t* __restrict__ flags, int64_t n,
                                     uint32_t* __restrict__ tile_counts);

// Block-wide exclusive scan of one uint32 per thread (kScanThreads threads).
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp /*[9]*/,
                                                         uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (kScanThreads / 32) ? s_warp[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < (kScanThreads / 32)) s_warp[lane] = winc - w;   // exclusive warp offsets
        if (lane == (kScanThreads / 32) - 1) s_warp[8] = winc;    // block total
    }
    __syncthreads();
    uint32_t res = s_warp[warp] + inc - v;
    if (total) *total = s_warp[8];
    __syncthreads();
    return res;
}

template <typename F>
__global__ void __launch_bounds__(kScanThreads)
compact_apply_kernel(const uint8_t* __restrict__ flags, int64_t n,
                     const uint32_t* __restrict__ tile_offsets, F f) {
    __shared__ uint32_t s_warp[9];
    const int64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * kScanTile + (int64_t)threadIdx.x * 16;
        uint8_t fl[16];
        uint32_t cnt = 0;
        if (base + 16 <= n) {
            uint4 v = *reinterpret_cast<const uint4*>(flags + base);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                fl[k] = (uint8_t)((w[k >> 2] >> (8 * (k & 3))) & 0xffu);
                cnt += fl[k] != 0;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                fl[k] = (base + k < n) ? flags[base + k] : (uint8_t)0;
                cnt += fl[k] != 0;
            }
        }
        uint32_t rank = block_exclusive_scan(cnt, s_warp, nullptr) + tile_offsets[tile];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (base + k < n) {
                f(base + k, fl[k] != 0, rank);
                rank += fl[k] != 0;
            }
        }
    }
}

// Runs the two phases; *d_total receives the number of selected flags.
template <typename F>
int compact_apply(const uint8_t* d_flags, int64_t n, uint32_t* d_total, uint32_t* d_ws,
                  cudaStream_t st, F f) {
    if (n <= 0) {
        if (d_total) UPCP_CUDA_OK(cudaMemsetAsync(d_total, 0, sizeof(uint32_t), st));
        return UPCP_OK;
    }
    const int64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    uint32_t* tile_counts = d_ws;
    uint32_t* scan_ws = d_ws + align_up((size_t)n_tiles * 4) / 4;
    int grid = (int)(n_tiles < (int64_t)num_sms() * 16 ? n_tiles : (int64_t)num_sms() * 16);
    compact_count_kernel<<<grid, kScanThreads, 0, st>>>(d_flags, n, tile_counts);
    UPCP_LAUNCH_OK();
    UPCP_TRY(exclusive_scan_u32(tile_counts, tile_counts, n_tiles, d_total, scan_ws, st));
    compact_apply_kernel<F><<<grid, kScanThreads, 0, st>>>(d_flags, n, tile_counts, f);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

// ---- radix sort -----------------------------------------------------------------
size_t sort_ws_words(int64_t n);

// Sorts bits [0, end_bit) of the keys, stable, LSD, 8 bits per pass.  Buffers ping-pong
// between (keys_a, vals_a) and (keys_b, vals_b); *result_in_b tells where the result is.
int radix_sort_pairs_u32(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                         int64_t n, int end_bit, uint32_t* d_ws, cudaStream_t st, bool* result_in_b);
int radix_sort_pairs_u64(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b,
                         int64_t n, int end_bit, uint32_t* d_ws, cudaStream_t st, bool* result_in_b);

}  // namespace upcp
