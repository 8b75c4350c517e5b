// S3 -- connected components of occupied octree cells (LabelConnectedComp).
//
// Restates what the reference obtains from CloudCompare through
//   pycc.ccPointCloud(xs, ys, zs)  +  cccorelib.AutoSegmentationTools
//   .labelConnectedComponents(cloud, level)            (label_connected_comp.py:53-97)
// followed by the min_component_size filter (:92-97) and the seed-fraction fill of
// _fill_components (:99-135):
//   1. selected points are cast to float32 (no shift) and assigned to cells of the
//      cubical octree exactly as DgmOctree does [ext, unverified -- see oracle/];
//   2. voxel keys (z,y,x packed, level bits per axis) are radix-sorted together with the
//      point index -> cell list (unique keys + first-point offsets) in HBM;
//   3. connected components of occupied cells under 26-connectivity by lock-free
//      union-find: every cell hooks onto its 13 "forward" neighbours (found by binary
//      search in the sorted unique-key array), roots are the minimum cell index;
//   4. component sizes and seed counts by atomic histograms on the root;
//   5. per-point component id / fill flag scattered back in original point order.
#include <math.h>
#include "common.cuh"
#include "primitives.cuh"
#include "upcp_math.cuh"

namespace upcp {

constexpr int kBlock = 256;
constexpr int kRowTableMaxLevel = 12;    // dense (y, z) row table of the link step: 4^level entries
static inline size_t row_table_rows(int level, int64_t m) {
    if (level > kRowTableMaxLevel) return 0;
    const size_t rows = (size_t)1 << (2 * level);
    return rows <= 8 * (size_t)m + 65536 ? rows : 0;        // not worth zeroing for a handful of cells
}

struct OctreeParams {
    float min_x, min_y, min_z, cs;
    int level, shift, max_cells_minus1;   // shift = max_level - level
};

template <typename K>
__device__ __forceinline__ K pack_key(int cx, int cy, int cz, int level) {
    return ((K)cz << (2 * level)) | ((K)cy << level) | (K)cx;
}

struct CompactIdxFunctor {
    uint32_t* sel_idx;
    __device__ void operator()(int64_t i, bool selected, uint32_t rank) const {
        if (selected) sel_idx[rank] = (uint32_t)i;
    }
};

template <typename K>
__global__ void __launch_bounds__(kBlock)
cc_keys_kernel(const double* __restrict__ x, const double* __restrict__ y,
               const double* __restrict__ z, const uint32_t* __restrict__ sel_idx, int64_t m,
               OctreeParams p, K* __restrict__ keys, uint32_t* __restrict__ vals) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x; j < m; j += stride) {
        int64_t i = sel_idx ? (int64_t)sel_idx[j] : j;
        // float32 cast of the unshifted coordinates (label_connected_comp.py:59-64)
        float fx = (float)x[i], fy = (float)y[i], fz = z ? (float)z[i] : 0.0f;
        int cx = octree_cell_pos(fx, p.min_x, p.cs, p.max_cells_minus1) >> p.shift;
        int cy = octree_cell_pos(fy, p.min_y, p.cs, p.max_cells_minus1) >> p.shift;
        int cz = octree_cell_pos(fz, p.min_z, p.cs, p.max_cells_minus1) >> p.shift;
        keys[j] = pack_key<K>(cx, cy, cz, p.level);
        vals[j] = (uint32_t)j;
    }
}

template <typename K>
__global__ void __launch_bounds__(kBlock)
cc_boundary_flags_kernel(const K* __restrict__ keys, int64_t m, uint8_t* __restrict__ flags) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t s = (int64_t)blockIdx.x * kBlock + threadIdx.x; s < m; s += stride)
        flags[s] = (s == 0 || keys[s] != keys[s - 1]) ? 1 : 0;
}

template <typename K>
struct CellFunctor {
    const K* keys;
    K* ucell;
    uint32_t* cell_first;
    uint32_t* cell_of;
    __device__ void operator()(int64_t s, bool is_first, uint32_t rank) const {
        if (is_first) { ucell[rank] = keys[s]; cell_first[rank] = (uint32_t)s; }
        cell_of[s] = rank + (is_first ? 1u : 0u) - 1u;
    }
};

__global__ void cc_finish_cells_kernel(const uint32_t* __restrict__ n_cells, uint32_t m,
                                       uint32_t* __restrict__ cell_first) {
    if (threadIdx.x == 0 && blockIdx.x == 0) cell_first[*n_cells] = m;
}

__global__ void __launch_bounds__(kBlock)
cc_init_parent_kernel(const uint32_t* __restrict__ n_cells, uint32_t* __restrict__ parent,
                      int32_t* __restrict__ comp_size, int32_t* __restrict__ seed_cnt) {
    const uint32_t c_total = *n_cells;
    uint32_t stride = gridDim.x * kBlock;
    for (uint32_t c = blockIdx.x * kBlock + threadIdx.x; c < c_total; c += stride) {
        parent[c] = c; comp_size[c] = 0; seed_cnt[c] = 0;
    }
}

// number of occupied cells per grid row (y, z); its exclusive prefix is the row table
template <typename K>
__global__ void __launch_bounds__(kBlock)
cc_row_count_kernel(const K* __restrict__ ucell, const uint32_t* __restrict__ n_cells, int level,
                    uint32_t* __restrict__ row_count) {
    const uint32_t c_total = *n_cells;
    uint32_t stride = gridDim.x * kBlock;
    for (uint32_t c = blockIdx.x * kBlock + threadIdx.x; c < c_total; c += stride)
        atomicAdd(row_count + (uint32_t)(ucell[c] >> level), 1u);
}

__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t a) {
    // path halving; racing writers only ever shorten paths towards smaller indices
    // L2 (.cg) loads: other SMs hook roots concurrently, never trust a stale L1 line
    uint32_t p = __ldcg(parent + a);
    while (p != a) {
        uint32_t gp = __ldcg(parent + p);
        if (gp != p) parent[a] = gp;
        a = p;
        p = gp;
    }
    return a;
}

__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
    // hook the larger root under the smaller one: roots end up as component minima
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }      // a > b
        uint32_t old = atomicCAS(&parent[a], a, b);
        if (old == a) return;
        a = old;                                          // lost the race: retry from there
    }
}

// first index in ucell[lo, hi) with ucell[idx] >= key
template <typename K>
__device__ __forceinline__ uint32_t lower_bound_key(const K* __restrict__ ucell, uint32_t lo,
                                                    uint32_t hi, K key) {
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (ucell[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <typename K>
__global__ void __launch_bounds__(kBlock)
cc_link_kernel(const K* __restrict__ ucell, const uint32_t* __restrict__ n_cells, int level,
               const uint32_t* __restrict__ row_first, uint32_t* __restrict__ parent) {
    const uint32_t c_total = *n_cells;
    const int dim = 1 << level;
    const K axis_mask = ((K)1 << level) - 1;
    const uint32_t stride = gridDim.x * kBlock;
    const uint32_t lane = threadIdx.x & 31;
    // The loop runs warp-uniform (every lane the same number of rounds) and the lanes are brought back together
    // with __syncwarp() after every union: a lane that loses a compare-and-swap race or walks a long parent
    // chain would otherwise run ahead of / behind the others for the REST of the body (independent thread
    // scheduling does not reconverge on its own here: ncu showed 1.6 of 32 lanes per issued instruction).
    for (uint32_t cw = blockIdx.x * kBlock + threadIdx.x - lane; cw < c_total; cw += stride) {
        const uint32_t c = cw + lane;
        const bool valid = c < c_total;
        const K key = valid ? ucell[c] : (K)0;
        const int cx = (int)(key & axis_mask);
        const int cy = (int)((key >> level) & axis_mask);
        const int cz = (int)((key >> (2 * level)) & axis_mask);
        // (dz, dy, dx) = (0, 0, +1): next entry of the sorted unique-key array
        if (valid && cx + 1 < dim && c + 1 < c_total && ucell[c + 1] == key + 1) uf_union(parent, c, c + 1);
        __syncwarp();
        // is the cell left of this one occupied (this cell continues an x-run of its row)?
        const bool own_left = valid && cx > 0 && c > 0 && ucell[c - 1] == key - 1;
        // rows (dz, dy) in {(0,+1), (+1,-1), (+1,0), (+1,+1)}: x in [cx-1, cx+1]
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int dz = r == 0 ? 0 : 1;
            const int dy = r == 0 ? 1 : r - 2;
            const int ny = cy + dy, nz = cz + dz;
            // which of the row's cells at x - 1 (L), x (M), x + 1 (R) are occupied, and where
            bool pl = false, pm = false, pr = false;
            uint32_t il = 0, im = 0, ir = 0;
            if (valid && !(ny < 0 || ny >= dim || nz >= dim)) {
                const int x0 = cx > 0 ? cx - 1 : 0;
                const int x1 = cx + 1 < dim ? cx + 1 : dim - 1;
                const K k0 = pack_key<K>(x0, ny, nz, level);
                const K k1 = pack_key<K>(x1, ny, nz, level);
                const K km = pack_key<K>(cx, ny, nz, level);
                // the cells of grid row (ny, nz) are one contiguous run of the sorted list: with the row
                // table (levels <= 12) the search is confined to that run, else to everything after c
                uint32_t lo = c + 1, hi = c_total;
                if (row_first) {
                    const uint32_t rowid = (uint32_t)nz * (uint32_t)dim + (uint32_t)ny;
                    lo = __ldg(row_first + rowid); hi = __ldg(row_first + rowid + 1);
                }
                uint32_t idx = lower_bound_key<K>(ucell, lo, hi, k0);
                for (int t = 0; t < 3 && idx < hi; ++t, ++idx) {
                    const K v = ucell[idx];
                    if (v > k1) break;
                    if (v < km) { pl = true; il = idx; }
                    else if (v == km) { pm = true; im = idx; }
                    else { pr = true; ir = idx; }
                }
            }
            __syncwarp();
            // Every pair of occupied cells with |dx| <= 1 is an edge of the cell graph, but x-neighbours of one
            // row are already joined (the union above), so ONE union per pair of adjacent x-runs gives the same
            // components: the first cell of this run joins the run(s) it touches, a later cell only a run that
            // STARTS at x + 1 (R occupied, M empty) -- any run through L or M already touches the cell at x - 1.
            if (!own_left && pl) uf_union(parent, c, il);
            if (!own_left && pm && !pl) uf_union(parent, c, im);
            if (pr && !pm) uf_union(parent, c, ir);
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(kBlock)
cc_compress_kernel(const uint32_t* __restrict__ n_cells, uint32_t* __restrict__ parent,
                   const uint32_t* __restrict__ cell_first, int32_t* __restrict__ comp_size,
                   unsigned long long* __restrict__ info) {
    const uint32_t c_total = *n_cells;
    uint32_t stride = gridDim.x * kBlock;
    unsigned int roots = 0;
    for (uint32_t c0 = blockIdx.x * kBlock; c0 < c_total; c0 += stride) {
        uint32_t c = c0 + threadIdx.x;
        uint32_t r = 0xffffffffu;
        int32_t cnt = 0;
        if (c < c_total) {
            r = c;
            while (true) { uint32_t p = parent[r]; if (p == r) break; r = p; }
            parent[c] = r;      // full compression (roots never change after linking)
            cnt = (int32_t)(cell_first[c + 1] - cell_first[c]);
            roots += (r == c);
        }
        // component-size histogram with warp-aggregated atomics: neighbouring cells of the
        // sorted cell list mostly share a root, so one atomic per distinct root per warp
        uint32_t peers = __match_any_sync(0xffffffffu, r);
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(peers) - 1;
        int32_t sum = cnt;
        if (peers == 0xffffffffu) {                 // warp-uniform: the whole warp shares a root
            sum = __reduce_add_sync(0xffffffffu, cnt);
        } else {                                    // leader gathers its peers one by one
            uint32_t rest = peers & ~(1u << leader);
            while (__any_sync(0xffffffffu, rest != 0)) {
                int src = rest ? (__ffs(rest) - 1) : lane;
                int32_t v = __shfl_sync(0xffffffffu, cnt, src);
                if (lane == leader && rest) sum += v;
                rest &= rest - 1;
            }
        }
        if (c < c_total && lane == leader) atomicAdd(&comp_size[r], sum);
    }
    for (int o = 16; o > 0; o >>= 1) roots += __shfl_xor_sync(0xffffffffu, roots, o);
    if ((threadIdx.x & 31) == 0 && roots) atomicAdd(&info[2], (unsigned long long)roots);
}

__global__ void __launch_bounds__(kBlock)
cc_seed_count_kernel(const uint32_t* __restrict__ sel_idx, const uint32_t* __restrict__ perm,
                     const uint32_t* __restrict__ cell_of, const uint32_t* __restrict__ parent,
                     const int32_t* __restrict__ labels, int32_t seed_label, int64_t m,
                     int32_t* __restrict__ seed_cnt) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t s0 = (int64_t)blockIdx.x * kBlock; s0 < m; s0 += stride) {
        int64_t s = s0 + threadIdx.x;
        bool is_seed = false;
        uint32_t root = 0xffffffffu;
        if (s < m) {
            uint32_t j = perm[s];
            int64_t i = sel_idx ? (int64_t)sel_idx[j] : (int64_t)j;
            is_seed = labels[i] == seed_label;
            root = parent[cell_of[s]];
        }
        // warp-aggregated atomics: sorted order => lanes mostly share one root
        uint32_t key = is_seed ? root : 0xffffffffu;
        uint32_t peers = __match_any_sync(0xffffffffu, key);
        if (is_seed && (threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1))
            atomicAdd(&seed_cnt[root], __popc(peers));
    }
}

__global__ void __launch_bounds__(kBlock)
cc_count_kept_kernel(const uint32_t* __restrict__ n_cells, const uint32_t* __restrict__ parent,
                     const int32_t* __restrict__ comp_size, int32_t min_size,
                     unsigned long long* __restrict__ info) {
    const uint32_t c_total = *n_cells;
    uint32_t stride = gridDim.x * kBlock;
    unsigned int kept = 0;
    for (uint32_t c = blockIdx.x * kBlock + threadIdx.x; c < c_total; c += stride)
        kept += (parent[c] == c && comp_size[c] >= min_size);
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(&info[3], (unsigned long long)kept);
    if (blockIdx.x == 0 && threadIdx.x == 0) info[1] = c_total;
}

__global__ void __launch_bounds__(kBlock)
cc_output_kernel(const uint32_t* __restrict__ sel_idx, const uint32_t* __restrict__ perm,
                 const uint32_t* __restrict__ cell_of, const uint32_t* __restrict__ parent,
                 const int32_t* __restrict__ comp_size, const int32_t* __restrict__ seed_cnt,
                 int64_t m, int32_t min_size, double threshold, bool do_fill,
                 int32_t* __restrict__ comp, uint8_t* __restrict__ fill) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t s = (int64_t)blockIdx.x * kBlock + threadIdx.x; s < m; s += stride) {
        uint32_t j = perm[s];
        int64_t i = sel_idx ? (int64_t)sel_idx[j] : (int64_t)j;
        uint32_t r = parent[cell_of[s]];
        int32_t sz = comp_size[r];
        bool keep = sz >= min_size;                       // label_connected_comp.py:92-97
        if (comp) comp[i] = keep ? (int32_t)r : -1;
        if (do_fill) {
            // label_connected_comp.py:123: float(seed_count) / cc_size > threshold
            bool f = keep && ((double)seed_cnt[r] / (double)sz > threshold);
            fill[i] = f ? 1 : 0;
        }
    }
}

__global__ void cc_set_info_kernel(long long* __restrict__ info, long long m) {
    if (threadIdx.x == 0 && blockIdx.x == 0) info[0] = m;
}

template <typename K>
static int lcc_impl(const double* d_x, const double* d_y, const double* d_z, int64_t n,
                    const uint8_t* d_sel, OctreeParams op, int32_t min_size,
                    const int32_t* d_labels, int32_t seed_label, double threshold,
                    int32_t* d_comp, uint8_t* d_fill, int64_t* d_info,
                    void* d_ws, size_t ws_bytes, cudaStream_t st) {
    Workspace ws(d_ws, ws_bytes);
    uint32_t* counters = ws.take<uint32_t>(64);           // [0] m, [1] n_cells
    uint32_t* sel_idx = d_sel ? ws.take<uint32_t>((size_t)n) : nullptr;
    uint32_t* prim_ws = ws.take<uint32_t>(compact_ws_words(n) + sort_ws_words(n) + scan_ws_words(((int64_t)1 << (2 * kRowTableMaxLevel)) + 1));
    if (!ws.ok) UPCP_FAIL(UPCP_E_WORKSPACE, "lcc: workspace too small");

    UPCP_CUDA_OK(cudaMemsetAsync(d_info, 0, 4 * sizeof(int64_t), st));
    if (d_comp) UPCP_CUDA_OK(cudaMemsetAsync(d_comp, 0xff, (size_t)n * sizeof(int32_t), st));
    if (d_fill) UPCP_CUDA_OK(cudaMemsetAsync(d_fill, 0, (size_t)n, st));

    // 1. compact the selection -> sel_idx (ascending original index)
    int64_t m = n;
    if (d_sel) {
        CompactIdxFunctor f{sel_idx};
        UPCP_TRY(compact_apply(d_sel, n, counters, prim_ws, st, f));
        uint32_t hm = 0;
        UPCP_CUDA_OK(cudaMemcpyAsync(&hm, counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        UPCP_CUDA_OK(cudaStreamSynchronize(st));
        m = hm;
    }
    cc_set_info_kernel<<<1, 32, 0, st>>>((long long*)d_info, (long long)m);      // d_info[0] = m, no host round trip
    UPCP_LAUNCH_OK();
    if (m == 0) return UPCP_OK;

    K* keys_a = ws.take<K>((size_t)m);
    K* keys_b = ws.take<K>((size_t)m);
    uint32_t* vals_a = ws.take<uint32_t>((size_t)m);
    uint32_t* vals_b = ws.take<uint32_t>((size_t)m);
    uint8_t* flags = ws.take<uint8_t>((size_t)m);
    uint32_t* cell_of = ws.take<uint32_t>((size_t)m);
    uint32_t* cell_first = ws.take<uint32_t>((size_t)m + 1);
    uint32_t* parent = ws.take<uint32_t>((size_t)m);
    int32_t* comp_size = ws.take<int32_t>((size_t)m);
    int32_t* seed_cnt = ws.take<int32_t>((size_t)m);
    const size_t n_rows = row_table_rows(op.level, m);
    uint32_t* row_first = n_rows ? ws.take<uint32_t>(n_rows + 1) : nullptr;
    if (!ws.ok) UPCP_FAIL(UPCP_E_WORKSPACE, "lcc: workspace too small");

    // 2. voxel keys + radix sort
    const int grid_m = grid_for(m, kBlock);
    cc_keys_kernel<K><<<grid_m, kBlock, 0, st>>>(d_x, d_y, d_z, sel_idx, m, op, keys_a, vals_a);
    UPCP_LAUNCH_OK();
    bool in_b = false;
    if (sizeof(K) == 4)
        UPCP_TRY(radix_sort_pairs_u32((uint32_t*)keys_a, vals_a, (uint32_t*)keys_b, vals_b, m, 3 * op.level, prim_ws, st, &in_b));
    else
        UPCP_TRY(radix_sort_pairs_u64((uint64_t*)keys_a, vals_a, (uint64_t*)keys_b, vals_b, m, 3 * op.level, prim_ws, st, &in_b));
    K* keys = in_b ? keys_b : keys_a;
    uint32_t* perm = in_b ? vals_b : vals_a;
    K* ucell = in_b ? keys_a : keys_b;                     // the other key buffer is free now

    // 3. cell list
    cc_boundary_flags_kernel<K><<<grid_m, kBlock, 0, st>>>(keys, m, flags);
    UPCP_LAUNCH_OK();
    uint32_t* n_cells = counters + 1;
    {
        CellFunctor<K> f{keys, ucell, cell_first, cell_of};
        UPCP_TRY(compact_apply(flags, m, n_cells, prim_ws, st, f));
    }
    cc_finish_cells_kernel<<<1, 32, 0, st>>>(n_cells, (uint32_t)m, cell_first);
    UPCP_LAUNCH_OK();

    // 4. union-find over cells
    cc_init_parent_kernel<<<grid_m, kBlock, 0, st>>>(n_cells, parent, comp_size, seed_cnt);
    UPCP_LAUNCH_OK();
    prof_begin(UPCP_PROF_LINK, st);
    if (row_first) {
        UPCP_CUDA_OK(cudaMemsetAsync(row_first, 0, (n_rows + 1) * 4, st));
        cc_row_count_kernel<K><<<grid_m, kBlock, 0, st>>>(ucell, n_cells, op.level, row_first);
        UPCP_LAUNCH_OK();
        UPCP_TRY(exclusive_scan_u32(row_first, row_first, (int64_t)n_rows + 1, nullptr, prim_ws, st));
    }
    cc_link_kernel<K><<<grid_m, kBlock, 0, st>>>(ucell, n_cells, op.level, row_first, parent);
    prof_end(UPCP_PROF_LINK, st, m);
    UPCP_LAUNCH_OK();
    cc_compress_kernel<<<grid_m, kBlock, 0, st>>>(n_cells, parent, cell_first, comp_size,
                                                  (unsigned long long*)d_info);
    UPCP_LAUNCH_OK();
    cc_count_kept_kernel<<<grid_m, kBlock, 0, st>>>(n_cells, parent, comp_size, min_size,
                                                    (unsigned long long*)d_info);
    UPCP_LAUNCH_OK();

    // 5. seed histogram + per-point output
    const bool do_fill = d_labels != nullptr && d_fill != nullptr;
    if (do_fill) {
        cc_seed_count_kernel<<<grid_m, kBlock, 0, st>>>(sel_idx, perm, cell_of, parent, d_labels,
                                                        seed_label, m, seed_cnt);
        UPCP_LAUNCH_OK();
    }
    cc_output_kernel<<<grid_m, kBlock, 0, st>>>(sel_idx, perm, cell_of, parent, comp_size, seed_cnt,
                                                m, min_size, threshold, do_fill, d_comp, d_fill);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

}  // namespace upcp

using namespace upcp;

extern "C" {

int upcp_octree_params(const double* h_bbox_sel6, int max_level, float* h_out4) {
    if (!h_bbox_sel6 || !h_out4 || max_level < 1 || max_level > 21)
        UPCP_FAIL(UPCP_E_INVALID, "octree_params: bad argument");
    // float32 bbox of the float32 cloud (the cast is monotone, so min/max commute with it)
    volatile float mn[3], mx[3];
    for (int d = 0; d < 3; ++d) { mn[d] = (float)h_bbox_sel6[d]; mx[d] = (float)h_bbox_sel6[3 + d]; }
    // CCMiscTools::MakeMinAndMaxCubical(dimMin, dimMax, 0.01)   [ext, unverified]
    volatile float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    volatile float max_dd = dx > dy ? dx : dy;
    max_dd = max_dd > dz ? max_dd : dz;
    max_dd = (float)((double)max_dd * (1.0 + 0.01));
    for (int d = 0; d < 3; ++d) {
        volatile float md = mx[d] + mn[d];
        volatile float diff = md - max_dd;
        h_out4[d] = diff * 0.5f;
    }
    // DgmOctree::updateCellSizeTable: cellSize[0] = dimMax.x - dimMin.x, halved per level
    volatile float dim_max_x = h_out4[0] + max_dd;
    volatile float cs = dim_max_x - h_out4[0];
    for (int k = 0; k < max_level; ++k) cs = cs / 2.0f;
    h_out4[3] = cs;
    return UPCP_OK;
}

int upcp_octree_level(const double* h_bbox_all6, int ncols, double grid_size) {
    // math_utils.get_octree_level (math_utils.py:21-33)
    double max_dim = 0.0;
    for (int d = 0; d < ncols && d < 3; ++d) {
        double e = h_bbox_all6[3 + d] - h_bbox_all6[d];
        if (d == 0 || e > max_dim) max_dim = e;
    }
    if (max_dim < 0.001) return 0;
    double lvl = rint(-log(grid_size / max_dim) / log(2.0));
    if (lvl > 0) return (int)lvl;
    return 1;
}

size_t upcp_lcc_ws_bytes(int64_t n, int64_t n_selected_upper, int level) {
    if (n < 1) n = 1;
    int64_t m = n_selected_upper < 1 || n_selected_upper > n ? n : n_selected_upper;
    size_t key_bytes = (3 * level <= 32) ? 4 : 8;
    size_t bytes = 0;
    bytes += align_up(64 * 4);
    bytes += align_up((size_t)n * 4);                                        // sel_idx
    bytes += align_up((compact_ws_words(n) + sort_ws_words(n) + scan_ws_words(((int64_t)1 << (2 * kRowTableMaxLevel)) + 1)) * 4);   // primitives
    bytes += 2 * align_up((size_t)m * key_bytes);                             // keys a/b
    bytes += 2 * align_up((size_t)m * 4);                                     // vals a/b
    bytes += align_up((size_t)m);                                             // flags
    bytes += align_up((size_t)m * 4);                                         // cell_of
    bytes += align_up(((size_t)m + 1) * 4);                                   // cell_first
    bytes += 3 * align_up((size_t)m * 4);                                     // parent, sizes, seeds
    bytes += align_up((row_table_rows(level, m) + 1) * 4);                      // row table of the link step
    return bytes + 4096;
}

int upcp_lcc(const double* d_x, const double* d_y, const double* d_z, int64_t n,
             const uint8_t* d_sel, const float* h_octree4, int level, int max_level,
             int32_t min_component_size, const int32_t* d_labels, int32_t seed_label,
             double threshold, int32_t* d_comp, uint8_t* d_fill, int64_t* d_info,
             void* d_ws, size_t ws_bytes, void* stream) {
    if (n < 0 || !h_octree4 || !d_info) UPCP_FAIL(UPCP_E_INVALID, "lcc: bad argument");
    if (max_level < 1 || max_level > 21 || level < 0 || level > max_level)
        UPCP_FAIL(UPCP_E_INVALID, "lcc: level must be in [0, max_level], max_level <= 21");
    if (n >= ((int64_t)1 << 31)) UPCP_FAIL(UPCP_E_CAPACITY, "lcc: n must be < 2^31");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        UPCP_CUDA_OK(cudaMemsetAsync(d_info, 0, 4 * sizeof(int64_t), st));
        return UPCP_OK;
    }
    OctreeParams op;
    op.min_x = h_octree4[0]; op.min_y = h_octree4[1]; op.min_z = h_octree4[2]; op.cs = h_octree4[3];
    op.level = level; op.shift = max_level - level; op.max_cells_minus1 = (1 << max_level) - 1;
    if (3 * level <= 32)
        return lcc_impl<uint32_t>(d_x, d_y, d_z, n, d_sel, op, min_component_size, d_labels, seed_label,
                                  threshold, d_comp, d_fill, d_info, d_ws, ws_bytes, st);
    return lcc_impl<uint64_t>(d_x, d_y, d_z, n, d_sel, op, min_component_size, d_labels, seed_label,
                              threshold, d_comp, d_fill, d_info, d_ws, ws_bytes, st);
}

}  // extern "C"
