// S4 -- exact k-nearest-neighbour graph + hybrid-search normals + curvature
// covariance over a uniform-grid cell list (RegionGrowing set-up).
//
// Restates what the reference obtains from Open3D (region_growing.py:59-73,86-92,
// 105-109): KDTreeFlann.search_knn_vector_3d(p, k) for every point, estimate_normals
// with KDTreeSearchParamHybrid(radius, max_nn) and compute_mean_and_covariance of the
// k-neighbourhood.  Neighbour ordering is ascending (squared distance, compact id)
// with the squared distance computed as ((dx*dx + dy*dy) + dz*dz) in fp64, no FMA
// [ext, unverified -- see oracle/].
//
// Data structure (all in HBM, built per call):
//   * selected points get a Morton key of their uniform-grid cell (cell size c);
//     (key, compact id) pairs are radix-sorted; coordinates are gathered into sorted
//     SoA arrays so that every cell is a contiguous, coalesced run;
//   * unique cells + first-point offsets (cell list) and an open-addressing hash table
//     cell key -> cell index (4 B slots, L2-resident);
//   * query groups: <= 32 consecutive sorted points of one 4x4x4-cell super-cell.
// Search: one warp per query group, one lane per query.  The warp walks the cells of the
// group's bounding box ring by ring; each lane looks up one cell (hash), cells whose
// distance to the group exceeds every lane's current k-th distance are pruned, the
// points of the surviving cells are staged through shared memory with coalesced loads
// and every lane tests every staged point (conflict-free broadcast reads).  Each lane
// keeps its k best in a shared-memory max-heap; a lane is finished when its k-th
// distance is smaller than its distance to the unexplored region.  The epilogue sorts
// the heap, writes the kNN row, and accumulates the Open3D cumulant covariances /
// FastEigen3x3 normal straight from the sorted list, so the 30-neighbour lists never
// touch HBM.
#include <math.h>
#include "common.cuh"
#include "primitives.cuh"
#include "upcp_math.cuh"

namespace upcp {

constexpr int kBlock = 256;
constexpr int kSearchWarps = 4;
constexpr int kSearchThreads = kSearchWarps * 32;
constexpr uint32_t kEmpty = 0xffffffffu;

struct GridParams {
    double ox, oy, oz, cell, inv_cell, slack;
    int nx, ny, nz;
};

__device__ __forceinline__ uint64_t expand_bits21(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}
__device__ __forceinline__ uint32_t compact_bits21(uint64_t x) {
    x &= 0x1249249249249249ULL;
    x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ULL;
    x = (x ^ (x >> 4)) & 0x100f00f00f00f00fULL;
    x = (x ^ (x >> 8)) & 0x1f0000ff0000ffULL;
    x = (x ^ (x >> 16)) & 0x1f00000000ffffULL;
    x = (x ^ (x >> 32)) & 0x1fffffULL;
    return (uint32_t)x;
}
__device__ __forceinline__ uint64_t morton3(int cx, int cy, int cz) {
    return expand_bits21((uint32_t)cx) | (expand_bits21((uint32_t)cy) << 1) | (expand_bits21((uint32_t)cz) << 2);
}
__device__ __forceinline__ uint64_t hash64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
__device__ __forceinline__ int cell_coord(double v, double origin, double inv_cell, int n) {
    double t = (v - origin) * inv_cell;
    if (!(t > 0.0)) return 0;
    if (t >= (double)n) return n - 1;
    return (int)floor(t);
}

struct KnnCompactIdx {
    uint32_t* sel_idx;
    __device__ void operator()(int64_t i, bool selected, uint32_t rank) const {
        if (selected) sel_idx[rank] = (uint32_t)i;
    }
};

template <typename K>
__global__ void __launch_bounds__(kBlock)
knn_keys_kernel(const double* __restrict__ x, const double* __restrict__ y,
                const double* __restrict__ z, const uint32_t* __restrict__ sel_idx, int64_t m,
                GridParams g, K* __restrict__ keys, uint32_t* __restrict__ vals) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x; j < m; j += stride) {
        int64_t i = sel_idx ? (int64_t)sel_idx[j] : j;
        int cx = cell_coord(x[i], g.ox, g.inv_cell, g.nx);
        int cy = cell_coord(y[i], g.oy, g.inv_cell, g.ny);
        int cz = cell_coord(z[i], g.oz, g.inv_cell, g.nz);
        keys[j] = (K)morton3(cx, cy, cz);
        vals[j] = (uint32_t)j;
    }
}

__global__ void __launch_bounds__(kBlock)
knn_gather_kernel(const double* __restrict__ x, const double* __restrict__ y,
                  const double* __restrict__ z, const uint32_t* __restrict__ sel_idx,
                  const uint32_t* __restrict__ perm, int64_t m,
                  double* __restrict__ sx, double* __restrict__ sy, double* __restrict__ sz) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t s = (int64_t)blockIdx.x * kBlock + threadIdx.x; s < m; s += stride) {
        uint32_t j = perm[s];
        int64_t i = sel_idx ? (int64_t)sel_idx[j] : (int64_t)j;
        sx[s] = x[i]; sy[s] = y[i]; sz[s] = z[i];
    }
}

// flags[s] = 1 where the key (shifted right by `shift`) differs from its predecessor
template <typename K>
__global__ void __launch_bounds__(kBlock)
knn_boundary_flags_kernel(const K* __restrict__ keys, int64_t m, int shift, uint8_t* __restrict__ flags) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t s = (int64_t)blockIdx.x * kBlock + threadIdx.x; s < m; s += stride)
        flags[s] = (s == 0 || (keys[s] >> shift) != (keys[s - 1] >> shift)) ? 1 : 0;
}

template <typename K>
struct KnnCellFunctor {
    const K* keys;
    uint64_t* ucell;
    uint32_t* cell_first;
    __device__ void operator()(int64_t s, bool is_first, uint32_t rank) const {
        if (is_first) { ucell[rank] = (uint64_t)keys[s]; cell_first[rank] = (uint32_t)s; }
    }
};

struct KnnSuperFunctor {     // super-cell run starts + super-cell index per sorted position
    uint32_t* sc_first;
    uint32_t* sc_of;
    __device__ void operator()(int64_t s, bool is_first, uint32_t rank) const {
        if (is_first) sc_first[rank] = (uint32_t)s;
        sc_of[s] = rank + (is_first ? 1u : 0u) - 1u;
    }
};

__global__ void __launch_bounds__(kBlock)
knn_group_flags_kernel(const uint32_t* __restrict__ sc_first, const uint32_t* __restrict__ sc_of,
                       int64_t m, uint8_t* __restrict__ flags) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t s = (int64_t)blockIdx.x * kBlock + threadIdx.x; s < m; s += stride)
        flags[s] = (((uint32_t)s - sc_first[sc_of[s]]) & 31u) == 0 ? 1 : 0;
}

struct KnnGroupFunctor {
    uint32_t* group_start;
    __device__ void operator()(int64_t s, bool is_first, uint32_t rank) const {
        if (is_first) group_start[rank] = (uint32_t)s;
    }
};

__global__ void knn_finish_kernel(const uint32_t* __restrict__ n_cells, uint32_t* __restrict__ cell_first,
                                  const uint32_t* __restrict__ n_groups, uint32_t* __restrict__ group_start,
                                  uint32_t m) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { cell_first[*n_cells] = m; group_start[*n_groups] = m; }
}

struct KnnScCellFunctor {    // super-cell runs over the fine-cell list
    const uint64_t* ucell;
    uint64_t* sc_key;
    uint32_t* sc_cell_first;
    __device__ void operator()(int64_t e, bool is_first, uint32_t rank) const {
        if (is_first) { sc_key[rank] = ucell[e] >> 6; sc_cell_first[rank] = (uint32_t)e; }
    }
};

__global__ void knn_finish_sc_kernel(const uint32_t* __restrict__ n_sc, uint32_t* __restrict__ sc_cell_first,
                                     uint32_t c_total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) sc_cell_first[*n_sc] = c_total;
}

__global__ void __launch_bounds__(kBlock)
knn_hash_insert_kernel(const uint64_t* __restrict__ ucell, const uint32_t* __restrict__ n_cells,
                       uint32_t* __restrict__ table, uint32_t table_mask) {
    const uint32_t c_total = *n_cells;
    uint32_t stride = gridDim.x * kBlock;
    for (uint32_t c = blockIdx.x * kBlock + threadIdx.x; c < c_total; c += stride) {
        uint32_t h = (uint32_t)hash64(ucell[c]) & table_mask;
        while (atomicCAS(&table[h], kEmpty, c) != kEmpty) h = (h + 1) & table_mask;
    }
}

// ------------------------------------------------------------------ search ------
constexpr int kCellCap = 160;      // per-warp list of candidate fine cells (flushed when nearly full)

struct SearchArgs {
    const double* sx; const double* sy; const double* sz;   // sorted coordinates
    const uint32_t* perm;          // sorted position -> compact id
    const uint64_t* ucell; const uint32_t* cell_first;       // fine cells (Morton keys, first point)
    const uint64_t* sc_key; const uint32_t* sc_cell_first;   // super-cells (key >> 6, first fine cell)
    const uint32_t* table; uint32_t table_mask;              // hash: super-cell key -> super-cell index
    const uint32_t* group_start; const uint32_t* n_groups;
    GridParams g;
    int64_t m;
    int k_heap;                    // max(knn, max_nn) clipped to m
    int knn, max_nn;
    double radius2;
    int32_t* out_knn; double* out_d2; double* out_normals; double* out_cov; double* out_curv;
};

template <int KMAX>
struct LaneHeap {
    // max-heap over (d2, compact id); storage [slot][lane] in shared memory
    double* d2; uint32_t* pos; const uint32_t* perm; int lane; int cnt;
    __device__ __forceinline__ double& D(int i) { return d2[i * 32 + lane]; }
    __device__ __forceinline__ uint32_t& P(int i) { return pos[i * 32 + lane]; }
    // (da, pa) ranks after (db, pb) in ascending (d2, compact id) order
    __device__ __forceinline__ bool after(double da, uint32_t pa, double db, uint32_t pb) const {
        if (da != db) return da > db;
        return __ldg(perm + pa) > __ldg(perm + pb);
    }
    __device__ __forceinline__ void sift_down(int i, int n, double vd, uint32_t vp) {
        while (true) {
            int c = 2 * i + 1;
            if (c >= n) break;
            double cd = D(c); uint32_t cp = P(c);
            if (c + 1 < n) {
                double rd = D(c + 1); uint32_t rp = P(c + 1);
                if (after(rd, rp, cd, cp)) { c = c + 1; cd = rd; cp = rp; }
            }
            if (!after(cd, cp, vd, vp)) break;
            D(i) = cd; P(i) = cp;
            i = c;
        }
        D(i) = vd; P(i) = vp;
    }
    __device__ __forceinline__ void push(double vd, uint32_t vp) {   // cnt < capacity
        int i = cnt++;
        while (i > 0) {
            int p = (i - 1) >> 1;
            double pd = D(p); uint32_t pp = P(p);
            if (!after(vd, vp, pd, pp)) break;
            D(i) = pd; P(i) = pp;
            i = p;
        }
        D(i) = vd; P(i) = vp;
    }
    __device__ __forceinline__ void sort_ascending() {
        for (int end = cnt - 1; end > 0; --end) {
            double vd = D(end); uint32_t vp = P(end);
            D(end) = D(0); P(end) = P(0);
            sift_down(0, end, vd, vp);
        }
    }
};

__device__ __forceinline__ double warp_max_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// conservative squared distance between the axis-aligned box [lo, hi] and the box [bmin, bmax]
__device__ __forceinline__ double box_gap2(double lox, double loy, double loz, double hix, double hiy, double hiz,
                                           double bminx, double bminy, double bminz,
                                           double bmaxx, double bmaxy, double bmaxz, double slack) {
    double gx = fmax(0.0, fmax(lox - bmaxx, bminx - hix) - slack);
    double gy = fmax(0.0, fmax(loy - bmaxy, bminy - hiy) - slack);
    double gz = fmax(0.0, fmax(loz - bmaxz, bminz - hiz) - slack);
    return (gx * gx + gy * gy + gz * gz) * (1.0 - 1e-9);
}

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// largest k in [0, 32) with off[k] <= i  (off is a non-decreasing exclusive scan in smem)
__device__ __forceinline__ int seg_search(const uint32_t* off, uint32_t i) {
    int k = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1)
        if (off[k + step] <= i) k += step;
    return k;
}

// per-warp shared memory carve-up (bytes)
template <int KMAX> struct SearchSmem {
    static constexpr size_t heap_d2 = 0;
    static constexpr size_t heap_pos = heap_d2 + (size_t)KMAX * 32 * 8;
    static constexpr size_t c_xy = heap_pos + (size_t)KMAX * 32 * 4;       // double2[32]
    static constexpr size_t c_z = c_xy + 32 * 16;                           // double[32]
    static constexpr size_t l_lb2 = c_z + 32 * 8;                           // double[kCellCap]
    static constexpr size_t l_cell = l_lb2 + (size_t)kCellCap * 8;          // u32[kCellCap]
    static constexpr size_t c_pos = l_cell + (size_t)kCellCap * 4;          // u32[32]
    static constexpr size_t seg_a = c_pos + 32 * 4;                         // u32[64 + 32] candidate expansion
    static constexpr size_t seg_b = seg_a + 96 * 4;                         // u32[64 + 32] cell listing
    static constexpr size_t per_warp = ((seg_b + 96 * 4) + 15) / 16 * 16;
};

template <int KMAX>
__global__ void __launch_bounds__(kSearchThreads)
knn_search_kernel(SearchArgs a) {
    extern __shared__ __align__(16) unsigned char knn_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    using L = SearchSmem<KMAX>;
    unsigned char* const wbase = knn_smem + (size_t)warp * L::per_warp;
    double* const w_heap_d2 = reinterpret_cast<double*>(wbase + L::heap_d2);
    uint32_t* const w_heap_pos = reinterpret_cast<uint32_t*>(wbase + L::heap_pos);
    double2* const c_xy = reinterpret_cast<double2*>(wbase + L::c_xy);
    double* const c_z = reinterpret_cast<double*>(wbase + L::c_z);
    double* const l_lb2 = reinterpret_cast<double*>(wbase + L::l_lb2);
    uint32_t* const l_cell = reinterpret_cast<uint32_t*>(wbase + L::l_cell);
    uint32_t* const c_pos = reinterpret_cast<uint32_t*>(wbase + L::c_pos);
    uint32_t* const sa_off = reinterpret_cast<uint32_t*>(wbase + L::seg_a);      // [64]: 32 offsets + 32 x total
    uint32_t* const sa_base = sa_off + 64;
    uint32_t* const sb_off = reinterpret_cast<uint32_t*>(wbase + L::seg_b);
    uint32_t* const sb_base = sb_off + 64;

    const GridParams g = a.g;
    const uint32_t n_groups = *a.n_groups;
    const int K = a.k_heap;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double scell = 4.0 * g.cell;                       // super-cell edge
    const int nsx = (g.nx + 3) >> 2, nsy = (g.ny + 3) >> 2, nsz = (g.nz + 3) >> 2;

    for (uint32_t grp = blockIdx.x * kSearchWarps + warp; grp < n_groups; grp += gridDim.x * kSearchWarps) {
        const uint32_t gs = a.group_start[grp];
        const int gcount = (int)(a.group_start[grp + 1] - gs);
        const bool valid = lane < gcount;
        const uint32_t qpos = gs + (valid ? lane : 0);
        const double qx = a.sx[qpos], qy = a.sy[qpos], qz = a.sz[qpos];
        // every query of a group lies in the same super-cell (groups never straddle one)
        const int sx = cell_coord(qx, g.ox, g.inv_cell, g.nx) >> 2;
        const int sy = cell_coord(qy, g.oy, g.inv_cell, g.ny) >> 2;
        const int sz = cell_coord(qz, g.oz, g.inv_cell, g.nz) >> 2;
        // bounding box of the group's queries (invalid lanes alias lane 0)
        double bminx = qx, bminy = qy, bminz = qz, bmaxx = qx, bmaxy = qy, bmaxz = qz;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bminx = fmin(bminx, __shfl_xor_sync(0xffffffffu, bminx, o)); bmaxx = fmax(bmaxx, __shfl_xor_sync(0xffffffffu, bmaxx, o));
            bminy = fmin(bminy, __shfl_xor_sync(0xffffffffu, bminy, o)); bmaxy = fmax(bmaxy, __shfl_xor_sync(0xffffffffu, bmaxy, o));
            bminz = fmin(bminz, __shfl_xor_sync(0xffffffffu, bminz, o)); bmaxz = fmax(bmaxz, __shfl_xor_sync(0xffffffffu, bmaxz, o));
        }

        LaneHeap<KMAX> heap;
        heap.d2 = w_heap_d2; heap.pos = w_heap_pos; heap.perm = a.perm; heap.lane = lane; heap.cnt = 0;
        bool done = !valid;
        double worst = inf;         // k-th best squared distance once the heap is full
        int list_n = 0;             // warp-uniform

        auto consider = [&](double d2, uint32_t cp) {
            if (heap.cnt < K) {
                heap.push(d2, cp);
                if (heap.cnt == K) worst = heap.D(0);
            } else if (heap.after(heap.D(0), heap.P(0), d2, cp)) {
                heap.sift_down(0, K, d2, cp);
                worst = heap.D(0);
            }
        };

        // Stream the points of up to 32 cells (one per lane: `first`, `cnt`) past every query:
        // load-balanced expansion, 32 coalesced candidate loads per batch staged in shared
        // memory, every lane tests every staged candidate (broadcast reads).
        auto process_cells = [&](uint32_t first, uint32_t cnt) {
            const uint32_t incl = warp_incl_scan_u32(cnt, lane);
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            sa_off[lane] = incl - cnt; sa_off[32 + lane] = total; sa_base[lane] = first;
            __syncwarp();
            for (uint32_t b = 0; b < total; b += 32) {
                const uint32_t i = b + lane;
                if (i < total) {
                    const int k = seg_search(sa_off, i);
                    const uint32_t p = sa_base[k] + (i - sa_off[k]);
                    c_xy[lane] = make_double2(a.sx[p], a.sy[p]); c_z[lane] = a.sz[p]; c_pos[lane] = p;
                }
                __syncwarp();
                const int nc = (int)min(32u, total - b);
                if (!done) {
                    for (int c = 0; c < nc; c += 4) {
                        double d[4]; uint32_t pp[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int cc = c + u < nc ? c + u : c;
                            const double2 xy = c_xy[cc];
                            d[u] = dist2(xy.x, xy.y, c_z[cc], qx, qy, qz);
                            pp[u] = c_pos[cc];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (c + u < nc && d[u] <= worst) consider(d[u], pp[u]);
                    }
                }
                __syncwarp();
            }
        };

        // Process the listed fine cells in three rounds of decreasing urgency -- cells that
        // overlap the group's bounding box, cells within half the current k-th distance, and
        // whatever still passes the prune test -- with the bound refreshed for every batch of 32.
        auto flush = [&]() {
            for (int round = 0; round < 3 && list_n > 0; ++round) {
                for (int j0 = 0; j0 < list_n; j0 += 32) {
                    const int e = j0 + lane;
                    const double wmax = warp_max_f64(done ? 0.0 : worst);
                    const double thr = round == 0 ? 0.0 : (round == 1 ? 0.25 * wmax : wmax);
                    const double lb = e < list_n ? l_lb2[e] : inf;
                    const bool elig = lb < inf && lb <= thr;      // inf marks processed / absent entries
                    if (!__any_sync(0xffffffffu, elig)) continue;
                    uint32_t first = 0, cnt = 0;
                    if (elig) {
                        l_lb2[e] = inf;
                        const uint32_t cell = l_cell[e];
                        first = __ldg(a.cell_first + cell);
                        cnt = __ldg(a.cell_first + cell + 1) - first;
                    }
                    process_cells(first, cnt);
                }
            }
            list_n = 0;
        };

        for (int R = 0;; ++R) {
            // ---- enumerate the super-cells of ring R (R = 0: the group's own super-cell)
            const int w = 2 * R + 1, v = 2 * R - 1;
            const int A = w * w, B = R > 0 ? w * v : 0, C = R > 0 ? v * v : 0;
            const int total = R == 0 ? 1 : 2 * A + 2 * B + 2 * C;
            for (int t0 = 0; t0 < total; t0 += 32) {
                int t = t0 + lane;
                uint32_t f0 = 0, f1 = 0;
                // prune bound: the largest k-th distance among the unfinished lanes
                const double wmax_chunk = warp_max_f64(done ? 0.0 : worst);
                if (t < total) {
                    int cx, cy, cz;
                    if (R == 0) { cx = sx; cy = sy; cz = sz; }
                    else if (t < 2 * A) {
                        cz = t < A ? sz - R : sz + R; if (t >= A) t -= A;
                        cx = sx - R + t % w; cy = sy - R + t / w;
                    } else if (t < 2 * A + 2 * B) {
                        t -= 2 * A;
                        cy = t < B ? sy - R : sy + R; if (t >= B) t -= B;
                        cx = sx - R + t % w; cz = sz - R + 1 + t / w;
                    } else {
                        t -= 2 * A + 2 * B;
                        cx = t < C ? sx - R : sx + R; if (t >= C) t -= C;
                        cy = sy - R + 1 + t % v; cz = sz - R + 1 + t / v;
                    }
                    if (cx >= 0 && cy >= 0 && cz >= 0 && cx < nsx && cy < nsy && cz < nsz) {
                        const double lb = box_gap2(g.ox + cx * scell, g.oy + cy * scell, g.oz + cz * scell,
                                                   g.ox + (cx + 1) * scell, g.oy + (cy + 1) * scell, g.oz + (cz + 1) * scell,
                                                   bminx, bminy, bminz, bmaxx, bmaxy, bmaxz, g.slack);
                        const uint64_t key = morton3(cx, cy, cz);
                        uint32_t h = (uint32_t)hash64(key) & a.table_mask;
                        while (!(lb > wmax_chunk)) {
                            uint32_t idx = __ldg(a.table + h);
                            if (idx == kEmpty) break;
                            if (__ldg(a.sc_key + idx) == key) {
                                f0 = __ldg(a.sc_cell_first + idx); f1 = __ldg(a.sc_cell_first + idx + 1);
                                break;
                            }
                            h = (h + 1) & a.table_mask;
                        }
                    }
                }
                // ---- list the fine cells of the super-cells found by this chunk (load-balanced)
                const uint32_t fcnt = f1 - f0;
                const uint32_t fincl = warp_incl_scan_u32(fcnt, lane);
                const uint32_t ftotal = __shfl_sync(0xffffffffu, fincl, 31);
                if (ftotal == 0) continue;
                sb_off[lane] = fincl - fcnt; sb_off[32 + lane] = ftotal; sb_base[lane] = f0;
                __syncwarp();
                for (uint32_t b = 0; b < ftotal; b += 32) {
                    const uint32_t i = b + lane;
                    bool keep = false;
                    double lb = inf;
                    uint32_t e = 0;
                    if (i < ftotal) {
                        const int k = seg_search(sb_off, i);
                        e = sb_base[k] + (i - sb_off[k]);
                        const uint64_t key = __ldg(a.ucell + e);
                        const int fx = (int)compact_bits21(key), fy = (int)compact_bits21(key >> 1), fz = (int)compact_bits21(key >> 2);
                        lb = box_gap2(g.ox + fx * g.cell, g.oy + fy * g.cell, g.oz + fz * g.cell,
                                      g.ox + (fx + 1) * g.cell, g.oy + (fy + 1) * g.cell, g.oz + (fz + 1) * g.cell,
                                      bminx, bminy, bminz, bmaxx, bmaxy, bmaxz, g.slack);
                        keep = !(lb > wmax_chunk);
                    }
                    const uint32_t kb = __ballot_sync(0xffffffffu, keep);
                    if (kb) {
                        if (list_n + 32 > kCellCap) flush();
                        if (keep) {
                            const int slot = list_n + __popc(kb & ((1u << lane) - 1u));
                            l_lb2[slot] = lb; l_cell[slot] = e;
                        }
                        list_n += __popc(kb);
                        __syncwarp();
                    }
                }
            }
            flush();
            // ---- termination: distance from the query to the unexplored region
            const bool open_x0 = sx - R > 0, open_x1 = sx + R < nsx - 1;
            const bool open_y0 = sy - R > 0, open_y1 = sy + R < nsy - 1;
            const bool open_z0 = sz - R > 0, open_z1 = sz + R < nsz - 1;
            if (!done) {
                double reach = inf;
                if (open_x0) reach = fmin(reach, qx - (g.ox + (sx - R) * scell));
                if (open_x1) reach = fmin(reach, (g.ox + (sx + R + 1) * scell) - qx);
                if (open_y0) reach = fmin(reach, qy - (g.oy + (sy - R) * scell));
                if (open_y1) reach = fmin(reach, (g.oy + (sy + R + 1) * scell) - qy);
                if (open_z0) reach = fmin(reach, qz - (g.oz + (sz - R) * scell));
                if (open_z1) reach = fmin(reach, (g.oz + (sz + R + 1) * scell) - qz);
                reach -= g.slack;
                if (reach == inf) done = true;                      // whole grid explored
                else if (heap.cnt == K && reach > 0.0 && worst < reach * reach * (1.0 - 1e-9)) done = true;
            }
            if (__all_sync(0xffffffffu, done)) break;
            if (!(open_x0 || open_x1 || open_y0 || open_y1 || open_z0 || open_z1)) break;
        }

        // --- epilogue: sort, write the kNN row, covariance / normal from the sorted list
        if (valid) {
            heap.sort_ascending();
            const int cnt = heap.cnt;
            const uint32_t cid = a.perm[qpos];
            if (a.out_knn) {
                for (int j = 0; j < a.knn; ++j)
                    a.out_knn[(size_t)cid * a.knn + j] = j < cnt ? (int32_t)__ldg(a.perm + heap.P(j)) : -1;
            }
            if (a.out_d2) {
                for (int j = 0; j < a.knn; ++j)
                    a.out_d2[(size_t)cid * a.knn + j] = j < cnt ? heap.D(j) : inf;
            }
            if (a.out_normals) {
                // KDTreeSearchParamHybrid: the max_nn nearest, of those the ones with d2 < r^2
                Cumulants cum; cum.clear();
                int nn = 0;
                const int lim = cnt < a.max_nn ? cnt : a.max_nn;
                for (int j = 0; j < lim; ++j) {
                    if (!(heap.D(j) < a.radius2)) break;
                    const uint32_t p = heap.P(j);
                    cum.add(a.sx[p], a.sy[p], a.sz[p]);
                    ++nn;
                }
                Vec3 nrm = {0.0, 0.0, 1.0};
                if (nn >= 3) {
                    double cov6[6];
                    cum.finish(nn, cov6);
                    nrm = normal_from_cov(cov6, nullptr);
                }
                a.out_normals[(size_t)cid * 3 + 0] = nrm.x;
                a.out_normals[(size_t)cid * 3 + 1] = nrm.y;
                a.out_normals[(size_t)cid * 3 + 2] = nrm.z;
            }
            if (a.out_cov || a.out_curv) {
                // compute_mean_and_covariance of the k-neighbourhood (region_growing.py:69-71)
                Cumulants cum; cum.clear();
                const int kk = cnt < a.knn ? cnt : a.knn;
                for (int j = 0; j < kk; ++j) {
                    const uint32_t p = heap.P(j);
                    cum.add(a.sx[p], a.sy[p], a.sz[p]);
                }
                double cov6[6];
                cum.finish(kk > 0 ? kk : 1, cov6);
                if (a.out_cov)
                    for (int q = 0; q < 6; ++q) a.out_cov[(size_t)cid * 6 + q] = cov6[q];
                if (a.out_curv) {
                    double ev[3];
                    fast_eigen3x3(cov6, ev);
                    double sum = ev[0] + ev[1] + ev[2];
                    for (int q = 0; q < 3; ++q) a.out_curv[(size_t)cid * 3 + q] = ev[q] / sum;
                }
            }
        }
        __syncwarp();
    }
}

static inline int bits_for(int n) { int b = 0; while ((1 << b) < n) ++b; return b; }
static inline size_t hash_capacity(int64_t cells) { size_t cap = 1024; while (cap < 2 * (size_t)cells) cap <<= 1; return cap; }
template <int KMAX> constexpr size_t search_smem_bytes() { return (size_t)kSearchWarps * SearchSmem<KMAX>::per_warp; }

template <typename K>
static int knn_impl(const double* d_x, const double* d_y, const double* d_z, int64_t n,
                    const uint8_t* d_sel, int64_t m, GridParams g, int key_bits,
                    int knn, int max_nn, double radius,
                    int32_t* d_knn, double* d_knn_d2, double* d_normals, double* d_cov, double* d_curv,
                    void* d_ws, size_t ws_bytes, cudaStream_t st) {
    Workspace ws(d_ws, ws_bytes);
    uint32_t* counters = ws.take<uint32_t>(64);          // [0] m, [1] n_cells, [2] n_super, [3] n_groups
    uint32_t* sel_idx = d_sel ? ws.take<uint32_t>((size_t)n) : nullptr;
    uint32_t* prim_ws = ws.take<uint32_t>(compact_ws_words(n) + sort_ws_words(m));
    K* keys_a = ws.take<K>((size_t)m);
    K* keys_b = ws.take<K>((size_t)m);
    uint32_t* vals_a = ws.take<uint32_t>((size_t)m);
    uint32_t* vals_b = ws.take<uint32_t>((size_t)m);
    double* sx = ws.take<double>((size_t)m);
    double* sy = ws.take<double>((size_t)m);
    double* sz = ws.take<double>((size_t)m);
    uint8_t* flags = ws.take<uint8_t>((size_t)m);
    uint64_t* ucell = ws.take<uint64_t>((size_t)m);
    uint32_t* cell_first = ws.take<uint32_t>((size_t)m + 1);
    uint32_t* sc_first = ws.take<uint32_t>((size_t)m);
    uint32_t* sc_of = ws.take<uint32_t>((size_t)m);
    uint32_t* group_start = ws.take<uint32_t>((size_t)m + 1);
    uint64_t* sc_key = ws.take<uint64_t>((size_t)m);
    uint32_t* sc_cell_first = ws.take<uint32_t>((size_t)m + 1);
    uint32_t* table = ws.take<uint32_t>(hash_capacity(m));
    if (!ws.ok) UPCP_FAIL(UPCP_E_WORKSPACE, "knn: workspace too small");

    if (d_sel) {
        KnnCompactIdx f{sel_idx};
        UPCP_TRY(compact_apply(d_sel, n, counters, prim_ws, st, f));
        uint32_t hm = 0;
        UPCP_CUDA_OK(cudaMemcpyAsync(&hm, counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        UPCP_CUDA_OK(cudaStreamSynchronize(st));
        if ((int64_t)hm != m) UPCP_FAIL(UPCP_E_INVALID, "knn: m does not match the number of selected points");
    } else if (m != n) {
        UPCP_FAIL(UPCP_E_INVALID, "knn: m must equal n when no selection is given");
    }

    const int grid_m = grid_for(m, kBlock);
    knn_keys_kernel<K><<<grid_m, kBlock, 0, st>>>(d_x, d_y, d_z, sel_idx, m, g, keys_a, vals_a);
    UPCP_LAUNCH_OK();
    bool in_b = false;
    if (sizeof(K) == 4)
        UPCP_TRY(radix_sort_pairs_u32((uint32_t*)keys_a, vals_a, (uint32_t*)keys_b, vals_b, m, key_bits, prim_ws, st, &in_b));
    else
        UPCP_TRY(radix_sort_pairs_u64((uint64_t*)keys_a, vals_a, (uint64_t*)keys_b, vals_b, m, key_bits, prim_ws, st, &in_b));
    K* keys = in_b ? keys_b : keys_a;
    uint32_t* perm = in_b ? vals_b : vals_a;
    knn_gather_kernel<<<grid_m, kBlock, 0, st>>>(d_x, d_y, d_z, sel_idx, perm, m, sx, sy, sz);
    UPCP_LAUNCH_OK();

    // cell list
    uint32_t* n_cells = counters + 1;
    knn_boundary_flags_kernel<K><<<grid_m, kBlock, 0, st>>>(keys, m, 0, flags);
    UPCP_LAUNCH_OK();
    {
        KnnCellFunctor<K> f{keys, ucell, cell_first};
        UPCP_TRY(compact_apply(flags, m, n_cells, prim_ws, st, f));
    }
    // super-cells (4x4x4 cells = low 6 Morton bits) and query groups
    uint32_t* n_super = counters + 2;
    uint32_t* n_groups = counters + 3;
    knn_boundary_flags_kernel<K><<<grid_m, kBlock, 0, st>>>(keys, m, 6, flags);
    UPCP_LAUNCH_OK();
    {
        KnnSuperFunctor f{sc_first, sc_of};
        UPCP_TRY(compact_apply(flags, m, n_super, prim_ws, st, f));
    }
    knn_group_flags_kernel<<<grid_m, kBlock, 0, st>>>(sc_first, sc_of, m, flags);
    UPCP_LAUNCH_OK();
    {
        KnnGroupFunctor f{group_start};
        UPCP_TRY(compact_apply(flags, m, n_groups, prim_ws, st, f));
    }
    knn_finish_kernel<<<1, 32, 0, st>>>(n_cells, cell_first, n_groups, group_start, (uint32_t)m);
    UPCP_LAUNCH_OK();

    // counts back to the host once: number of fine cells (sizes the super-cell pass and the hash)
    uint32_t h_counts[4] = {0, 0, 0, 0};
    UPCP_CUDA_OK(cudaMemcpyAsync(h_counts, counters, sizeof(h_counts), cudaMemcpyDeviceToHost, st));
    UPCP_CUDA_OK(cudaStreamSynchronize(st));
    const uint32_t c_total = h_counts[1];
    // super-cells over the fine-cell list: contiguous runs of ucell with equal key >> 6
    uint32_t* n_sc = counters + 4;
    knn_boundary_flags_kernel<uint64_t><<<grid_for(c_total, kBlock), kBlock, 0, st>>>(ucell, c_total, 6, flags);
    UPCP_LAUNCH_OK();
    {
        KnnScCellFunctor f{ucell, sc_key, sc_cell_first};
        UPCP_TRY(compact_apply(flags, c_total, n_sc, prim_ws, st, f));
    }
    knn_finish_sc_kernel<<<1, 32, 0, st>>>(n_sc, sc_cell_first, c_total);
    UPCP_LAUNCH_OK();
    uint32_t cap = 1024;
    while (cap < 2u * c_total) cap <<= 1;            // c_total >= number of super-cells
    UPCP_CUDA_OK(cudaMemsetAsync(table, 0xff, (size_t)cap * 4, st));
    knn_hash_insert_kernel<<<grid_for(c_total, kBlock), kBlock, 0, st>>>(sc_key, n_sc, table, cap - 1);
    UPCP_LAUNCH_OK();

    SearchArgs a;
    a.sx = sx; a.sy = sy; a.sz = sz; a.perm = perm; a.ucell = ucell; a.cell_first = cell_first;
    a.sc_key = sc_key; a.sc_cell_first = sc_cell_first;
    a.table = table; a.table_mask = cap - 1; a.group_start = group_start; a.n_groups = n_groups;
    a.g = g; a.m = m;
    int kh = knn > max_nn ? knn : max_nn;
    if ((int64_t)kh > m) kh = (int)m;
    a.k_heap = kh; a.knn = knn; a.max_nn = max_nn; a.radius2 = radius * radius;
    a.out_knn = d_knn; a.out_d2 = d_knn_d2; a.out_normals = d_normals; a.out_cov = d_cov; a.out_curv = d_curv;
    const uint32_t g_total = h_counts[3];
    int blocks = (int)((g_total + kSearchWarps - 1) / kSearchWarps);
    int cap_blocks = num_sms() * 32;
    if (blocks > cap_blocks) blocks = cap_blocks;
    if (blocks < 1) blocks = 1;
    prof_begin(UPCP_PROF_KNN, st);
    if (kh <= 16) {
        UPCP_CUDA_OK(cudaFuncSetAttribute(knn_search_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)search_smem_bytes<16>()));
        knn_search_kernel<16><<<blocks, kSearchThreads, search_smem_bytes<16>(), st>>>(a);
    } else {
        UPCP_CUDA_OK(cudaFuncSetAttribute(knn_search_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)search_smem_bytes<32>()));
        knn_search_kernel<32><<<blocks, kSearchThreads, search_smem_bytes<32>(), st>>>(a);
    }
    prof_end(UPCP_PROF_KNN, st, m);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

}  // namespace upcp

using namespace upcp;

extern "C" {

size_t upcp_knn_ws_bytes(int64_t n, int64_t m, int kmax) {
    (void)kmax;
    if (n < 1) n = 1;
    if (m < 1 || m > n) m = n;
    size_t bytes = align_up(64 * 4);
    bytes += align_up((size_t)n * 4);
    bytes += align_up((compact_ws_words(n) + sort_ws_words(m)) * 4);
    bytes += 2 * align_up((size_t)m * 8) + 2 * align_up((size_t)m * 4);     // keys, vals
    bytes += 3 * align_up((size_t)m * 8);                                   // sorted coords
    bytes += align_up((size_t)m);                                           // flags
    bytes += align_up((size_t)m * 8);                                       // ucell
    bytes += 2 * align_up(((size_t)m + 1) * 4);                             // cell_first, group_start
    bytes += 2 * align_up((size_t)m * 4);                                   // sc_first, sc_of
    bytes += align_up((size_t)m * 8) + align_up(((size_t)m + 1) * 4);       // sc_key, sc_cell_first
    { size_t cap = 1024; while (cap < 2 * (size_t)m) cap <<= 1; bytes += align_up(cap * 4); }  // hash table
    return bytes + 4096;
}

int upcp_knn_normals(const double* d_x, const double* d_y, const double* d_z, int64_t n,
                     const uint8_t* d_sel, int64_t m, const double* h_bbox_sel6,
                     int knn, int max_nn, double radius, double cell_size,
                     int32_t* d_knn, double* d_knn_d2, double* d_normals, double* d_cov, double* d_curv,
                     void* d_ws, size_t ws_bytes, void* stream) {
    if (n < 0 || m < 0 || m > n || !h_bbox_sel6) UPCP_FAIL(UPCP_E_INVALID, "knn: bad argument");
    if (knn < 1 || max_nn < 0) UPCP_FAIL(UPCP_E_INVALID, "knn: knn >= 1 and max_nn >= 0 required");
    if (knn > 32 || max_nn > 32) UPCP_FAIL(UPCP_E_CAPACITY, "knn: knn and max_nn are limited to 32");
    if (!(cell_size > 0.0) || !(radius >= 0.0)) UPCP_FAIL(UPCP_E_INVALID, "knn: cell_size > 0 and radius >= 0 required");
    if (n >= ((int64_t)1 << 31)) UPCP_FAIL(UPCP_E_CAPACITY, "knn: n must be < 2^31");
    if (m == 0) return UPCP_OK;
    GridParams g;
    g.ox = h_bbox_sel6[0]; g.oy = h_bbox_sel6[1]; g.oz = h_bbox_sel6[2];
    double ex = h_bbox_sel6[3] - g.ox, ey = h_bbox_sel6[4] - g.oy, ez = h_bbox_sel6[5] - g.oz;
    if (!(ex >= 0) || !(ey >= 0) || !(ez >= 0) || !isfinite(ex + ey + ez))
        UPCP_FAIL(UPCP_E_INVALID, "knn: bounding box is not finite");
    // keep every axis within 2^20 cells (Morton keys of 3 x 21 bits)
    double emax = fmax(ex, fmax(ey, ez));
    double cell = cell_size;
    if (emax / cell > 1000000.0) cell = emax / 1000000.0;
    g.cell = cell; g.inv_cell = 1.0 / cell;
    // ... and the whole grid within 2^30 cells so that box enumeration stays in int32
    while ((floor(ex / cell) + 1) * (floor(ey / cell) + 1) * (floor(ez / cell) + 1) > 1073741824.0) cell *= 2.0;
    g.cell = cell; g.inv_cell = 1.0 / cell;
    g.nx = (int)floor(ex / cell) + 1; g.ny = (int)floor(ey / cell) + 1; g.nz = (int)floor(ez / cell) + 1;
    double amax = fmax(fmax(fabs(g.ox), fabs(h_bbox_sel6[3])), fmax(fmax(fabs(g.oy), fabs(h_bbox_sel6[4])),
                                                                     fmax(fabs(g.oz), fabs(h_bbox_sel6[5]))));
    g.slack = cell * 1e-6 + amax * 1e-12;
    int bits = bits_for(g.nx > g.ny ? (g.nx > g.nz ? g.nx : g.nz) : (g.ny > g.nz ? g.ny : g.nz));
    if (bits < 2) bits = 2;
    int key_bits = 3 * bits;
    cudaStream_t st = (cudaStream_t)stream;
    if (key_bits <= 32)
        return knn_impl<uint32_t>(d_x, d_y, d_z, n, d_sel, m, g, key_bits, knn, max_nn, radius, d_knn,
                                  d_knn_d2, d_normals, d_cov, d_curv, d_ws, ws_bytes, st);
    return knn_impl<uint64_t>(d_x, d_y, d_z, n, d_sel, m, g, key_bits, knn, max_nn, radius, d_knn,
                              d_knn_d2, d_normals, d_cov, d_curv, d_ws, ws_bytes, st);
}

}  // extern "C"
