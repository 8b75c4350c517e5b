// S4 -- exact k-nearest-neighbour graph + hybrid-search normals + curvature
// covariance over a two-level uniform grid (RegionGrowing set-up).
//
// Restates what the reference obtains from Open3D (region_growing.py:59-73,86-92,
// 105-109): KDTreeFlann.search_knn_vector_3d(p, k) for every point, estimate_normals
// with KDTreeSearchParamHybrid(radius, max_nn) and compute_mean_and_covariance of the
// k-neighbourhood.  Neighbour ordering is ascending (squared distance, compact id)
// with the squared distance computed as ((dx*dx + dy*dy) + dz*dz) in fp64, no FMA
// [ext, unverified -- see oracle/].
//
// Data structure (all in HBM, built per call):
//   * selected points get the Morton key of their fine grid cell (edge c); (key, compact
//     id) pairs are radix-sorted; coordinates are gathered into sorted SoA arrays so that
//     every cell is a contiguous, coalesced run;
//   * fine-cell list (unique keys + first-point offsets); super-cells (4x4x4 fine cells =
//     key >> 6) are contiguous runs of the fine-cell list; an open-addressing hash table
//     (4 B slots, L2-resident) maps a super-cell key to its run.
// Search (knn_select_kernel): ONE WARP PER QUERY, one lane per candidate.
//   * the warp walks super-cell rings around the query; each lane looks one super-cell up,
//     the fine cells of the hits are listed (load-balanced) with their exact lower-bound
//     distance to the query and pruned against the current k-th distance;
//   * listed cells are processed nearest-first; 32 candidates are loaded coalesced per
//     step, every lane computes one squared distance;
//   * the k best are a warp-distributed sorted list in REGISTERS (lane j holds the j-th
//     neighbour): batches with many hits are merged with a bitonic sort + bitonic merge,
//     sparse hits are inserted with a ballot + shuffle shift.  No shared-memory heap, no
//     divergence, occupancy bounded by registers only;
//   * a query is finished when its k-th distance is smaller than its distance to the
//     unexplored region (strict, with a conservative slack) -- so the result is exact.
// Epilogue (knn_epilogue_kernel): one thread per query folds its sorted neighbour list in
// order into the Open3D cumulant covariances / FastEigen3x3 normal and writes the kNN row.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"
#include "primitives.cuh"
#include "upcp_math.cuh"

namespace upcp {

constexpr int kBlock = 256;
constexpr int kSelWarps = 8;
constexpr int kSelThreads = kSelWarps * 32;
constexpr int kCellCap = 128;           // per-warp list of candidate fine cells
constexpr int kRow = 48;                // neighbour-candidate row stride (u32 sorted positions)
constexpr int kGrpWarps = 4;
constexpr int kGrpThreads = kGrpWarps * 32;
constexpr int kBuckets = 64;            // radial histogram resolution of the group kernel
constexpr int kEpiThreads = 128;
constexpr uint32_t kEmpty = 0xffffffffu;
constexpr uint32_t kFull = 0xffffffffu;

struct GridParams {
    double ox, oy, oz, cell, inv_cell, slack;
    int nx, ny, nz;
    int sc_log;        // super-cell = (1 << sc_log)^3 fine cells (2 or 3)
};

// Sorted points as 32-byte records (x, y, z, compact id): one sector per gathered neighbour.
struct __align__(32) P4 { double x, y, z, w; };
__device__ __forceinline__ P4 load_p4(const P4* __restrict__ p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    P4 r; r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y;
    return r;
}
__device__ __forceinline__ uint32_t p4_cid(const P4& v) { return (uint32_t)__double_as_longlong(v.w); }

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

// ------------------------------------------------------------------ build ------
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
                  const uint32_t* __restrict__ perm, int64_t m, P4* __restrict__ spts) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t s = (int64_t)blockIdx.x * kBlock + threadIdx.x; s < m; s += stride) {
        uint32_t j = perm[s];
        int64_t i = sel_idx ? (int64_t)sel_idx[j] : (int64_t)j;
        double2* o = reinterpret_cast<double2*>(spts + s);
        o[0] = make_double2(x[i], y[i]);
        o[1] = make_double2(z[i], __longlong_as_double((long long)j));
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

struct KnnScCellFunctor {    // super-cell runs over the fine-cell list
    const uint64_t* ucell;
    uint64_t* sc_key;
    uint32_t* sc_cell_first;
    int sc_shift;
    __device__ void operator()(int64_t e, bool is_first, uint32_t rank) const {
        if (is_first) { sc_key[rank] = ucell[e] >> sc_shift; sc_cell_first[rank] = (uint32_t)e; }
    }
};

__global__ void knn_finish_cells_kernel(const uint32_t* __restrict__ n_cells, uint32_t* __restrict__ cell_first,
                                        uint32_t m) {
    if (threadIdx.x == 0 && blockIdx.x == 0) cell_first[*n_cells] = m;
}

__global__ void knn_finish_sc_kernel(const uint32_t* __restrict__ n_sc, uint32_t* __restrict__ sc_cell_first,
                                     uint32_t c_total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) sc_cell_first[*n_sc] = c_total;
}

__global__ void __launch_bounds__(kBlock)
knn_hash_insert_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ n_keys,
                       uint32_t* __restrict__ table, uint32_t table_mask) {
    const uint32_t total = *n_keys;
    uint32_t stride = gridDim.x * kBlock;
    for (uint32_t c = blockIdx.x * kBlock + threadIdx.x; c < total; c += stride) {
        uint32_t h = (uint32_t)hash64(keys[c]) & table_mask;
        while (atomicCAS(&table[h], kEmpty, c) != kEmpty) h = (h + 1) & table_mask;
    }
}

// ------------------------------------------------------------------ search ------
struct SelectArgs {
    const P4* spts;                // sorted points (x, y, z, compact id)
    const uint64_t* ucell; const uint32_t* cell_first;       // fine cells (Morton keys, first point)
    const uint64_t* sc_key; const uint32_t* sc_cell_first;   // super-cells (key >> 6, first fine cell)
    const uint32_t* table; uint32_t table_mask;              // hash: super-cell key -> super-cell index
    GridParams g;
    int64_t m;
    int k;                         // neighbours kept per query (<= 32), clipped to m
    const uint32_t* todo_idx;      // queries (sorted positions) to process, or NULL = all m
    const uint32_t* todo_cnt;      // device count of todo_idx
    uint32_t* out_pos;             // [kRow][m] sorted positions of the k nearest, ascending
    uint8_t* out_cnt;              // [m] number of valid entries of the row
};

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(kFull, v, o);
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

// conservative squared distance between the point q and the axis-aligned box [lo, hi]
__device__ __forceinline__ double point_box_gap2(double qx, double qy, double qz,
                                                 double lox, double loy, double loz,
                                                 double hix, double hiy, double hiz, double slack) {
    double gx = fmax(0.0, fmax(lox - qx, qx - hix) - slack);
    double gy = fmax(0.0, fmax(loy - qy, qy - hiy) - slack);
    double gz = fmax(0.0, fmax(loz - qz, qz - hiz) - slack);
    return (gx * gx + gy * gy + gz * gz) * (1.0 - 1e-9);
}

// conservative squared distance between two axis-aligned boxes
__device__ __forceinline__ double box_box_gap2(double lox, double loy, double loz, double hix, double hiy, double hiz,
                                               double bminx, double bminy, double bminz,
                                               double bmaxx, double bmaxy, double bmaxz, double slack) {
    double gx = fmax(0.0, fmax(lox - bmaxx, bminx - hix) - slack);
    double gy = fmax(0.0, fmax(loy - bmaxy, bminy - hiy) - slack);
    double gz = fmax(0.0, fmax(loz - bmaxz, bminz - hiz) - slack);
    return (gx * gx + gy * gy + gz * gz) * (1.0 - 1e-9);
}

// One entry of the warp-distributed neighbour list: (squared distance, compact id) is the
// sort key, p the sorted position of the point.
struct Nb {
    double d; uint32_t c; uint32_t p;
};
__device__ __forceinline__ bool nb_less(double da, uint32_t ca, double db, uint32_t cb) {
    return da < db || (da == db && ca < cb);
}
__device__ __forceinline__ Nb nb_shfl_xor(const Nb& v, int j) {
    Nb r;
    r.d = __shfl_xor_sync(kFull, v.d, j); r.c = __shfl_xor_sync(kFull, v.c, j); r.p = __shfl_xor_sync(kFull, v.p, j);
    return r;
}
__device__ __forceinline__ Nb nb_shfl(const Nb& v, int src) {
    Nb r;
    r.d = __shfl_sync(kFull, v.d, src); r.c = __shfl_sync(kFull, v.c, src); r.p = __shfl_sync(kFull, v.p, src);
    return r;
}
// compare-exchange with lane ^ j; keep the smaller element when `keep_min`
__device__ __forceinline__ void nb_cx(Nb& v, int j, bool keep_min) {
    const Nb o = nb_shfl_xor(v, j);
    const bool o_less = nb_less(o.d, o.c, v.d, v.c);
    if (o_less == keep_min) v = o;
}
// full ascending bitonic sort of 32 entries (one per lane)
__device__ __forceinline__ void nb_sort32(Nb& v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool asc = (lane & k) == 0 || k == 32;
            const bool lower = (lane & j) == 0;
            nb_cx(v, j, lower == asc);
        }
    }
}
// ascending sort of a bitonic sequence of 32 entries
__device__ __forceinline__ void nb_merge32(Nb& v, int lane) {
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) nb_cx(v, j, (lane & j) == 0);
}

__global__ void __launch_bounds__(kSelThreads)
knn_select_kernel(SelectArgs a) {
    // per-warp shared memory: candidate-cell list + two load-balancing scratch areas
    __shared__ double s_lb2[kSelWarps][kCellCap];
    __shared__ uint32_t s_cell[kSelWarps][kCellCap];
    __shared__ uint32_t s_seg[kSelWarps][2][96];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* const l_lb2 = s_lb2[warp];
    uint32_t* const l_cell = s_cell[warp];
    uint32_t* const sa_off = s_seg[warp][0]; uint32_t* const sa_base = sa_off + 64;
    uint32_t* const sb_off = s_seg[warp][1]; uint32_t* const sb_base = sb_off + 64;

    const GridParams g = a.g;
    const int K = a.k;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const int scl = g.sc_log, scm = (1 << g.sc_log) - 1;
    const double scell = (double)(1 << g.sc_log) * g.cell;   // super-cell edge
    const int nsx = (g.nx + scm) >> scl, nsy = (g.ny + scm) >> scl, nsz = (g.nz + scm) >> scl;

    const int64_t n_todo = a.todo_idx ? (int64_t)*a.todo_cnt : a.m;
    for (int64_t qi = (int64_t)blockIdx.x * kSelWarps + warp; qi < n_todo; qi += (int64_t)gridDim.x * kSelWarps) {
        const int64_t q = a.todo_idx ? (int64_t)a.todo_idx[qi] : qi;
        const P4 qp = load_p4(a.spts + q);
        const double qx = qp.x, qy = qp.y, qz = qp.z;
        const int sx = cell_coord(qx, g.ox, g.inv_cell, g.nx) >> scl;
        const int sy = cell_coord(qy, g.oy, g.inv_cell, g.ny) >> scl;
        const int sz = cell_coord(qz, g.oz, g.inv_cell, g.nz) >> scl;

        Nb mine; mine.d = inf; mine.c = kFull; mine.p = 0;   // lane j: j-th nearest so far
        double kth_d = inf; uint32_t kth_c = kFull;          // the K-th entry (lane K-1), warp-uniform
        int list_n = 0;                                      // warp-uniform

        auto refresh_kth = [&]() {
            kth_d = __shfl_sync(kFull, mine.d, K - 1);
            kth_c = __shfl_sync(kFull, mine.c, K - 1);
        };

        // Stream the points of up to 32 cells (one per lane: first, cnt) past the query.
        auto process_cells = [&](uint32_t first, uint32_t cnt) {
            const uint32_t incl = warp_incl_scan_u32(cnt, lane);
            const uint32_t total = __shfl_sync(kFull, incl, 31);
            sa_off[lane] = incl - cnt; sa_off[32 + lane] = total; sa_base[lane] = first;
            __syncwarp();
            for (uint32_t b = 0; b < total; b += 32) {
                const uint32_t i = b + lane;
                Nb cand; cand.d = inf; cand.c = kFull; cand.p = 0;
                if (i < total) {
                    const int kk = seg_search(sa_off, i);
                    const uint32_t p = sa_base[kk] + (i - sa_off[kk]);
                    const P4 cp = load_p4(a.spts + p);
                    cand.d = dist2(cp.x, cp.y, cp.z, qx, qy, qz);
                    cand.c = p4_cid(cp);
                    cand.p = p;
                }
                bool pass = nb_less(cand.d, cand.c, kth_d, kth_c);
                uint32_t hits = __ballot_sync(kFull, pass);
                if (hits == 0) continue;
                if (__popc(hits) > 8) {
                    // batch merge: sort the 32 candidates, keep the 32 smallest of list + batch
                    if (!pass) { cand.d = inf; cand.c = kFull; }
                    nb_sort32(cand, lane);
                    const Nb rev = nb_shfl(cand, 31 - lane);
                    if (nb_less(rev.d, rev.c, mine.d, mine.c)) mine = rev;
                    nb_merge32(mine, lane);
                    if (lane >= K) { mine.d = inf; mine.c = kFull; }
                    refresh_kth();
                } else {
                    // sparse hits: insert one by one (ballot + shuffle shift)
                    while (hits) {
                        const int src = __ffs(hits) - 1;
                        hits &= hits - 1;
                        const Nb nw = nb_shfl(cand, src);
                        if (!nb_less(nw.d, nw.c, kth_d, kth_c)) continue;        // k-th moved meanwhile
                        const bool before = nb_less(mine.d, mine.c, nw.d, nw.c);   // my entry stays in front
                        const int at = __popc(__ballot_sync(kFull, before));       // insertion rank
                        Nb up;
                        up.d = __shfl_up_sync(kFull, mine.d, 1); up.c = __shfl_up_sync(kFull, mine.c, 1);
                        up.p = __shfl_up_sync(kFull, mine.p, 1);
                        if (lane == at) mine = nw; else if (lane > at) mine = up;
                        if (lane >= K) { mine.d = inf; mine.c = kFull; }
                        refresh_kth();
                    }
                }
            }
            __syncwarp();
        };

        // Process the listed fine cells nearest-first in three rounds (cells containing the
        // query, cells within a quarter of the current k-th squared distance, the rest), each
        // batch re-checked against the current k-th distance.
        auto flush = [&]() {
            for (int round = 0; round < 3 && list_n > 0; ++round) {
                for (int j0 = 0; j0 < list_n; j0 += 32) {
                    const int e = j0 + lane;
                    const double thr = round == 0 ? 0.0 : (round == 1 ? 0.25 * kth_d : kth_d);
                    const double lb = e < list_n ? l_lb2[e] : inf;
                    const bool elig = lb < inf && lb <= thr;      // inf marks processed / absent entries
                    if (!__any_sync(kFull, elig)) continue;
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
            // ---- enumerate the super-cells of ring R (R = 0: the query's own super-cell)
            const int w = 2 * R + 1, v = 2 * R - 1;
            const int A = w * w, B = R > 0 ? w * v : 0, C = R > 0 ? v * v : 0;
            const int total = R == 0 ? 1 : 2 * A + 2 * B + 2 * C;
            for (int t0 = 0; t0 < total; t0 += 32) {
                int t = t0 + lane;
                uint32_t f0 = 0, f1 = 0;
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
                        const double lb = point_box_gap2(qx, qy, qz, g.ox + cx * scell, g.oy + cy * scell, g.oz + cz * scell,
                                                         g.ox + (cx + 1) * scell, g.oy + (cy + 1) * scell,
                                                         g.oz + (cz + 1) * scell, g.slack);
                        if (!(lb > kth_d)) {
                            const uint64_t key = morton3(cx, cy, cz);
                            uint32_t h = (uint32_t)hash64(key) & a.table_mask;
                            while (true) {
                                const uint32_t idx = __ldg(a.table + h);
                                if (idx == kEmpty) break;
                                if (__ldg(a.sc_key + idx) == key) {
                                    f0 = __ldg(a.sc_cell_first + idx); f1 = __ldg(a.sc_cell_first + idx + 1);
                                    break;
                                }
                                h = (h + 1) & a.table_mask;
                            }
                        }
                    }
                }
                // ---- list the fine cells of the super-cells found by this chunk (load-balanced)
                const uint32_t fcnt = f1 - f0;
                const uint32_t fincl = warp_incl_scan_u32(fcnt, lane);
                const uint32_t ftotal = __shfl_sync(kFull, fincl, 31);
                if (ftotal == 0) continue;
                sb_off[lane] = fincl - fcnt; sb_off[32 + lane] = ftotal; sb_base[lane] = f0;
                __syncwarp();
                for (uint32_t b = 0; b < ftotal; b += 32) {
                    const uint32_t i = b + lane;
                    bool keep = false;
                    double lb = inf;
                    uint32_t e = 0;
                    if (i < ftotal) {
                        const int kk = seg_search(sb_off, i);
                        e = sb_base[kk] + (i - sb_off[kk]);
                        const uint64_t key = __ldg(a.ucell + e);
                        const int fx = (int)compact_bits21(key), fy = (int)compact_bits21(key >> 1), fz = (int)compact_bits21(key >> 2);
                        lb = point_box_gap2(qx, qy, qz, g.ox + fx * g.cell, g.oy + fy * g.cell, g.oz + fz * g.cell,
                                            g.ox + (fx + 1) * g.cell, g.oy + (fy + 1) * g.cell, g.oz + (fz + 1) * g.cell,
                                            g.slack);
                        keep = !(lb > kth_d);
                    }
                    const uint32_t kb = __ballot_sync(kFull, keep);
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
                __syncwarp();
            }
            flush();
            // ---- termination: distance from the query to the unexplored region (warp-uniform)
            const bool open_x0 = sx - R > 0, open_x1 = sx + R < nsx - 1;
            const bool open_y0 = sy - R > 0, open_y1 = sy + R < nsy - 1;
            const bool open_z0 = sz - R > 0, open_z1 = sz + R < nsz - 1;
            if (!(open_x0 || open_x1 || open_y0 || open_y1 || open_z0 || open_z1)) break;   // whole grid explored
            double reach = inf;
            if (open_x0) reach = fmin(reach, qx - (g.ox + (sx - R) * scell));
            if (open_x1) reach = fmin(reach, (g.ox + (sx + R + 1) * scell) - qx);
            if (open_y0) reach = fmin(reach, qy - (g.oy + (sy - R) * scell));
            if (open_y1) reach = fmin(reach, (g.oy + (sy + R + 1) * scell) - qy);
            if (open_z0) reach = fmin(reach, qz - (g.oz + (sz - R) * scell));
            if (open_z1) reach = fmin(reach, (g.oz + (sz + R + 1) * scell) - qz);
            reach -= g.slack;
            if (kth_d < inf && reach > 0.0 && kth_d < reach * reach * (1.0 - 1e-9)) break;
        }
        // sorted positions of the neighbours, ascending; one coalesced 128-byte row
        const bool have = lane < K && mine.d < inf;
        a.out_pos[(size_t)lane * a.m + q] = have ? mine.p : kFull;
        const uint32_t hv = __ballot_sync(kFull, have);
        if (lane == 0) a.out_cnt[q] = (uint8_t)__popc(hv);
    }
}

// ------------------------------------------------------------------ group kernel --
// Dense bulk: ONE WARP PER GROUP of <= 32 Morton-consecutive queries of one super-cell, one
// lane per query; every staged candidate is tested by all lanes (broadcast reads), so
// candidate loads and the cell walk are amortised over the group.  Selection is a two-pass
// radial histogram ("radix select") instead of a per-candidate priority queue:
//   bound    all members of the group lie in their bounding box, so (with >= k members) the
//            k-th neighbour of lane i is no farther than tau_i = the farthest corner of the box;
//   pass A   stream every point of the cells within tau_max of the box; each lane counts the
//            candidates with d2 <= tau_i into 64 equal-width d2 buckets (shared-memory u16);
//   select   b*_i = first bucket whose cumulative count reaches k; the bucket index is a monotone
//            function of d2, so every candidate outside buckets <= b*_i is strictly farther than
//            every candidate inside: the k nearest are among the (few more than k) inside;
//   pass B   stream the same cells again and append the inside candidates to the query's row.
// Both passes only FILTER, so they run in fp32 on coordinates relative to the box corner (full-rate
// pipe, 16-byte staged records); pass B keeps one bucket more than b*_i, a margin thousands of
// times larger than the fp32 rounding of a squared distance, so the kept set provably contains
// the k nearest under the exact fp64 ordering.  (Precisely: pass B keeps a quarter of the next
// bucket beyond b*_i; the fp32 error of a squared distance is < 1e-3 of a bucket width.)
// The row is sorted exactly by (fp64 d2, compact id) in the epilogue kernel.  Queries the scheme
// cannot serve (groups with < k members, rows that would overflow, elongated groups) are flagged
// and handled by the warp-per-query kernel above.  Both paths are exact.
struct GroupArgs {
    const P4* spts;
    const uint64_t* ucell; const uint32_t* cell_first;
    const uint64_t* sc_key; const uint32_t* sc_cell_first;
    const uint32_t* table; uint32_t table_mask;
    const uint4* groups; const uint32_t* n_groups;     // member_start, n_members, chunk_start, chunk_cnt
    GridParams g;
    int64_t m;
    int k;
    uint32_t* out_pos; uint8_t* out_cnt; uint8_t* todo;
};

__global__ void __launch_bounds__(kGrpThreads)
knn_group_kernel(GroupArgs a) {
    __shared__ uint16_t s_hist[kGrpWarps][kBuckets * 32];
    __shared__ float4 s_cf[kGrpWarps][32];          // staged candidate: box-relative x, y, z, sorted position
    __shared__ double s_lb2[kGrpWarps][kCellCap];
    __shared__ uint32_t s_cells[kGrpWarps][kCellCap];
    __shared__ uint32_t s_gseg[kGrpWarps][2][96];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t* const hist = s_hist[warp];
    float4* const c_f = s_cf[warp];
    double* const l_lb2 = s_lb2[warp];
    uint32_t* const l_cell = s_cells[warp];
    uint32_t* const sa_off = s_gseg[warp][0]; uint32_t* const sa_base = sa_off + 64;
    uint32_t* const sb_off = s_gseg[warp][1]; uint32_t* const sb_base = sb_off + 64;

    const GridParams g = a.g;
    const int K = a.k;
    const uint32_t n_groups = *a.n_groups;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const int scl = g.sc_log, scm = (1 << g.sc_log) - 1;
    const double scell = (double)(1 << g.sc_log) * g.cell;
    const int nsx = (g.nx + scm) >> scl, nsy = (g.ny + scm) >> scl, nsz = (g.nz + scm) >> scl;

    for (uint32_t grp = blockIdx.x * kGrpWarps + warp; grp < n_groups; grp += gridDim.x * kGrpWarps) {
        const uint4 gd = a.groups[grp];
        const uint32_t member_start = gd.x, n_members = gd.y, chunk_start = gd.z, chunk_cnt = gd.w;
        const bool member = (uint32_t)lane < n_members;
        const uint32_t pos = member_start + (member ? lane : 0);
        const bool is_query = member && pos >= chunk_start && pos < chunk_start + chunk_cnt;
        const P4 qp = load_p4(a.spts + pos);
        const double qx = qp.x, qy = qp.y, qz = qp.z;
        // bounding box of the members (non-member lanes alias member 0)
        double bminx = qx, bminy = qy, bminz = qz, bmaxx = qx, bmaxy = qy, bmaxz = qz;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bminx = fmin(bminx, __shfl_xor_sync(kFull, bminx, o)); bmaxx = fmax(bmaxx, __shfl_xor_sync(kFull, bmaxx, o));
            bminy = fmin(bminy, __shfl_xor_sync(kFull, bminy, o)); bmaxy = fmax(bmaxy, __shfl_xor_sync(kFull, bmaxy, o));
            bminz = fmin(bminz, __shfl_xor_sync(kFull, bminz, o)); bmaxz = fmax(bmaxz, __shfl_xor_sync(kFull, bmaxz, o));
        }
        // tau_i: squared distance to the farthest corner of the box (>= d2 to every member, with slack)
        double tau;
        {
            const double ex = fmax(qx - bminx, bmaxx - qx), ey = fmax(qy - bminy, bmaxy - qy), ez = fmax(qz - bminz, bmaxz - qz);
            tau = (ex * ex + ey * ey + ez * ez) * (1.0 + 1e-9) + 1e-300;
        }
        double tau_max = tau;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tau_max = fmax(tau_max, __shfl_xor_sync(kFull, tau_max, o));
        const float scale = (float)((double)kBuckets / tau);
        const float qfx = (float)(qx - bminx), qfy = (float)(qy - bminy), qfz = (float)(qz - bminz);
        const int sx = cell_coord(qx, g.ox, g.inv_cell, g.nx) >> scl;       // same for all members
        const int sy = cell_coord(qy, g.oy, g.inv_cell, g.ny) >> scl;
        const int sz = cell_coord(qz, g.oz, g.inv_cell, g.nz) >> scl;
        const int gsx = __shfl_sync(kFull, sx, 0), gsy = __shfl_sync(kFull, sy, 0), gsz = __shfl_sync(kFull, sz, 0);
        bool active = is_query;
        // (a degenerate box -- members within a fraction of a millimetre -- leaves no room for the
        // fp32 margin: hand it over as well)
        if ((int)ceil((sqrt(tau_max) + g.slack) / scell) > 2 || (int)n_members < K || !(tau_max > 1e-7)) {
            if (is_query) a.todo[pos] = 1;               // not served here: per-query kernel
            continue;
        }

        int bstar = -1;            // last (partially) kept bucket
        float tkeep = 0.0f;        // pass B keeps candidates with d2f * scale < tkeep
        bool overflow = false;
        double bound = tau;        // current upper bound of this lane's k-th squared distance
        double bound_max = tau_max;
        int list_n = 0;

        // first bucket whose cumulative count reaches K -> tightened bound (upper bucket edge)
        auto tighten = [&]() {
            uint32_t cum = 0;
            int b1 = -1;
#pragma unroll 1
            for (int b = 0; b < kBuckets; ++b) {
                cum += hist[b * 32 + lane];
                if (cum >= (uint32_t)K) { b1 = b; break; }
            }
            // pruning bound: upper edge of bucket b1 plus the quarter-bucket margin pass B keeps
            if (b1 >= 0) bound = fmin(bound, ((double)(b1 + 1) + 0.25) / (double)scale * (1.0 + 1e-6));
            bound_max = active ? bound : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bound_max = fmax(bound_max, __shfl_xor_sync(kFull, bound_max, o));
        };

        // stream the points of the eligible listed cells past every lane (pass 0: count into the
        // histogram, pass 1: counting-sort placement into the row)
        auto stream = [&](int pass, double lo_excl, double hi_incl) {
            for (int j0 = 0; j0 < list_n; j0 += 32) {
                const int e = j0 + lane;
                uint32_t first = 0, cnt = 0;
                if (e < list_n) {
                    const double lb = l_lb2[e];
                    if (lb > lo_excl && lb <= hi_incl) {
                        const uint32_t cell = l_cell[e];
                        first = __ldg(a.cell_first + cell);
                        cnt = __ldg(a.cell_first + cell + 1) - first;
                    }
                }
                const uint32_t incl = warp_incl_scan_u32(cnt, lane);
                const uint32_t total = __shfl_sync(kFull, incl, 31);
                if (total == 0) continue;
                sa_off[lane] = incl - cnt; sa_off[32 + lane] = total; sa_base[lane] = first;
                __syncwarp();
                for (uint32_t b = 0; b < total; b += 32) {
                    const uint32_t i = b + lane;
                    if (i < total) {
                        const int kk = seg_search(sa_off, i);
                        const uint32_t p = sa_base[kk] + (i - sa_off[kk]);
                        const P4 cp = load_p4(a.spts + p);
                        c_f[lane] = make_float4((float)(cp.x - bminx), (float)(cp.y - bminy), (float)(cp.z - bminz),
                                                __uint_as_float(p));
                    }
                    __syncwarp();
                    const int nc = (int)min(32u, total - b);
                    if (active) {
#pragma unroll 4
                        for (int c = 0; c < nc; ++c) {
                            const float4 cf = c_f[c];
                            const float dx = cf.x - qfx, dy = cf.y - qfy, dz = cf.z - qfz;
                            const float d2f = dx * dx + dy * dy + dz * dz;
                            const float tb = d2f * scale;                  // monotone in d2f
                            int bk = __float2int_rz(tb);                   // saturating
                            bk = bk > kBuckets - 1 ? kBuckets - 1 : bk;
                            if (pass == 0) {
                                hist[bk * 32 + lane] += 1;
                            } else if (tb < tkeep) {                       // bk <= bstar
                                const uint32_t slot = hist[bk * 32 + lane];
                                hist[bk * 32 + lane] = (uint16_t)(slot + 1);
                                if (slot < (uint32_t)kRow) a.out_pos[(size_t)slot * a.m + pos] = __float_as_uint(cf.w);
                                else overflow = true;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        };

        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 0) {
                for (int i = lane; i < kBuckets * 32; i += 32) hist[i] = 0;      // [bucket][lane]
                __syncwarp();
            }
            const int rmax = (int)ceil((sqrt(bound_max) + g.slack) / scell);
            // Enumeration state machine: fill the cell list until it is full or the rings are
            // exhausted, then flush it -- ONE flush / stream site keeps the kernel small enough
            // for the instruction cache.
            int R = 0, t0 = 0;
            uint32_t fb = 0, ftotal = 0;
            bool have_chunk = false, enum_done = false;
            while (!enum_done) {
                while (true) {
                    if (!have_chunk) {
                        if (R > rmax) { enum_done = true; break; }
                        const int w = 2 * R + 1, v = 2 * R - 1;
                        const int A = w * w, B = R > 0 ? w * v : 0, C = R > 0 ? v * v : 0;
                        const int total = R == 0 ? 1 : 2 * A + 2 * B + 2 * C;
                        int t = t0 + lane;
                        uint32_t f0 = 0, f1 = 0;
                        if (t < total) {
                            int cx, cy, cz;
                            if (R == 0) { cx = gsx; cy = gsy; cz = gsz; }
                            else if (t < 2 * A) {
                                cz = t < A ? gsz - R : gsz + R; if (t >= A) t -= A;
                                cx = gsx - R + t % w; cy = gsy - R + t / w;
                            } else if (t < 2 * A + 2 * B) {
                                t -= 2 * A;
                                cy = t < B ? gsy - R : gsy + R; if (t >= B) t -= B;
                                cx = gsx - R + t % w; cz = gsz - R + 1 + t / w;
                            } else {
                                t -= 2 * A + 2 * B;
                                cx = t < C ? gsx - R : gsx + R; if (t >= C) t -= C;
                                cy = gsy - R + 1 + t % v; cz = gsz - R + 1 + t / v;
                            }
                            if (cx >= 0 && cy >= 0 && cz >= 0 && cx < nsx && cy < nsy && cz < nsz) {
                                const double lb = box_box_gap2(g.ox + cx * scell, g.oy + cy * scell, g.oz + cz * scell,
                                                               g.ox + (cx + 1) * scell, g.oy + (cy + 1) * scell, g.oz + (cz + 1) * scell,
                                                               bminx, bminy, bminz, bmaxx, bmaxy, bmaxz, g.slack);
                                if (!(lb > bound_max)) {
                                    const uint64_t key = morton3(cx, cy, cz);
                                    uint32_t h = (uint32_t)hash64(key) & a.table_mask;
                                    while (true) {
                                        const uint32_t idx = __ldg(a.table + h);
                                        if (idx == kEmpty) break;
                                        if (__ldg(a.sc_key + idx) == key) {
                                            f0 = __ldg(a.sc_cell_first + idx); f1 = __ldg(a.sc_cell_first + idx + 1);
                                            break;
                                        }
                                        h = (h + 1) & a.table_mask;
                                    }
                                }
                            }
                        }
                        t0 += 32;
                        if (t0 >= total) { ++R; t0 = 0; }
                        const uint32_t fcnt = f1 - f0;
                        const uint32_t fincl = warp_incl_scan_u32(fcnt, lane);
                        ftotal = __shfl_sync(kFull, fincl, 31);
                        if (ftotal == 0) continue;
                        __syncwarp();
                        sb_off[lane] = fincl - fcnt; sb_off[32 + lane] = ftotal; sb_base[lane] = f0;
                        __syncwarp();
                        fb = 0;
                        have_chunk = true;
                    }
                    // expand the fine cells of the current chunk into the list
                    while (fb < ftotal && list_n + 32 <= kCellCap) {
                        const uint32_t i = fb + lane;
                        bool keep = false;
                        double lb = inf;
                        uint32_t e = 0;
                        if (i < ftotal) {
                            const int kk = seg_search(sb_off, i);
                            e = sb_base[kk] + (i - sb_off[kk]);
                            const uint64_t key = __ldg(a.ucell + e);
                            const int fx = (int)compact_bits21(key), fy = (int)compact_bits21(key >> 1), fz = (int)compact_bits21(key >> 2);
                            lb = box_box_gap2(g.ox + fx * g.cell, g.oy + fy * g.cell, g.oz + fz * g.cell,
                                              g.ox + (fx + 1) * g.cell, g.oy + (fy + 1) * g.cell, g.oz + (fz + 1) * g.cell,
                                              bminx, bminy, bminz, bmaxx, bmaxy, bmaxz, g.slack);
                            keep = !(lb > bound_max);
                        }
                        const uint32_t kb = __ballot_sync(kFull, keep);
                        if (keep) {
                            const int slot = list_n + __popc(kb & ((1u << lane) - 1u));
                            l_lb2[slot] = lb; l_cell[slot] = e;
                        }
                        list_n += __popc(kb);
                        fb += 32;
                    }
                    __syncwarp();
                    if (fb >= ftotal) have_chunk = false;
                    else break;                      // list full: flush, then resume this chunk
                }
                // ---- flush: pass 0 counts (cells overlapping the members' box first, which hold
                // >= k points and tighten the bound before anything farther is touched), pass 1 places
                const int n_rounds = pass == 0 ? 2 : 1;
                for (int round = 0; round < n_rounds && list_n > 0; ++round) {
                    const double lo_excl = (pass == 0 && round == 1) ? 0.0 : -1.0;
                    const double hi_incl = (pass == 0 && round == 0) ? 0.0 : bound_max;
                    stream(pass, lo_excl, hi_incl);
                    if (pass == 0) tighten();
                }
                list_n = 0;
            }
            if (pass == 0) {
                // select: first bucket whose cumulative count reaches K; turn the histogram into
                // exclusive offsets so that pass B places candidates bucket-sorted (counting sort)
                __syncwarp();
                uint32_t cum = 0;
                int bs = -1;       // last bucket kept = (first bucket whose cumulative count reaches K) + 1
#pragma unroll 1
                for (int b = 0; b < kBuckets; ++b) {
                    const uint32_t hb = hist[b * 32 + lane];
                    hist[b * 32 + lane] = (uint16_t)cum;
                    const bool reached = cum >= (uint32_t)K;      // K reached by the buckets before b
                    cum += hb;
                    if (reached) { bs = b; break; }
                }
                // bs = (first bucket whose cumulative count reaches K) + 1; pass B keeps that bucket's
                // first quarter.  bs < 0: K only reached inside the clamp bucket (or never); bs == last:
                // the margin would reach into the clamp bucket, which also holds far candidates.
                if (active && (bs < 0 || bs >= kBuckets - 1)) {
                    a.todo[pos] = 1;
                    active = false;
                }
                bstar = bs;
                tkeep = (float)bs + 0.25f;
                if (active) bound = fmin(bound, ((double)bs + 0.25) / (double)scale * (1.0 + 1e-6));
                bound_max = active ? bound : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) bound_max = fmax(bound_max, __shfl_xor_sync(kFull, bound_max, o));
                if (!__any_sync(kFull, active)) break;
                __syncwarp();
            }
        }
        if (active) {
            // kept candidates = running counter of the last bucket (its offset started at the count
            // of all buckets before it)
            const uint32_t kept = hist[bstar * 32 + lane];
            if (overflow || kept > (uint32_t)kRow) a.todo[pos] = 1;
            else a.out_cnt[pos] = (uint8_t)kept;
        }
    }
}

// ---- query groups: <= 32 consecutive sorted points of one super-cell run
__global__ void __launch_bounds__(kBlock)
knn_group_count_kernel(const uint32_t* __restrict__ run_first, const uint32_t* __restrict__ n_runs,
                       uint32_t* __restrict__ ng) {
    const uint32_t total = *n_runs;
    uint32_t stride = gridDim.x * kBlock;
    for (uint32_t r = blockIdx.x * kBlock + threadIdx.x; r < total; r += stride)
        ng[r] = (run_first[r + 1] - run_first[r] + 31) / 32;
}

__global__ void __launch_bounds__(kBlock)
knn_group_fill_kernel(const uint32_t* __restrict__ run_first, const uint32_t* __restrict__ n_runs,
                      const uint32_t* __restrict__ goff, uint4* __restrict__ groups) {
    const uint32_t total = *n_runs;
    uint32_t stride = gridDim.x * kBlock;
    for (uint32_t r = blockIdx.x * kBlock + threadIdx.x; r < total; r += stride) {
        const uint32_t s0 = run_first[r], s1 = run_first[r + 1], len = s1 - s0;
        uint32_t o = goff[r];
        for (uint32_t c = s0; c < s1; c += 32, ++o) {
            const uint32_t ccnt = s1 - c < 32 ? s1 - c : 32;
            // a short tail borrows members from the preceding chunk of the same run so that the
            // bounding-box bound still covers 32 points
            const uint32_t mstart = (ccnt == 32 || len < 32) ? c : s1 - 32;
            const uint32_t mcnt = len < 32 ? len : 32;
            groups[o] = make_uint4(mstart, mcnt, c, ccnt);
        }
    }
}

struct RunFunctor {
    uint32_t* run_first;
    __device__ void operator()(int64_t s, bool is_first, uint32_t rank) const {
        if (is_first) run_first[rank] = (uint32_t)s;
    }
};
struct TodoFunctor {
    uint32_t* todo_idx;
    __device__ void operator()(int64_t s, bool is_todo, uint32_t rank) const {
        if (is_todo) todo_idx[rank] = (uint32_t)s;
    }
};
__global__ void knn_finish_runs_kernel(const uint32_t* __restrict__ n_runs, uint32_t* __restrict__ run_first, uint32_t m) {
    if (threadIdx.x == 0 && blockIdx.x == 0) run_first[*n_runs] = m;
}

// ------------------------------------------------------------------ epilogue ----
struct EpilogueArgs {
    const P4* spts;
    const uint32_t* pos; const uint8_t* cnt;      // pos: [kRow][m] (slot-major: coalesced per query thread)
    int64_t m;
    int knn, max_nn;
    double radius2;
    int32_t* out_knn; double* out_d2; double* out_normals; double* out_cov; double* out_curv;
    double thr_curve; uint8_t* out_flag;
};

// One thread per query.  The candidate row (k exact neighbours from the per-query kernel, or
// the slightly larger bucket-ordered superset from the group kernel) is ordered by
// (fp64 d2, compact id) and folded IN ORDER into the cumulant covariances (fp summation order
// is part of the contract).
// Shared memory holds ONE 8-byte key per candidate ([slot][thread] layout, conflict-free):
// the bit pattern of the non-negative fp64 squared distance -- monotone as an unsigned integer
// -- with its 6 low mantissa bits replaced by the candidate's slot in the row.  An insertion
// sort on these integers (rows arrive nearly sorted) orders every pair of candidates whose
// distances differ above the 58th bit; runs of keys that agree in the upper 58 bits (exact
// ties, duplicates) are re-ordered by the exact (d2, compact id) comparison afterwards.
// Everything else (positions, coordinates) is re-read through L1/L2 instead of being parked
// in shared memory: 384 B instead of 768 B per query, twice the resident warps.
__device__ __forceinline__ void epi_exact(const EpilogueArgs& a, int64_t s, uint64_t key, double qx, double qy, double qz,
                                          double& d2, uint32_t& c) {
    const uint32_t p = __ldg(a.pos + (size_t)(key & 63u) * a.m + s);
    const P4 cp = load_p4(a.spts + p);
    d2 = dist2(cp.x, cp.y, cp.z, qx, qy, qz);
    c = p4_cid(cp);
}

__global__ void __launch_bounds__(kEpiThreads, 4)
knn_epilogue_kernel(EpilogueArgs a) {
    extern __shared__ __align__(16) unsigned char epi_smem[];
    uint64_t* const s_key = reinterpret_cast<uint64_t*>(epi_smem);                       // [kRow][threads]
    const int tx = threadIdx.x;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    int64_t stride = (int64_t)gridDim.x * kEpiThreads;
    for (int64_t s = (int64_t)blockIdx.x * kEpiThreads + tx; s < a.m; s += stride) {
        const uint32_t* col = a.pos + s;               // slot j of this query: col[j * m]
        const P4 qp = load_p4(a.spts + s);
        const double qx = qp.x, qy = qp.y, qz = qp.z;
        const uint32_t cid = p4_cid(qp);
        const int n = (int)a.cnt[s];
        // phase 1: candidate records (independent loads, software-pipelined by 4) -> keys
        for (int j0 = 0; j0 < n; j0 += 4) {
            uint32_t pp[4]; P4 cp[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) pp[u] = j0 + u < n ? __ldg(col + (size_t)(j0 + u) * a.m) : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) cp[u] = load_p4(a.spts + pp[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (j0 + u < n) {
                    const double d2 = dist2(cp[u].x, cp[u].y, cp[u].z, qx, qy, qz);
                    s_key[(j0 + u) * kEpiThreads + tx] =
                        ((uint64_t)__double_as_longlong(d2) & ~(uint64_t)63) | (uint64_t)(j0 + u);
                }
            }
        }
        // phase 2: insertion sort of the packed keys
        bool tie = false;
        for (int j = 1; j < n; ++j) {
            const uint64_t kj = s_key[j * kEpiThreads + tx];
            int i = j;
            while (i > 0) {
                const uint64_t kp = s_key[(i - 1) * kEpiThreads + tx];
                tie |= ((kp ^ kj) >> 6) == 0;
                if (kp <= kj) break;
                s_key[i * kEpiThreads + tx] = kp;
                --i;
            }
            if (i != j) s_key[i * kEpiThreads + tx] = kj;
        }
        if (tie) {
            // exact order inside runs of keys that agree in their upper 58 bits (rare)
            for (int j = 1; j < n; ++j) {
                const uint64_t kj = s_key[j * kEpiThreads + tx];
                if (((s_key[(j - 1) * kEpiThreads + tx] ^ kj) >> 6) != 0) continue;
                double dj; uint32_t cj;
                epi_exact(a, s, kj, qx, qy, qz, dj, cj);
                int i = j;
                while (i > 0) {
                    const uint64_t kp = s_key[(i - 1) * kEpiThreads + tx];
                    if (((kp ^ kj) >> 6) != 0) break;
                    double dp; uint32_t cp;
                    epi_exact(a, s, kp, qx, qy, qz, dp, cp);
                    if (!(dj < dp || (dj == dp && cj < cp))) break;
                    s_key[i * kEpiThreads + tx] = kp;
                    --i;
                }
                if (i != j) s_key[i * kEpiThreads + tx] = kj;
            }
        }
        Cumulants nrm_cum; nrm_cum.clear();       // hybrid search: max_nn nearest with d2 < r^2
        Cumulants cov_cum; cov_cum.clear();       // the knn-neighbourhood (region_growing.py:69-71)
        int nn = 0, kk = 0;
        bool in_radius = true;
        const int lim = a.knn > a.max_nn ? a.knn : a.max_nn;
        for (int j0 = 0; j0 < lim; j0 += 4) {
            P4 cp[4]; uint32_t pp[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                pp[u] = j < n ? __ldg(col + (size_t)(s_key[j * kEpiThreads + tx] & 63u) * a.m) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) cp[u] = load_p4(a.spts + pp[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                if (j >= lim) break;
                const bool present = j < n;
                const double px = cp[u].x, py = cp[u].y, pz = cp[u].z;
                double d2 = inf;
                uint32_t c = 0;
                if (present) { d2 = dist2(px, py, pz, qx, qy, qz); c = p4_cid(cp[u]); }
                if (j < a.knn) {
                    if (a.out_knn) a.out_knn[(size_t)cid * a.knn + j] = present ? (int32_t)c : -1;
                    if (a.out_d2) a.out_d2[(size_t)cid * a.knn + j] = d2;
                    if (present) { cov_cum.add(px, py, pz); ++kk; }
                }
                if (j < a.max_nn && present && in_radius) {
                    if (d2 < a.radius2) { nrm_cum.add(px, py, pz); ++nn; }
                    else in_radius = false;            // sorted ascending: nothing closer follows
                }
            }
        }
        if (a.out_normals) {
            Vec3 nrm = {0.0, 0.0, 1.0};
            if (nn >= 3) {
                double cov6[6];
                nrm_cum.finish(nn, cov6);
                nrm = normal_from_cov(cov6, nullptr);
            }
            a.out_normals[(size_t)cid * 3 + 0] = nrm.x;
            a.out_normals[(size_t)cid * 3 + 1] = nrm.y;
            a.out_normals[(size_t)cid * 3 + 2] = nrm.z;
        }
        if (a.out_cov || a.out_curv || a.out_flag) {
            double cov6[6];
            cov_cum.finish(kk > 0 ? kk : 1, cov6);
            if (a.out_cov)
                for (int t = 0; t < 6; ++t) a.out_cov[(size_t)cid * 6 + t] = cov6[t];
            if (a.out_curv || a.out_flag) {
                double ev[3];
                fast_eigen3x3(cov6, ev);
                const double sum = ev[0] + ev[1] + ev[2];
                const double r0 = ev[0] / sum, r1 = ev[1] / sum, r2 = ev[2] / sum;
                if (a.out_curv) {
                    a.out_curv[(size_t)cid * 3 + 0] = r0; a.out_curv[(size_t)cid * 3 + 1] = r1;
                    a.out_curv[(size_t)cid * 3 + 2] = r2;
                }
                if (a.out_flag) {
                    const double tol = 1e-9;
                    const bool fin = isfinite(r0) && isfinite(r1) && isfinite(r2);
                    const bool below = fin && r0 < a.thr_curve - tol && r1 < a.thr_curve - tol && r2 < a.thr_curve - tol;
                    const bool above = fin && r0 >= a.thr_curve + tol && r1 >= a.thr_curve + tol && r2 >= a.thr_curve + tol;
                    a.out_flag[cid] = below ? 1 : (above ? 0 : 2);
                }
            }
        }
    }
}

static inline int bits_for(int n) { int b = 0; while ((1 << b) < n) ++b; return b; }
static inline size_t hash_capacity(int64_t cells) { size_t cap = 1024; while (cap < 2 * (size_t)cells) cap <<= 1; return cap; }

template <typename K>
static int knn_impl(const double* d_x, const double* d_y, const double* d_z, int64_t n,
                    const uint8_t* d_sel, int64_t m, GridParams g, int key_bits,
                    int knn, int max_nn, double radius,
                    int32_t* d_knn, double* d_knn_d2, double* d_normals, double* d_cov, double* d_curv,
                    double thr_curve, uint8_t* d_flag, void* d_ws, size_t ws_bytes, cudaStream_t st) {
    Workspace ws(d_ws, ws_bytes);
    uint32_t* counters = ws.take<uint32_t>(64);          // [0] m, [1] n_cells, [4] n_super
    uint32_t* sel_idx = d_sel ? ws.take<uint32_t>((size_t)n) : nullptr;
    uint32_t* prim_ws = ws.take<uint32_t>(compact_ws_words(n) + sort_ws_words(m));
    K* keys_a = ws.take<K>((size_t)m);
    K* keys_b = ws.take<K>((size_t)m);
    uint32_t* vals_a = ws.take<uint32_t>((size_t)m);
    uint32_t* vals_b = ws.take<uint32_t>((size_t)m);
    P4* spts = ws.take<P4>((size_t)m);
    uint8_t* flags = ws.take<uint8_t>((size_t)m);
    uint64_t* ucell = ws.take<uint64_t>((size_t)m);
    uint32_t* cell_first = ws.take<uint32_t>((size_t)m + 1);
    uint64_t* sc_key = ws.take<uint64_t>((size_t)m);
    uint32_t* sc_cell_first = ws.take<uint32_t>((size_t)m + 1);
    uint32_t* table = ws.take<uint32_t>(hash_capacity(m));
    uint32_t* nb_pos = ws.take<uint32_t>((size_t)m * kRow);
    uint8_t* nb_cnt = ws.take<uint8_t>((size_t)m);
    uint8_t* todo = ws.take<uint8_t>((size_t)m);
    uint32_t* todo_idx = ws.take<uint32_t>((size_t)m);
    uint32_t* run_first = ws.take<uint32_t>((size_t)m + 1);
    uint32_t* run_ng = ws.take<uint32_t>((size_t)m + 1);
    uint4* groups = ws.take<uint4>((size_t)m / 32 + (size_t)m + 2);
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
    knn_gather_kernel<<<grid_m, kBlock, 0, st>>>(d_x, d_y, d_z, sel_idx, perm, m, spts);
    UPCP_LAUNCH_OK();

    // fine-cell list
    uint32_t* n_cells = counters + 1;
    knn_boundary_flags_kernel<K><<<grid_m, kBlock, 0, st>>>(keys, m, 0, flags);
    UPCP_LAUNCH_OK();
    {
        KnnCellFunctor<K> f{keys, ucell, cell_first};
        UPCP_TRY(compact_apply(flags, m, n_cells, prim_ws, st, f));
    }
    knn_finish_cells_kernel<<<1, 32, 0, st>>>(n_cells, cell_first, (uint32_t)m);
    UPCP_LAUNCH_OK();
    // super-cell runs over the sorted POINTS (query groups are cut from them)
    uint32_t* n_runs = counters + 2;
    knn_boundary_flags_kernel<K><<<grid_m, kBlock, 0, st>>>(keys, m, 3 * g.sc_log, flags);
    UPCP_LAUNCH_OK();
    {
        RunFunctor f{run_first};
        UPCP_TRY(compact_apply(flags, m, n_runs, prim_ws, st, f));
    }
    knn_finish_runs_kernel<<<1, 32, 0, st>>>(n_runs, run_first, (uint32_t)m);
    UPCP_LAUNCH_OK();
    // number of fine cells / runs back to the host once (sizes the following passes)
    uint32_t h_counts[4] = {0, 0, 0, 0};
    UPCP_CUDA_OK(cudaMemcpyAsync(h_counts, counters, sizeof(h_counts), cudaMemcpyDeviceToHost, st));
    UPCP_CUDA_OK(cudaStreamSynchronize(st));
    const uint32_t c_total = h_counts[1];
    const uint32_t r_total = h_counts[2];
    // query groups: ceil(len / 32) per run
    uint32_t* n_groups = counters + 3;
    knn_group_count_kernel<<<grid_for(r_total, kBlock), kBlock, 0, st>>>(run_first, n_runs, run_ng);
    UPCP_LAUNCH_OK();
    UPCP_TRY(exclusive_scan_u32(run_ng, run_ng, r_total, n_groups, prim_ws, st));
    knn_group_fill_kernel<<<grid_for(r_total, kBlock), kBlock, 0, st>>>(run_first, n_runs, run_ng, groups);
    UPCP_LAUNCH_OK();
    // super-cells: contiguous runs of the fine-cell list with equal key >> 6
    uint32_t* n_sc = counters + 4;
    knn_boundary_flags_kernel<uint64_t><<<grid_for(c_total, kBlock), kBlock, 0, st>>>(ucell, c_total, 3 * g.sc_log, flags);
    UPCP_LAUNCH_OK();
    {
        KnnScCellFunctor f{ucell, sc_key, sc_cell_first, 3 * g.sc_log};
        UPCP_TRY(compact_apply(flags, c_total, n_sc, prim_ws, st, f));
    }
    knn_finish_sc_kernel<<<1, 32, 0, st>>>(n_sc, sc_cell_first, c_total);
    UPCP_LAUNCH_OK();
    uint32_t cap = 1024;
    while (cap < 2u * c_total) cap <<= 1;            // c_total >= number of super-cells
    UPCP_CUDA_OK(cudaMemsetAsync(table, 0xff, (size_t)cap * 4, st));
    knn_hash_insert_kernel<<<grid_for(c_total, kBlock), kBlock, 0, st>>>(sc_key, n_sc, table, cap - 1);
    UPCP_LAUNCH_OK();

    int kh = knn > max_nn ? knn : max_nn;
    if ((int64_t)kh > m) kh = (int)m;
    UPCP_CUDA_OK(cudaMemsetAsync(todo, 0, (size_t)m, st));
    prof_begin(UPCP_PROF_KNN, st);
    // dense bulk: warp per group, lane per query, two-pass radial histogram selection
    GroupArgs ga;
    ga.spts = spts; ga.ucell = ucell; ga.cell_first = cell_first;
    ga.sc_key = sc_key; ga.sc_cell_first = sc_cell_first; ga.table = table; ga.table_mask = cap - 1;
    ga.groups = groups; ga.n_groups = n_groups; ga.g = g; ga.m = m; ga.k = kh;
    ga.out_pos = nb_pos; ga.out_cnt = nb_cnt; ga.todo = todo;
    {
        int64_t upper = (int64_t)m / 32 + r_total + 1;            // >= number of groups
        int64_t need = (upper + kGrpWarps - 1) / kGrpWarps;
        int64_t cap_blocks = (int64_t)num_sms() * 32;
        knn_group_kernel<<<(int)(need < cap_blocks ? need : cap_blocks), kGrpThreads, 0, st>>>(ga);
        UPCP_LAUNCH_OK();
    }
    // the rest (sparse / short / overflowing queries): warp per query
    uint32_t* n_todo = counters + 5;
    {
        TodoFunctor f{todo_idx};
        UPCP_TRY(compact_apply(todo, m, n_todo, prim_ws, st, f));
    }
    SelectArgs a;
    a.spts = spts; a.ucell = ucell; a.cell_first = cell_first;
    a.sc_key = sc_key; a.sc_cell_first = sc_cell_first; a.table = table; a.table_mask = cap - 1;
    a.g = g; a.m = m; a.k = kh; a.todo_idx = todo_idx; a.todo_cnt = n_todo; a.out_pos = nb_pos; a.out_cnt = nb_cnt;
    knn_select_kernel<<<num_sms() * 16, kSelThreads, 0, st>>>(a);
    prof_end(UPCP_PROF_KNN, st, m);
    UPCP_LAUNCH_OK();
    if (getenv("UPCP_KNN_DEBUG")) {       // development aid: sizes of the search structures
        uint32_t h[8];
        cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        fprintf(stderr, "[upcp knn] m=%lld cells=%u runs=%u groups=%u super=%u todo=%u cell=%.4f\n", (long long)m, h[1], h[2],
                h[3], h[4], h[5], g.cell);
    }

    EpilogueArgs e;
    e.spts = spts; e.pos = nb_pos; e.cnt = nb_cnt; e.m = m;
    e.knn = knn; e.max_nn = max_nn; e.radius2 = radius * radius;
    e.out_knn = d_knn; e.out_d2 = d_knn_d2; e.out_normals = d_normals; e.out_cov = d_cov; e.out_curv = d_curv;
    e.thr_curve = thr_curve; e.out_flag = d_flag;
    const size_t epi_smem = (size_t)kRow * kEpiThreads * 8;
    UPCP_CUDA_OK(cudaFuncSetAttribute(knn_epilogue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)epi_smem));
    knn_epilogue_kernel<<<grid_for(m, kEpiThreads, 2), kEpiThreads, epi_smem, st>>>(e);
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
    bytes += align_up((size_t)m * 32);                                      // sorted points (x, y, z, id)
    bytes += align_up((size_t)m);                                           // flags
    bytes += align_up((size_t)m * 8);                                       // ucell
    bytes += align_up(((size_t)m + 1) * 4);                                 // cell_first
    bytes += align_up((size_t)m * 8) + align_up(((size_t)m + 1) * 4);       // sc_key, sc_cell_first
    { size_t cap = 1024; while (cap < 2 * (size_t)m) cap <<= 1; bytes += align_up(cap * 4); }  // hash table
    bytes += align_up((size_t)m * kRow * 4) + 2 * align_up((size_t)m);      // neighbour rows, counts, todo flags
    bytes += 3 * align_up(((size_t)m + 1) * 4);                             // todo_idx, run_first, run_ng
    bytes += align_up(((size_t)m / 32 + (size_t)m + 2) * 16);               // groups
    return bytes + 4096;
}

int upcp_knn_normals(const double* d_x, const double* d_y, const double* d_z, int64_t n,
                     const uint8_t* d_sel, int64_t m, const double* h_bbox_sel6,
                     int knn, int max_nn, double radius, double cell_size,
                     int32_t* d_knn, double* d_knn_d2, double* d_normals, double* d_cov, double* d_curv,
                     double threshold_curve, uint8_t* d_seed_flag, void* d_ws, size_t ws_bytes, void* stream) {
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
    // ... and the whole grid within 2^30 cells
    while ((floor(ex / cell) + 1) * (floor(ey / cell) + 1) * (floor(ez / cell) + 1) > 1073741824.0) cell *= 2.0;
    g.cell = cell; g.inv_cell = 1.0 / cell;
    g.nx = (int)floor(ex / cell) + 1; g.ny = (int)floor(ey / cell) + 1; g.nz = (int)floor(ez / cell) + 1;
    double amax = fmax(fmax(fabs(g.ox), fabs(h_bbox_sel6[3])), fmax(fmax(fabs(g.oy), fabs(h_bbox_sel6[4])),
                                                                     fmax(fabs(g.oz), fabs(h_bbox_sel6[5]))));
    g.slack = cell * 1e-6 + amax * 1e-12;
    g.sc_log = 2;
    if (const char* env = getenv("UPCP_KNN_SCLOG")) { int v = atoi(env); if (v == 2 || v == 3) g.sc_log = v; }   // tuning knob
    int bits = bits_for(g.nx > g.ny ? (g.nx > g.nz ? g.nx : g.nz) : (g.ny > g.nz ? g.ny : g.nz));
    if (bits < 2) bits = 2;
    int key_bits = 3 * bits;
    cudaStream_t st = (cudaStream_t)stream;
    if (key_bits <= 32)
        return knn_impl<uint32_t>(d_x, d_y, d_z, n, d_sel, m, g, key_bits, knn, max_nn, radius, d_knn,
                                  d_knn_d2, d_normals, d_cov, d_curv, threshold_curve, d_seed_flag, d_ws, ws_bytes, st);
    return knn_impl<uint64_t>(d_x, d_y, d_z, n, d_sel, m, g, key_bits, knn, max_nn, radius, d_knn,
                              d_knn_d2, d_normals, d_cov, d_curv, threshold_curve, d_seed_flag, d_ws, ws_bytes, st);
}

}  // extern "C"
