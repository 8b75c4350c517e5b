// S4 -- exact k-nearest-neighbour graph + hybrid-search normals + curvature
// covariance over a dense uniform grid (RegionGrowing set-up).
//
// Restates what the reference obtains from Open3D (region_growing.py:59-73,86-92,
// 105-109): KDTreeFlann.search_knn_vector_3d(p, k) for every point, estimate_normals
// with KDTreeSearchParamHybrid(radius, max_nn) and compute_mean_and_covariance of the
// k-neighbourhood.  Neighbour ordering is ascending (squared distance, compact id)
// with the squared distance computed as ((dx*dx + dy*dy) + dz*dz) in fp64, no FMA
// [ext, unverified -- see oracle/].
//
// Data structure (all in HBM, built per call):
//   * a DENSE uniform grid over the bounding box of the selected points, cells numbered
//     row-major (x fastest).  Selected points are counting-sorted by cell number into sorted
//     records, so every cell -- and every run of x-adjacent cells -- is one contiguous,
//     coalesced range of points;
//   * cell_first[G + 1]: exclusive prefix of the per-cell point counts (empty cells
//     included), so the points of cells [x0, x1] of a grid row are
//     cell_first[row + x0] .. cell_first[row + x1 + 1]: TWO loads per row, no hashing, no
//     tree.  With 180 GB of HBM the table (4 B per cell, a few cells per point) is cheap.
// What has to be found per query is what the consumers read: the `knn` nearest (kNN row, curvature
// covariance) and those of the `max_nn` nearest that lie within `radius` (hybrid-search normal).
// Beyond BOTH the knn-th distance and the radius nothing matters, so every tier certifies
// "the max_nn nearest are known" OR "the knn nearest and everything within the radius are known".
// Search, three exact tiers:
//   1. knn_block_kernel<1>: one LANE per query, candidates = the query's own 3x3x3 cell block
//      (9 contiguous row ranges).  Selection is a two-pass radial histogram: pass A counts the
//      candidates into 32 equal-width buckets of squared distance over [0, tau), tau <= reach^2,
//      reach = the distance from the query to the nearest face of its block (everything outside
//      the block is farther); the cumulative counts give the bucket below which everything needed
//      lies; pass B appends the candidates below it (plus a quarter bucket of margin) to the
//      query's row.  Both passes only FILTER, so they run in fp32 on origin-relative coordinates;
//      the margin is at least two rigorous fp32 error bounds (checked per query), so the kept set
//      provably contains what is needed under the exact fp64 ordering.  If the needed candidates
//      do not end below the last bucket the block cannot certify the answer and the query moves on to
//   2. knn_block_kernel<2>: the same over the 5x5x5 block (25 row ranges), then
//   3. knn_select_kernel: one WARP per query, thick shells of grid rows (on a 2x coarser second
//      grid when this tier has much to do) with a register-resident sorted top-k list (bitonic
//      merge / shuffle insert) in fp64, until what is still needed is nearer than the unexplored region.
// Epilogue (knn_epilogue_kernel): one thread per query orders its row exactly by
// (fp64 d2, compact id), folds it in order into the Open3D cumulant covariances /
// FastEigen3x3 normal and writes the kNN row.
#include <math.h>
#include <stdio.h>
#include "common.cuh"
#include "primitives.cuh"
#include "upcp_math.cuh"

namespace upcp {

constexpr int kBlock = 256;
constexpr int kSelWarps = 8;
constexpr int kSelThreads = kSelWarps * 32;
constexpr int kRow = 48;                // candidate-row capacity (sorted positions, slot-major)
constexpr int kBlkWarps = 4;
constexpr int kBlkThreads = kBlkWarps * 32;
constexpr int kBuckets = 32;            // radial histogram resolution
constexpr int kEpiThreads = 128;
constexpr uint32_t kFull = 0xffffffffu;

struct GridParams {
    double ox, oy, oz, cell, inv_cell, slack;
    int nx, ny, nz;
    double extent;     // largest bounding-box edge (fp32 error bound of the filter passes)
};

// Sorted points as 32-byte records (x, y, z, compact id): one sector per gathered neighbour.
struct __align__(32) P4 { double x, y, z, w; };
__device__ __forceinline__ P4 load_p4(const P4* __restrict__ p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    P4 r; r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y;
    return r;
}
// w packs the compact id (low 32 bits) and the point's position in the fine-grid order (high 32 bits)
__device__ __forceinline__ uint32_t p4_cid(const P4& v) { return (uint32_t)__double_as_longlong(v.w); }
__device__ __forceinline__ uint32_t p4_pos(const P4& v) { return (uint32_t)((unsigned long long)__double_as_longlong(v.w) >> 32); }

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

// Counting sort by cell number in two kernels around one scan: the first computes every point's
// cell and its arrival rank within the cell (one atomic per distinct cell per warp), the scan turns
// the per-cell counts into the dense prefix table, the second places the records.  The order of
// the points INSIDE a cell is the arrival order -- not reproducible from run to run, and
// irrelevant: every consumer orders candidates by (exact squared distance, compact id).
__device__ __forceinline__ uint32_t cell_rank(uint32_t* __restrict__ cell_count, uint32_t key, bool valid) {
    const uint32_t peers = __match_any_sync(__activemask(), valid ? key : 0xffffffffu);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (valid && lane == leader) base = atomicAdd(cell_count + key, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(kBlock)
knn_keys_kernel(const double* __restrict__ x, const double* __restrict__ y,
                const double* __restrict__ z, const uint32_t* __restrict__ sel_idx, int64_t n, int64_t m,
                GridParams g, uint32_t* __restrict__ keys, uint32_t* __restrict__ rank,
                uint32_t* __restrict__ cell_count) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    const int64_t m_warp = (m + 31) & ~(int64_t)31;
    for (int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x; j < m_warp; j += stride) {
        const bool valid = j < m;
        uint32_t key = 0;
        if (valid) {
            int64_t i = sel_idx ? (int64_t)sel_idx[j] : j;
            if (i >= n) i = n - 1;                  // (only with a wrong m: stay inside the arrays)
            int cx = cell_coord(x[i], g.ox, g.inv_cell, g.nx);
            int cy = cell_coord(y[i], g.oy, g.inv_cell, g.ny);
            int cz = cell_coord(z[i], g.oz, g.inv_cell, g.nz);
            key = ((uint32_t)cz * (uint32_t)g.ny + (uint32_t)cy) * (uint32_t)g.nx + (uint32_t)cx;
        }
        const uint32_t r = cell_rank(cell_count, key, valid);
        if (valid) { keys[j] = key; rank[j] = r; }
    }
}

__global__ void __launch_bounds__(kBlock)
knn_place_kernel(const double* __restrict__ x, const double* __restrict__ y,
                 const double* __restrict__ z, const uint32_t* __restrict__ sel_idx,
                 const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rank,
                 const uint32_t* __restrict__ cell_first, int64_t n, int64_t m, GridParams g,
                 P4* __restrict__ spts, float4* __restrict__ spf, uint32_t* __restrict__ order) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x; j < m; j += stride) {
        const uint32_t s = __ldg(cell_first + keys[j]) + rank[j];
        int64_t i = sel_idx ? (int64_t)sel_idx[j] : (int64_t)j;
        if (i >= n) i = n - 1;
        const double px = x[i], py = y[i], pz = z[i];
        double2* o = reinterpret_cast<double2*>(spts + s);
        o[0] = make_double2(px, py);
        o[1] = make_double2(pz, __longlong_as_double((long long)(((unsigned long long)s << 32) | (unsigned long long)j)));
        // origin-relative fp32 copy for the filter passes (error bound: see knn_block_kernel)
        spf[s] = make_float4((float)(px - g.ox), (float)(py - g.oy), (float)(pz - g.oz), 0.0f);
        if (order) order[s] = (uint32_t)i;          // original point index of sorted position s
    }
}

// Second, coarser grid for the per-query tier, built the same way from the fine-sorted records
// (they keep their fine-grid position in w).
__global__ void __launch_bounds__(kBlock)
knn_keys2_kernel(const P4* __restrict__ spts, int64_t m, GridParams g, uint32_t* __restrict__ keys,
                 uint32_t* __restrict__ rank, uint32_t* __restrict__ cell_count,
                 const uint32_t* __restrict__ todo_cnt, int min_share) {
    if ((int64_t)*todo_cnt * min_share < m || *todo_cnt == 0) return;     // the per-query tier stays on the fine grid
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    const int64_t m_warp = (m + 31) & ~(int64_t)31;
    for (int64_t s = (int64_t)blockIdx.x * kBlock + threadIdx.x; s < m_warp; s += stride) {
        const bool valid = s < m;
        uint32_t key = 0;
        if (valid) {
            const P4 p = load_p4(spts + s);
            int cx = cell_coord(p.x, g.ox, g.inv_cell, g.nx);
            int cy = cell_coord(p.y, g.oy, g.inv_cell, g.ny);
            int cz = cell_coord(p.z, g.oz, g.inv_cell, g.nz);
            key = ((uint32_t)cz * (uint32_t)g.ny + (uint32_t)cy) * (uint32_t)g.nx + (uint32_t)cx;
        }
        const uint32_t r = cell_rank(cell_count, key, valid);
        if (valid) { keys[s] = key; rank[s] = r; }
    }
}
__global__ void __launch_bounds__(kBlock)
knn_place2_kernel(const P4* __restrict__ spts, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rank,
                  const uint32_t* __restrict__ cell_first, int64_t m, P4* __restrict__ spts2,
                  const uint32_t* __restrict__ todo_cnt, int min_share) {
    if ((int64_t)*todo_cnt * min_share < m || *todo_cnt == 0) return;
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t s = (int64_t)blockIdx.x * kBlock + threadIdx.x; s < m; s += stride) {
        const uint32_t t = __ldg(cell_first + keys[s]) + rank[s];
        const double2* src = reinterpret_cast<const double2*>(spts + s);
        double2* dst = reinterpret_cast<double2*>(spts2 + t);
        dst[0] = __ldg(src); dst[1] = __ldg(src + 1);
    }
}

// ------------------------------------------------------------------ tier 1 / 2 ----
struct BlockArgs {
    const P4* spts; const float4* spf; const uint32_t* cell_first;
    GridParams g;
    int64_t m;
    int k;
    const uint32_t* q_idx; const uint32_t* q_cnt;      // queries (sorted positions) or NULL = all m
    uint32_t* out_pos; uint8_t* out_cnt;               // rows [kRow][m], counts [m]
    uint8_t* todo;                                     // [m] 1 = not served by this tier
    int k_row; double radius2;                         // kNN row length, squared normal-estimation radius
    uint32_t max_total;                                // blocks with more candidates go to the next tier (<= 60000: u16 counters)
    double guess_factor, guess_max_share;              // first histogram range (upcp_tuning.knn_guess_*)
};

// One lane per query; see the header comment.  The per-lane histogram is kBuckets counters in
// shared memory (u16, laid out so that a lane only ever touches its own bank).
//   tau      upper edge of the histogram (squared distance).  The block can certify anything up to
//            reach^2, but its candidates mostly lie much farther than the k-th neighbour, so the
//            first attempt uses a data-driven guess -- three times the k-th squared distance a planar
//            cloud of the block's point count would have -- and only falls back to reach^2 when k is
//            not reached below it: fine buckets where the cloud is dense, few candidates kept.
//   fp32     |d2f - d2| <= 3.47 sqrt(d2) delta + 3 delta^2 + 2^-22 d2 with delta = 2^-23 extent (origin-
//            relative coordinates rounded to fp32, differences and products in fp32); a query is only
//            served when this bound is below an eighth of a bucket (pass B keeps a quarter bucket
//            beyond b*, i.e. two error bounds).
template <int R>
__global__ void __launch_bounds__(kBlkThreads)
knn_block_kernel(BlockArgs a) {
    __shared__ uint16_t s_hist[kBlkWarps][kBuckets * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t* const hist = s_hist[warp] + lane * 2;          // counter of bucket b: HIST(b), bank = lane
#define HIST(b) hist[(((b) >> 1) << 6) + ((b) & 1)]
    const GridParams g = a.g;
    const int K = a.k;
    const int64_t n_q = a.q_idx ? (int64_t)*a.q_cnt : a.m;
    const int64_t n_q_warp = (n_q + 31) & ~(int64_t)31;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double delta = g.extent * (1.01 / 8388608.0);
    const float nbf = (float)kBuckets;

    for (int64_t qb = ((int64_t)blockIdx.x * kBlkWarps + warp) * 32; qb < n_q_warp;
         qb += (int64_t)gridDim.x * kBlkThreads) {
        const int64_t qi = qb + lane;
        const bool valid = qi < n_q;
        const int64_t q = valid ? (a.q_idx ? (int64_t)a.q_idx[qi] : qi) : 0;
        const P4 qp = load_p4(a.spts + q);
        const float4 qf = __ldg(a.spf + q);
        const int cx = cell_coord(qp.x, g.ox, g.inv_cell, g.nx);
        const int cy = cell_coord(qp.y, g.oy, g.inv_cell, g.ny);
        const int cz = cell_coord(qp.z, g.oz, g.inv_cell, g.nz);
        // reach: distance to the nearest face of the (2R+1)^3 block that has grid beyond it
        double reach = inf;
        if (cx - R > 0) reach = fmin(reach, qp.x - (g.ox + (double)(cx - R) * g.cell));
        if (cx + R < g.nx - 1) reach = fmin(reach, (g.ox + (double)(cx + R + 1) * g.cell) - qp.x);
        if (cy - R > 0) reach = fmin(reach, qp.y - (g.oy + (double)(cy - R) * g.cell));
        if (cy + R < g.ny - 1) reach = fmin(reach, (g.oy + (double)(cy + R + 1) * g.cell) - qp.y);
        if (cz - R > 0) reach = fmin(reach, qp.z - (g.oz + (double)(cz - R) * g.cell));
        if (cz + R < g.nz - 1) reach = fmin(reach, (g.oz + (double)(cz + R + 1) * g.cell) - qp.z);
        // block covers the grid on every side: every pair of points is within the block diagonal
        if (!(reach < inf)) reach = 1.7320508075688774 * (double)(2 * R + 1) * g.cell * 1.001 + g.slack;
        reach -= g.slack;
        const double tau_full = reach * reach * (1.0 - 1e-9);
        const int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
        bool active = valid && reach > 0.0;
        // Pruning inside the block.  Both passes only look at candidates below a squared distance (the
        // histogram range, then the keep limit) that in dense surroundings is far smaller than a cell: a
        // grid row -- or the outer cells of its x range -- whose nearest face lies beyond that distance
        // cannot hold such a candidate and is not walked.  near(d, f): lower bound of the coordinate
        // difference to any point of the d-th cell from the query's own, f = the query's offset inside its
        // cell; fp32 with the grid's assignment slack and a 2^-6 margin on the limit (the fp32 distance
        // error is below 2^-8 of the range, see above), so the pruned set is exactly what the unpruned
        // passes would have counted / kept.
        const float cellf = (float)g.cell, slackf = (float)(g.slack + g.cell * 1e-6);
        const float fxq = (float)(qp.x - (g.ox + (double)cx * g.cell));
        const float fyq = (float)(qp.y - (g.oy + (double)cy * g.cell));
        const float fzq = (float)(qp.z - (g.oz + (double)cz * g.cell));
        auto near = [&](int d, float f) -> float {
            if (d == 0) return 0.0f;
            const float v = (d < 0 ? f + (float)(-d - 1) * cellf : (cellf - f) + (float)(d - 1) * cellf) - slackf;
            return v > 0.0f ? v : 0.0f;
        };
        // x range of row (dy, dz) that can hold a candidate within sqrt(lim2): false = none
        auto row_range = [&](int dy, int dz, float lim2, int& xa, int& xb) -> bool {
            const float ay = near(dy, fyq), az = near(dz, fzq);
            const float yz2 = ay * ay + az * az;
            if (yz2 > lim2) return false;
            int ea = 0, eb = 0;
#pragma unroll
            for (int d = 1; d <= R; ++d) {
                const float a = near(-d, fxq), b = near(d, fxq);
                if (a * a + yz2 <= lim2) ea = d;
                if (b * b + yz2 <= lim2) eb = d;
            }
            xa = max(cx - ea, x0); xb = min(cx + eb, x1);
            return true;
        };

        // cheap certificate first: a block with fewer than k_row points cannot serve the query
        // (2 prefix lookups per row instead of a wasted counting pass)
        uint32_t total = 0;
#pragma unroll 1
        for (int dz = -R; dz <= R; ++dz) {
#pragma unroll 1
            for (int dy = -R; dy <= R; ++dy) {
                const int yy = cy + dy, zz = cz + dz;
                if (active && yy >= 0 && yy < g.ny && zz >= 0 && zz < g.nz) {
                    const uint32_t row = ((uint32_t)zz * (uint32_t)g.ny + (uint32_t)yy) * (uint32_t)g.nx;
                    total += __ldg(a.cell_first + row + x1 + 1) - __ldg(a.cell_first + row + x0);
                }
            }
        }
        // (u16 counters: a block with more candidates than they can count goes to the next tier)
        if (total < (uint32_t)a.k_row || total > a.max_total) active = false;
        if (!__any_sync(kFull, active)) {
            if (valid) a.todo[q] = 1;
            continue;
        }
        // histogram edge: data-driven guess first, reach^2 as the fall-back
        double tau = tau_full;
        {
            const double side = (double)(2 * R + 1) * g.cell;
            const double guess = a.guess_factor * (double)K * side * side / (3.141592653589793 * (double)(total > 0 ? total : 1));
            if (guess < a.guess_max_share * tau) tau = guess;        // only where the full range would be too coarse
        }
        auto fp32_ok = [&](double t) {
            const double err = 3.47 * sqrt(t) * delta + 3.0 * delta * delta + t * (1.0 / 4194304.0);
            return 8.0 * err <= t * (1.0 / kBuckets);
        };
        if (!fp32_ok(tau)) tau = tau_full;
        if (!fp32_ok(tau)) active = false;
        float scale = (float)((double)kBuckets / tau);

        float tkeep = 0.0f;        // pass B keeps the candidates with d2f * scale below this
        bool counting = active;    // this lane takes part in the current counting pass
        for (int attempt = 0; attempt < 2; ++attempt) {
            if (!__any_sync(kFull, counting)) break;
            if (counting) {
#pragma unroll
                for (int b = 0; b < kBuckets; ++b) HIST(b) = 0;
            }
            // ---- pass A: radial histogram of the block's candidates
            const float lim_a = (float)(tau * (1.0 + 1.0 / 64.0));
#pragma unroll 1
            for (int dz = -R; dz <= R; ++dz) {
#pragma unroll 1
                for (int dy = -R; dy <= R; ++dy) {
                    const int yy = cy + dy, zz = cz + dz;
                    uint32_t p = 0, pb = 0;
                    int xa = x0, xb = x1;
                    if (counting && yy >= 0 && yy < g.ny && zz >= 0 && zz < g.nz && row_range(dy, dz, lim_a, xa, xb)) {
                        const uint32_t row = ((uint32_t)zz * (uint32_t)g.ny + (uint32_t)yy) * (uint32_t)g.nx;
                        p = __ldg(a.cell_first + row + xa);
                        pb = __ldg(a.cell_first + row + xb + 1);
                    }
                    for (; p + 1 < pb; p += 2) {
                        const float4 c0 = __ldg(a.spf + p), c1 = __ldg(a.spf + p + 1);
                        const float ax = c0.x - qf.x, ay = c0.y - qf.y, az = c0.z - qf.z;
                        const float bx = c1.x - qf.x, by = c1.y - qf.y, bz = c1.z - qf.z;
                        const float t0 = fmaf(az, az, fmaf(ay, ay, ax * ax)) * scale;
                        const float t1 = fmaf(bz, bz, fmaf(by, by, bx * bx)) * scale;
                        if (t0 < nbf) HIST(__float2int_rz(t0)) += 1;
                        if (t1 < nbf) HIST(__float2int_rz(t1)) += 1;
                    }
                    if (p < pb) {
                        const float4 c0 = __ldg(a.spf + p);
                        const float ax = c0.x - qf.x, ay = c0.y - qf.y, az = c0.z - qf.z;
                        const float t0 = fmaf(az, az, fmaf(ay, ay, ax * ax)) * scale;
                        if (t0 < nbf) HIST(__float2int_rz(t0)) += 1;
                    }
                }
            }
            // ---- select.  What the consumers use is the k_row nearest (kNN row, covariance) and, of the
            // K nearest, those within the normal radius.  The block certifies that when
            //   (a) the K-th candidate falls below the last bucket (everything needed is nearer), or
            //   (b) the k_row-th does AND the radius lies inside the histogram range (then every point
            //       within the radius is in the block, and there are fewer than K of them);
            // pass B keeps the candidates below the smaller of the two limits, plus the fp32 margin.
            // The counters become exclusive offsets so that pass B places the kept candidates bucket-ordered.
            if (counting) {
                uint32_t cum = 0;
                int bK = -1, bk = -1;           // first bucket whose cumulative count reaches K / k_row
#pragma unroll 1
                for (int b = 0; b < kBuckets; ++b) {
                    const uint32_t hb = HIST(b);
                    HIST(b) = (uint16_t)cum;
                    cum += hb;
                    if (bk < 0 && cum >= (uint32_t)a.k_row) bk = b;
                    if (bK < 0 && cum >= (uint32_t)K) bK = b;
                }
                const float rb = (float)(a.radius2 * (double)scale);           // the radius in bucket units
                float lim = -1.0f;
                if (bK >= 0 && bK <= kBuckets - 2) lim = (float)bK + 1.25f;
                if (bk >= 0 && bk <= kBuckets - 2 && rb + 0.25f <= (float)kBuckets) {
                    const float alt = fmaxf((float)bk + 1.25f, rb + 0.25f);
                    if (lim < 0.0f || alt < lim) lim = alt;
                }
                // otherwise: retry with the full range, then give up
                if (lim >= 0.0f) { tkeep = lim; counting = false; }
                else if (attempt == 0 && tau < tau_full && fp32_ok(tau_full)) {
                    tau = tau_full;
                    scale = (float)((double)kBuckets / tau);
                } else { active = false; counting = false; }
            }
        }
        // ---- pass B: append the candidates of buckets <= b* (+ a quarter bucket) to the row
        bool overflow = false;
        uint32_t kept = 0;
        // row slot s of query q lives at out_pos[s * m + q]: one 32 x 32 -> 64-bit multiply-add per store
        // (m < 2^30 is checked at entry)
        char* const out_q = reinterpret_cast<char*>(a.out_pos + q);
        const uint32_t row_stride = (uint32_t)a.m * 4u;
        if (__any_sync(kFull, active)) {
            const float lim_b = (float)(((double)tkeep / (double)kBuckets + 1.0 / 64.0) * tau);
#pragma unroll 1
            for (int dz = -R; dz <= R; ++dz) {
#pragma unroll 1
                for (int dy = -R; dy <= R; ++dy) {
                    const int yy = cy + dy, zz = cz + dz;
                    uint32_t p = 0, pb = 0;
                    int xa = x0, xb = x1;
                    if (active && yy >= 0 && yy < g.ny && zz >= 0 && zz < g.nz && row_range(dy, dz, lim_b, xa, xb)) {
                        const uint32_t row = ((uint32_t)zz * (uint32_t)g.ny + (uint32_t)yy) * (uint32_t)g.nx;
                        p = __ldg(a.cell_first + row + xa);
                        pb = __ldg(a.cell_first + row + xb + 1);
                    }
                    for (; p < pb; ++p) {
                        const float4 c0 = __ldg(a.spf + p);
                        const float ax = c0.x - qf.x, ay = c0.y - qf.y, az = c0.z - qf.z;
                        const float t0 = fmaf(az, az, fmaf(ay, ay, ax * ax)) * scale;
                        if (t0 < tkeep) {
                            const int b0 = __float2int_rz(t0);
                            const uint32_t slot = HIST(b0);
                            HIST(b0) = (uint16_t)(slot + 1);
                            ++kept;
                            if (slot < (uint32_t)kRow) *reinterpret_cast<uint32_t*>(out_q + (uint64_t)slot * row_stride) = p;
                            else overflow = true;
                        }
                    }
                }
            }
        }
        if (valid) {
            bool served = false;
            if (active && !overflow && kept <= (uint32_t)kRow) { a.out_cnt[q] = (uint8_t)kept; served = true; }
            a.todo[q] = served ? 0 : 1;
        }
    }
}
#undef HIST

// ------------------------------------------------------------------ tier 3 ------
struct SelectArgs {
    const P4* qpts;                // query records, fine-grid order
    // candidate records, cell prefix and geometry of [0] the fine grid and [1] the coarse second grid;
    // the kernel uses [1] when at least m / coarse_min_share queries are left (the builders of the
    // second grid test the same device counter: no host round trip), 0 = no second grid
    const P4* spts[2];
    const uint32_t* cell_first[2];
    GridParams g[2];
    int coarse_min_share;
    int64_t m;
    int k;                         // neighbours kept per query (<= 32), clipped to m
    const uint32_t* todo_idx;      // queries (sorted positions) to process, or NULL = all m
    const uint32_t* todo_cnt;      // device count of todo_idx
    uint32_t* out_pos;             // [kRow][m] sorted positions of the k nearest, ascending
    uint8_t* out_cnt;              // [m] number of valid entries of the row
    // What the consumers use: the k_row nearest (kNN row, covariance) and, of the k nearest, those
    // within radius2 (hybrid-search normals).  Beyond BOTH the k_row-th distance and the radius
    // nothing matters, so the walk may stop there instead of at the k-th distance.
    int k_row; double radius2;
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
    __shared__ uint32_t s_seg[kSelWarps][96];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* const sa_off = s_seg[warp]; uint32_t* const sa_base = sa_off + 64;

    const int K = a.k;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);

    const int64_t n_todo = a.todo_idx ? (int64_t)*a.todo_cnt : a.m;
    const int which = (a.coarse_min_share > 0 && n_todo > 0 && n_todo * a.coarse_min_share >= a.m) ? 1 : 0;
    const GridParams g = a.g[which];
    const P4* const cand_pts = a.spts[which];
    const uint32_t* const cfirst = a.cell_first[which];
    for (int64_t qi = (int64_t)blockIdx.x * kSelWarps + warp; qi < n_todo; qi += (int64_t)gridDim.x * kSelWarps) {
        const int64_t q = a.todo_idx ? (int64_t)a.todo_idx[qi] : qi;
        const P4 qp = load_p4(a.qpts + q);
        const double qx = qp.x, qy = qp.y, qz = qp.z;
        const int cx = cell_coord(qx, g.ox, g.inv_cell, g.nx);
        const int cy = cell_coord(qy, g.oy, g.inv_cell, g.ny);
        const int cz = cell_coord(qz, g.oz, g.inv_cell, g.nz);

        Nb mine; mine.d = inf; mine.c = kFull; mine.p = 0;   // lane j: j-th nearest so far
        double kth_d = inf; uint32_t kth_c = kFull;          // the K-th entry (lane K-1), warp-uniform

        double need_d = inf;                                  // min(k-th, max(k_row-th, radius^2)): see SelectArgs
        auto refresh_kth = [&]() {
            kth_d = __shfl_sync(kFull, mine.d, K - 1);
            kth_c = __shfl_sync(kFull, mine.c, K - 1);
            need_d = fmin(kth_d, fmax(__shfl_sync(kFull, mine.d, a.k_row - 1), a.radius2));
        };

        // Stream the points of up to 32 ranges (one per lane: first, cnt) past the query.
        auto process_ranges = [&](uint32_t first, uint32_t cnt) {
            const uint32_t incl = warp_incl_scan_u32(cnt, lane);
            const uint32_t total = __shfl_sync(kFull, incl, 31);
            if (total == 0) return;
            sa_off[lane] = incl - cnt; sa_off[32 + lane] = total; sa_base[lane] = first;
            __syncwarp();
            for (uint32_t b = 0; b < total; b += 32) {
                const uint32_t i = b + lane;
                Nb cand; cand.d = inf; cand.c = kFull; cand.p = 0;
                if (i < total) {
                    const int kk = seg_search(sa_off, i);
                    const uint32_t p = sa_base[kk] + (i - sa_off[kk]);
                    const P4 cp = load_p4(cand_pts + p);
                    cand.d = dist2(cp.x, cp.y, cp.z, qx, qy, qz);
                    cand.c = p4_cid(cp);
                    cand.p = p4_pos(cp);           // position in the fine-grid order (what the rows hold)
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

        // points of the cells lo..hi (flat cell numbers of one grid row)
        auto range = [&](uint32_t lo, uint32_t hi, uint32_t& first, uint32_t& cnt) {
            first = __ldg(cfirst + lo);
            cnt = __ldg(cfirst + hi + 1) - first;
        };
        // squared gap between the query and the slab of cell index c along one axis (conservative)
        auto axis_gap = [&](double qv, double origin, int c, int cq) -> double {
            if (c == cq) return 0.0;
            const double gap = c > cq ? (origin + (double)c * g.cell) - qv : qv - (origin + (double)(c + 1) * g.cell);
            return fmax(0.0, gap - g.slack);
        };

        // Thick shells of the dense grid: stage (Rdone, Rn] covers the rows (dy, dz) in [-Rn, Rn]^2;
        // rows outside the finished block contribute their whole x-range [cx - Rn, cx + Rn] (ONE
        // contiguous range), rows inside it the two end pieces.  Rows and range ends that cannot
        // hold anything nearer than the current k-th distance are dropped before any load.  The
        // next radius is a jump, not a step: straight to the radius that certifies the current k-th
        // distance once k neighbours are known, else a density-driven guess.
        int Rdone = -1, Rn = 1;
        const int Rmax = max(g.nx, max(g.ny, g.nz));
        while (true) {
            // Rows are taken ring by ring around the query's own row, nearest first: what the inner rings find
            // tightens need_d before the outer rings are set up, and most of those are dropped before any load
            // (an isolated point next to a dense surface meets the surface in a few rows, not in all of them).
            for (int r = 0; r <= Rn; ++r) {
                // the four edges of ring r: dy = -r, dy = +r, dz = -r, dz = +r
                if (cy - r < 0 && cy + r >= g.ny && cz - r < 0 && cz + r >= g.nz) break;      // beyond the grid
                if (need_d < inf) {
                    const double gmin = fmax(0.0, (double)(r - 1) * g.cell - g.slack);          // (y, z) gap of every row of the ring
                    if (gmin * gmin * (1.0 - 1e-9) > need_d) break;
                }
                const int ring_n = r == 0 ? 1 : 8 * r;
                for (int j0 = 0; j0 < ring_n; j0 += 32) {
                    const int j = j0 + lane;
                    uint32_t f1 = 0, n1 = 0, f2 = 0, n2 = 0;
                    int dy = 0, dz = 0;
                    if (r > 0 && j < ring_n) {
                        const int side = j / (2 * r), o = j - side * 2 * r;
                        dy = side == 0 ? o - r : side == 1 ? r : side == 2 ? r - o : -r;
                        dz = side == 0 ? -r : side == 1 ? o - r : side == 2 ? r : r - o;
                    }
                    const int yy = cy + dy, zz = cz + dz;
                    if (j < ring_n && yy >= 0 && yy < g.ny && zz >= 0 && zz < g.nz) {
                        const double gy = axis_gap(qy, g.oy, yy, cy), gz = axis_gap(qz, g.oz, zz, cz);
                        const double lb_yz = (gy * gy + gz * gz) * (1.0 - 1e-9);
                        if (!(lb_yz > need_d)) {
                            const uint32_t row = ((uint32_t)zz * (uint32_t)g.ny + (uint32_t)yy) * (uint32_t)g.nx;
                            // along x only the cells that overlap [qx - s, qx + s] can hold anything still needed,
                            // s^2 = what the row's (y, z) gap leaves of need_d (plus the assignment slack)
                            int xa = max(cx - Rn, 0), xb = min(cx + Rn, g.nx - 1);
                            if (need_d < inf) {
                                const double sx = sqrt(need_d - lb_yz) * (1.0 + 1e-9) + 2.0 * g.slack;
                                xa = max(xa, cell_coord(qx - sx, g.ox, g.inv_cell, g.nx));
                                xb = min(xb, cell_coord(qx + sx, g.ox, g.inv_cell, g.nx));
                            }
                            if (r > Rdone) {
                                range(row + xa, row + xb, f1, n1);
                            } else {                                   // inside the finished block: the two end pieces
                                const int xl = cx - Rdone - 1, xr = cx + Rdone + 1;
                                if (xa <= xl) range(row + xa, row + xl, f1, n1);
                                if (xr <= xb) range(row + xr, row + xb, f2, n2);
                            }
                        }
                    }
                    if (__any_sync(kFull, n1 != 0)) process_ranges(f1, n1);
                    if (__any_sync(kFull, n2 != 0)) process_ranges(f2, n2);
                }
            }
            Rdone = Rn;
            // ---- termination: distance from the query to the unexplored region (warp-uniform)
            const bool open_x0 = cx - Rdone > 0, open_x1 = cx + Rdone < g.nx - 1;
            const bool open_y0 = cy - Rdone > 0, open_y1 = cy + Rdone < g.ny - 1;
            const bool open_z0 = cz - Rdone > 0, open_z1 = cz + Rdone < g.nz - 1;
            if (!(open_x0 || open_x1 || open_y0 || open_y1 || open_z0 || open_z1)) break;   // whole grid explored
            double reach = inf;
            if (open_x0) reach = fmin(reach, qx - (g.ox + (double)(cx - Rdone) * g.cell));
            if (open_x1) reach = fmin(reach, (g.ox + (double)(cx + Rdone + 1) * g.cell) - qx);
            if (open_y0) reach = fmin(reach, qy - (g.oy + (double)(cy - Rdone) * g.cell));
            if (open_y1) reach = fmin(reach, (g.oy + (double)(cy + Rdone + 1) * g.cell) - qy);
            if (open_z0) reach = fmin(reach, qz - (g.oz + (double)(cz - Rdone) * g.cell));
            if (open_z1) reach = fmin(reach, (g.oz + (double)(cz + Rdone + 1) * g.cell) - qz);
            reach -= g.slack;
            if (need_d < inf && reach > 0.0 && need_d < reach * reach * (1.0 - 1e-9)) break;
            // ---- next radius
            if (need_d < inf) {
                // reach(R) >= R * cell - slack: this radius certifies what is still needed
                const double need = (sqrt(need_d) + 2.0 * g.slack) * g.inv_cell;
                Rn = need < (double)Rmax ? (int)need + 1 : Rmax;
            } else {
                const int found = __popc(__ballot_sync(kFull, mine.d < inf));
                Rn = 2 * Rdone + 2;
                if (found > 0) {
                    // planar guess: the count grows with the square of the block edge
                    const int guess = (int)(((double)(2 * Rdone + 1) * sqrt((double)K / (double)found) * 1.15 - 1.0) * 0.5) + 1;
                    if (guess < Rn) Rn = guess;
                }
            }
            if (Rn <= Rdone) Rn = Rdone + 1;
            if (Rn > Rmax) Rn = Rmax;
        }
        // sorted positions of the neighbours, ascending
        const bool have = lane < K && mine.d < inf;
        a.out_pos[(size_t)lane * a.m + q] = have ? mine.p : kFull;
        const uint32_t hv = __ballot_sync(kFull, have);
        if (lane == 0) a.out_cnt[q] = (uint8_t)__popc(hv);
    }
}

struct TodoFunctor {
    uint32_t* todo_idx;
    __device__ void operator()(int64_t s, bool is_todo, uint32_t rank) const {
        if (is_todo) todo_idx[rank] = (uint32_t)s;
    }
};

// ------------------------------------------------------------------ epilogue ----
struct EpilogueArgs {
    const P4* spts;
    const uint32_t* pos; const uint8_t* cnt;      // pos: [kRow][m] (slot-major: coalesced per query thread)
    int64_t m;
    int knn, max_nn;
    double radius2;
    int32_t* out_knn; double* out_d2; double* out_normals; double* out_cov; double* out_curv;
    double thr_curve; uint8_t* out_flag;
    int sorted_out;     // 0: outputs indexed by compact id, neighbours as compact ids; 1: both as sorted positions
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
        const uint32_t cid = a.sorted_out ? (uint32_t)s : p4_cid(qp);     // output row of this query
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
        // phase 2: insertion sort of the packed keys.  ONE flat loop whose step is either "shift one
        // entry" or "place and fetch the next": a lane's trip count is its own n + inversions, so the
        // warp pays the maximum over its lanes once instead of the maximum per inserted element.
        bool tie = false;
        {
            int j = 1, i = 1;
            uint64_t kj = n > 1 ? s_key[kEpiThreads + tx] : 0;
            while (j < n) {
                const uint64_t kp = i > 0 ? s_key[(i - 1) * kEpiThreads + tx] : 0;
                const bool shift = i > 0 && kp > kj;
                tie |= i > 0 && ((kp ^ kj) >> 6) == 0;
                if (shift) {
                    s_key[i * kEpiThreads + tx] = kp;
                    --i;
                } else {
                    if (i != j) s_key[i * kEpiThreads + tx] = kj;
                    ++j;
                    i = j;
                    if (j < n) kj = s_key[j * kEpiThreads + tx];
                }
            }
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
                if (present) { d2 = dist2(px, py, pz, qx, qy, qz); c = a.sorted_out ? pp[u] : p4_cid(cp[u]); }
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
            bool want_cov = a.out_cov != nullptr && !a.sorted_out;      // sorted space: only where the flag is undecided
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
                    if (!below && !above) want_cov = a.out_cov != nullptr;
                }
            } else {
                want_cov = a.out_cov != nullptr;
            }
            if (want_cov)
                for (int t = 0; t < 6; ++t) a.out_cov[(size_t)cid * 6 + t] = cov6[t];
        }
    }
}

// the dense grid may hold up to this many cells for m selected points
static inline uint64_t grid_cell_cap(int64_t m) {
    uint64_t cap = 16ull * (uint64_t)(m < 1 ? 1 : m) + (1ull << 20);
    return cap < (1ull << 28) ? cap : (1ull << 28);
}

static int knn_impl(const double* d_x, const double* d_y, const double* d_z, int64_t n,
                    const uint8_t* d_sel, int64_t m, GridParams g,
                    int knn, int max_nn, double radius,
                    int32_t* d_knn, double* d_knn_d2, double* d_normals, double* d_cov, double* d_curv,
                    double thr_curve, uint8_t* d_flag, uint32_t* d_order, void* d_ws, size_t ws_bytes, cudaStream_t st) {
    const uint64_t G = (uint64_t)g.nx * (uint64_t)g.ny * (uint64_t)g.nz;
    Workspace ws(d_ws, ws_bytes);
    uint32_t* counters = ws.take<uint32_t>(64);          // [0] m, [1] tier-2 queries, [2] tier-3 queries
    uint32_t* sel_idx = d_sel ? ws.take<uint32_t>((size_t)n) : nullptr;
    uint32_t* prim_ws = ws.take<uint32_t>(compact_ws_words(n) + sort_ws_words(m) + scan_ws_words((int64_t)grid_cell_cap(m) + 1));
    uint32_t* keys = ws.take<uint32_t>((size_t)m);
    uint32_t* rank = ws.take<uint32_t>((size_t)m);
    P4* spts = ws.take<P4>((size_t)m);
    float4* spf = ws.take<float4>((size_t)m);
    uint32_t* cell_first = ws.take<uint32_t>((size_t)grid_cell_cap(m) + 1);
    P4* spts2 = ws.take<P4>((size_t)m);                                   // coarse grid of the per-query tier
    uint32_t* cell_first2 = ws.take<uint32_t>((size_t)grid_cell_cap(m) / 4 + 2);
    uint32_t* nb_pos = ws.take<uint32_t>((size_t)m * kRow);
    uint8_t* nb_cnt = ws.take<uint8_t>((size_t)m);
    uint8_t* todo[2] = {ws.take<uint8_t>((size_t)m), ws.take<uint8_t>((size_t)m)};
    uint32_t* todo_idx[2] = {ws.take<uint32_t>((size_t)m), ws.take<uint32_t>((size_t)m)};
    if (!ws.ok) UPCP_FAIL(UPCP_E_WORKSPACE, "knn: workspace too small");
    if (m >= ((int64_t)1 << 30)) UPCP_FAIL(UPCP_E_CAPACITY, "knn: at most 2^30 - 1 selected points");
    if (G > grid_cell_cap(m)) UPCP_FAIL(UPCP_E_INTERNAL, "knn: grid larger than its cap");

    const upcp_tuning tune = tuning();
    prof_begin(UPCP_PROF_KNN, st);
    prof_begin(UPCP_PROF_KNN_BUILD, st);
    if (d_sel) {
        // (m is the caller's count of the same mask -- DeviceTile.bbox delivers it with the bounding box --;
        // it is not re-read from the device here: no host round trip inside the search.  The builders
        // clamp their gathers, so a wrong m gives unspecified rows but never an out-of-bounds access.)
        KnnCompactIdx f{sel_idx};
        UPCP_TRY(compact_apply(d_sel, n, counters, prim_ws, st, f));
    } else if (m != n) {
        UPCP_FAIL(UPCP_E_INVALID, "knn: m must equal n when no selection is given");
    }

    // cell numbers + arrival ranks + per-cell counts, exclusive prefix over ALL cells, placement
    const int grid_m = grid_for(m, kBlock);
    UPCP_CUDA_OK(cudaMemsetAsync(cell_first, 0, ((size_t)G + 1) * 4, st));
    knn_keys_kernel<<<grid_m, kBlock, 0, st>>>(d_x, d_y, d_z, sel_idx, n, m, g, keys, rank, cell_first);
    UPCP_LAUNCH_OK();
    UPCP_TRY(exclusive_scan_u32(cell_first, cell_first, (int64_t)G + 1, nullptr, prim_ws, st));
    knn_place_kernel<<<grid_m, kBlock, 0, st>>>(d_x, d_y, d_z, sel_idx, keys, rank, cell_first, n, m, g, spts, spf, d_order);
    UPCP_LAUNCH_OK();
    prof_end(UPCP_PROF_KNN_BUILD, st, m);

    int kh = knn > max_nn ? knn : max_nn;
    if ((int64_t)kh > m) kh = (int)m;
    prof_begin(UPCP_PROF_KNN_SEARCH, st);
    // block tiers: lane per query over its (2R+1)^3 cell block; each tier takes the queries the
    // previous one could not certify
    BlockArgs ba;
    ba.spts = spts; ba.spf = spf; ba.cell_first = cell_first; ba.g = g; ba.m = m; ba.k = kh;
    ba.q_idx = nullptr; ba.q_cnt = nullptr; ba.out_pos = nb_pos; ba.out_cnt = nb_cnt; ba.todo = todo[0];
    ba.k_row = knn < kh ? knn : kh; ba.radius2 = radius * radius;
    // A lane that has to scan a crowded 5^3 block alone holds its whole warp up; such queries are better
    // served by the warp-per-query tier, which prunes (swept on the bench tile: 60000 / 4000 / 1500 / 600 /
    // 300 candidates -> 17.6 / 17.4 / 17.2 / 16.5 / 17.0 ms for the search): upcp_tuning.knn_tier2_max_block.
    ba.max_total = (uint32_t)tune.knn_tier1_max_block;
    ba.guess_factor = 0.1 * (double)tune.knn_guess_tenths; ba.guess_max_share = 0.01 * (double)tune.knn_guess_max_pct;
    // the filter passes live on L1 hits (each candidate is read by ~150 queries, twice): keep the
    // shared-memory carve-out at what the resident CTAs need, the rest of the 256 KB stays L1
    UPCP_CUDA_OK(cudaFuncSetAttribute(knn_block_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 40));
    UPCP_CUDA_OK(cudaFuncSetAttribute(knn_block_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 40));
    {
        const int64_t cap_blocks = (int64_t)num_sms() * 32;
        const int64_t need = (m + kBlkThreads - 1) / kBlkThreads;
        knn_block_kernel<1><<<(int)(need < cap_blocks ? need : cap_blocks), kBlkThreads, 0, st>>>(ba);
        UPCP_LAUNCH_OK();
    }
    int last = 0;                                           // todo[last]: what the tiers so far left over
    if (tune.knn_tiers >= 2) {
        TodoFunctor f{todo_idx[0]};
        UPCP_TRY(compact_apply(todo[0], m, counters + 1, prim_ws, st, f));
        UPCP_CUDA_OK(cudaMemsetAsync(todo[1], 0, (size_t)m, st));
        ba.q_idx = todo_idx[0]; ba.q_cnt = counters + 1; ba.todo = todo[1];
        ba.max_total = (uint32_t)tune.knn_tier2_max_block;
        knn_block_kernel<2><<<num_sms() * 8, kBlkThreads, 0, st>>>(ba);
        UPCP_LAUNCH_OK();
        last = 1;
    }
    // last tier: the rest (sparse surroundings, overflowing rows): warp per query
    uint32_t* n_todo_last = counters + 2;
    uint32_t* todo_last_idx = todo_idx[last];
    {
        TodoFunctor f{todo_last_idx};
        UPCP_TRY(compact_apply(todo[last], m, n_todo_last, prim_ws, st, f));
    }
    // The per-query tier walks grid rows; where it has much to do (sparse clutter, noise) it gets a
    // second grid with `coarse` times larger cells -- coarse^2 times fewer rows per shell -- built
    // from the fine-sorted records (one more counting sort).  Whether it pays is decided ON THE DEVICE
    // from the number of queries left (the builders and the search read the same counter): no host
    // round trip.
    SelectArgs a;
    a.qpts = spts; a.m = m; a.k = kh;
    a.spts[0] = spts; a.cell_first[0] = cell_first; a.g[0] = g;
    a.spts[1] = spts; a.cell_first[1] = cell_first; a.g[1] = g; a.coarse_min_share = 0;
    a.todo_idx = todo_last_idx; a.todo_cnt = n_todo_last; a.out_pos = nb_pos; a.out_cnt = nb_cnt;
    a.k_row = knn < kh ? knn : kh; a.radius2 = radius * radius;
    const int coarse = tune.knn_coarse;
    if (coarse > 1) {
        GridParams g2 = g;
        g2.cell = g.cell * (double)coarse; g2.inv_cell = 1.0 / g2.cell;
        g2.nx = (g.nx + coarse - 1) / coarse; g2.ny = (g.ny + coarse - 1) / coarse; g2.nz = (g.nz + coarse - 1) / coarse;
        g2.slack = g2.cell * 1e-6 + (g.slack - g.cell * 1e-6);
        const uint64_t G2 = (uint64_t)g2.nx * (uint64_t)g2.ny * (uint64_t)g2.nz;
        if (G2 + 1 <= grid_cell_cap(m) / 4 + 2) {
            // (keys / rank are free again)
            const int share = 200;                      // second grid when at least m / 200 queries are left
            UPCP_CUDA_OK(cudaMemsetAsync(cell_first2, 0, ((size_t)G2 + 1) * 4, st));
            knn_keys2_kernel<<<grid_m, kBlock, 0, st>>>(spts, m, g2, keys, rank, cell_first2, n_todo_last, share);
            UPCP_LAUNCH_OK();
            UPCP_TRY(exclusive_scan_u32(cell_first2, cell_first2, (int64_t)G2 + 1, nullptr, prim_ws, st));
            knn_place2_kernel<<<grid_m, kBlock, 0, st>>>(spts, keys, rank, cell_first2, m, spts2, n_todo_last, share);
            UPCP_LAUNCH_OK();
            a.spts[1] = spts2; a.cell_first[1] = cell_first2; a.g[1] = g2; a.coarse_min_share = share;
        }
    }
    knn_select_kernel<<<num_sms() * 8, kSelThreads, 0, st>>>(a);
    UPCP_LAUNCH_OK();
    prof_end(UPCP_PROF_KNN_SEARCH, st, m);

    prof_begin(UPCP_PROF_KNN_EPILOGUE, st);
    EpilogueArgs e;
    e.spts = spts; e.pos = nb_pos; e.cnt = nb_cnt; e.m = m;
    e.knn = knn; e.max_nn = max_nn; e.radius2 = radius * radius;
    e.out_knn = d_knn; e.out_d2 = d_knn_d2; e.out_normals = d_normals; e.out_cov = d_cov; e.out_curv = d_curv;
    e.thr_curve = thr_curve; e.out_flag = d_flag; e.sorted_out = d_order ? 1 : 0;
    const size_t epi_smem = (size_t)kRow * kEpiThreads * 8;
    UPCP_CUDA_OK(cudaFuncSetAttribute(knn_epilogue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)epi_smem));
    knn_epilogue_kernel<<<grid_for(m, kEpiThreads, 4), kEpiThreads, epi_smem, st>>>(e);
    UPCP_LAUNCH_OK();
    prof_end(UPCP_PROF_KNN_EPILOGUE, st, m);
    prof_end(UPCP_PROF_KNN, st, m);
    return UPCP_OK;
}

}  // namespace upcp

using namespace upcp;

extern "C" {

size_t upcp_knn_ws_bytes(int64_t n, int64_t m, int kmax) {
    (void)kmax;
    if (n < 1) n = 1;
    if (m < 1 || m > n) m = n;
    const size_t gcap = (size_t)grid_cell_cap(m);
    size_t bytes = align_up(64 * 4);
    bytes += align_up((size_t)n * 4);
    bytes += align_up((compact_ws_words(n) + sort_ws_words(m) + scan_ws_words((int64_t)gcap + 1)) * 4);
    bytes += 2 * align_up((size_t)m * 4);                                   // cell numbers, arrival ranks
    bytes += align_up((size_t)m * 32);                                      // sorted points (x, y, z, id)
    bytes += align_up((size_t)m * 16);                                      // fp32 origin-relative copy
    bytes += align_up((gcap + 1) * 4);                                      // dense cell prefix
    bytes += align_up((size_t)m * 32) + align_up((gcap / 4 + 2) * 4);       // coarse grid of the per-query tier
    bytes += align_up((size_t)m * kRow * 4) + 3 * align_up((size_t)m);      // neighbour rows, counts, todo flags
    bytes += 2 * align_up((size_t)m * 4);                                   // todo lists
    return bytes + 4096;
}

static int knn_entry(const double* d_x, const double* d_y, const double* d_z, int64_t n,
                     const uint8_t* d_sel, int64_t m, const double* h_bbox_sel6,
                     int knn, int max_nn, double radius, double cell_size,
                     int32_t* d_knn, double* d_knn_d2, double* d_normals, double* d_cov, double* d_curv,
                     double threshold_curve, uint8_t* d_seed_flag, uint32_t* d_order,
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
    // dense grid: at most grid_cell_cap(m) cells (and 2^20 per axis); coarsen the cells until it fits
    double emax = fmax(ex, fmax(ey, ez));
    double cell = cell_size;
    if (emax / cell > 1000000.0) cell = emax / 1000000.0;
    const double gcap = (double)grid_cell_cap(m);
    while ((floor(ex / cell) + 1) * (floor(ey / cell) + 1) * (floor(ez / cell) + 1) > gcap) cell *= 1.0625;
    g.cell = cell; g.inv_cell = 1.0 / cell;
    g.nx = (int)floor(ex / cell) + 1; g.ny = (int)floor(ey / cell) + 1; g.nz = (int)floor(ez / cell) + 1;
    double amax = fmax(fmax(fabs(g.ox), fabs(h_bbox_sel6[3])), fmax(fmax(fabs(g.oy), fabs(h_bbox_sel6[4])),
                                                                     fmax(fabs(g.oz), fabs(h_bbox_sel6[5]))));
    g.slack = cell * 1e-6 + amax * 1e-12;
    g.extent = emax > cell ? emax : cell;
    return knn_impl(d_x, d_y, d_z, n, d_sel, m, g, knn, max_nn, radius, d_knn, d_knn_d2, d_normals, d_cov,
                    d_curv, threshold_curve, d_seed_flag, d_order, d_ws, ws_bytes, (cudaStream_t)stream);
}

int upcp_knn_normals(const double* d_x, const double* d_y, const double* d_z, int64_t n,
                     const uint8_t* d_sel, int64_t m, const double* h_bbox_sel6,
                     int knn, int max_nn, double radius, double cell_size,
                     int32_t* d_knn, double* d_knn_d2, double* d_normals, double* d_cov, double* d_curv,
                     double threshold_curve, uint8_t* d_seed_flag, void* d_ws, size_t ws_bytes, void* stream) {
    return knn_entry(d_x, d_y, d_z, n, d_sel, m, h_bbox_sel6, knn, max_nn, radius, cell_size, d_knn, d_knn_d2,
                     d_normals, d_cov, d_curv, threshold_curve, d_seed_flag, nullptr, d_ws, ws_bytes, stream);
}

int upcp_knn_normals_sorted(const double* d_x, const double* d_y, const double* d_z, int64_t n,
                            const uint8_t* d_sel, int64_t m, const double* h_bbox_sel6,
                            int knn, int max_nn, double radius, double cell_size,
                            int32_t* d_knn, double* d_normals, double* d_cov,
                            double threshold_curve, uint8_t* d_seed_flag, uint32_t* d_order,
                            void* d_ws, size_t ws_bytes, void* stream) {
    if (!d_order) UPCP_FAIL(UPCP_E_INVALID, "knn: d_order is required");
    return knn_entry(d_x, d_y, d_z, n, d_sel, m, h_bbox_sel6, knn, max_nn, radius, cell_size, d_knn, nullptr,
                     d_normals, d_cov, nullptr, threshold_curve, d_seed_flag, d_order, d_ws, ws_bytes, stream);
}

}  // extern "C"
