// Region growing over FIXED-RADIUS neighbourhoods: RegionGrowing._region_growing(method='radius')
// (reference region_growing.py:59-73,104-109: KDTreeFlann.search_radius_vector_3d(seed, grow_region_radius) --
// every point with squared distance < radius^2, ascending, the first returned one dropped).
//
// A radius neighbourhood of a dense mobile-LiDAR cloud holds hundreds to thousands of points, so no rows are
// stored: the selected points are counting-sorted into a dense uniform grid with cells of about radius / 2
// (the same structure as the kNN stage, knn.cu), and whoever needs a neighbourhood walks the (2R+1)^3 cell
// block around the point -- (2R+1)^2 contiguous ranges of the sorted records:
//   * radius_curv_kernel (lane per point): cumulant covariance of the neighbourhood and the curvature seed
//     flag eig_val[0] / sum < threshold_curve.  The reference folds the neighbours in ascending distance
//     order; here they come in grid order, so the covariance is accumulated RELATIVE TO THE POINT (no
//     cancellation, ~1e-13 accurate) and the flag is only decided on the device when it holds for every
//     matrix within the rounding noise the reference's unshifted single-pass sum can carry (a statistical
//     bound: 6 sigma with sigma = u max|p|^2 sqrt(n / 3)); everything else is flagged 2 = undecided and resolved by the
//     caller from the exact, ordered neighbour list (radius_neighbours_kernel) with LAPACK, as the
//     reference computes it;
//   * radius_grow_kernel: the persistent work-queue kernel of region.cu with a WARP per seed: lanes stride over
//     the ranges of the block, test distance, membership and the normal angle, claim accepted points with an
//     atomic exchange and push new seeds.  The region is the same least fixpoint as for the kNN graph, so the
//     expansion order is free.
// "The first returned index is dropped" (region_growing.py:112): the nearest point of a seed is the seed
// itself unless exact duplicates exist, in which case it is the duplicate with the smallest compact id
// (ordering (squared distance, compact id), as everywhere).
#include <math.h>
#include "common.cuh"
#include "primitives.cuh"
#include "upcp_math.cuh"

namespace upcp {

constexpr int kRBlock = 256;
constexpr int kRGrowThreads = 256;
constexpr uint32_t kRFull = 0xffffffffu;

struct RGrid {
    double ox, oy, oz, cell, inv_cell;
    int nx, ny, nz, R;
};
struct __align__(32) RRec { double x, y, z; uint32_t cid, pad; };

struct RCounters {      // same protocol as GrowCounters (region.cu)
    unsigned int head, tail;
    int pending;
    unsigned int error;
    unsigned long long region_size;
    unsigned long long nb_total;       // radius_neighbours: list cursor
};

__device__ __forceinline__ int rcell(double v, double origin, double inv_cell, int n) {
    double t = (v - origin) * inv_cell;
    if (!(t > 0.0)) return 0;
    if (t >= (double)n) return n - 1;
    return (int)floor(t);
}
__device__ __forceinline__ RRec load_rec(const RRec* __restrict__ p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    RRec r; r.x = a.x; r.y = a.y; r.z = b.x;
    r.cid = (uint32_t)__double_as_longlong(b.y); r.pad = 0;
    return r;
}

struct RCompactIdx {
    uint32_t* sel_idx;
    __device__ void operator()(int64_t i, bool selected, uint32_t rank) const { if (selected) sel_idx[rank] = (uint32_t)i; }
};

__global__ void __launch_bounds__(kRBlock)
radius_keys_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                   const uint32_t* __restrict__ sel_idx, int64_t n, int64_t m, RGrid g, uint32_t* __restrict__ keys,
                   uint32_t* __restrict__ rank, uint32_t* __restrict__ cell_count) {
    const int64_t stride = (int64_t)gridDim.x * kRBlock;
    for (int64_t j = (int64_t)blockIdx.x * kRBlock + threadIdx.x; j < m; j += stride) {
        int64_t i = sel_idx ? (int64_t)sel_idx[j] : j;
        if (i >= n) i = n - 1;
        const uint32_t key = ((uint32_t)rcell(z[i], g.oz, g.inv_cell, g.nz) * (uint32_t)g.ny + (uint32_t)rcell(y[i], g.oy, g.inv_cell, g.ny)) *
                                 (uint32_t)g.nx + (uint32_t)rcell(x[i], g.ox, g.inv_cell, g.nx);
        keys[j] = key;
        rank[j] = atomicAdd(cell_count + key, 1u);
    }
}

__global__ void __launch_bounds__(kRBlock)
radius_place_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                    const uint32_t* __restrict__ sel_idx, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rank,
                    const uint32_t* __restrict__ cell_first, int64_t n, int64_t m, const double* __restrict__ normals_c,
                    const uint8_t* __restrict__ seed_c, RRec* __restrict__ recs, double* __restrict__ normals_s,
                    int32_t* __restrict__ state, uint32_t* __restrict__ pos_of_cid) {
    const int64_t stride = (int64_t)gridDim.x * kRBlock;
    for (int64_t j = (int64_t)blockIdx.x * kRBlock + threadIdx.x; j < m; j += stride) {
        const uint32_t s = cell_first[keys[j]] + rank[j];
        int64_t i = sel_idx ? (int64_t)sel_idx[j] : j;
        if (i >= n) i = n - 1;
        double2* o = reinterpret_cast<double2*>(recs + s);
        o[0] = make_double2(x[i], y[i]);
        o[1] = make_double2(z[i], __longlong_as_double((long long)(unsigned long long)j));
        normals_s[(size_t)s * 3] = normals_c[(size_t)j * 3]; normals_s[(size_t)s * 3 + 1] = normals_c[(size_t)j * 3 + 1];
        normals_s[(size_t)s * 3 + 2] = normals_c[(size_t)j * 3 + 2];
        state[s] = seed_c[j] != 0 ? 1 : 0;
        pos_of_cid[j] = s;
    }
}

// ---- curvature seed flags ---------------------------------------------------------------
__global__ void __launch_bounds__(kRBlock)
radius_curv_kernel(const RRec* __restrict__ recs, const uint32_t* __restrict__ cell_first, RGrid g, int64_t m, double r2,
                   double thr, uint8_t* __restrict__ flag_c, uint32_t* __restrict__ nb_count_c) {
    const int64_t stride = (int64_t)gridDim.x * kRBlock;
    const int R = g.R;
    for (int64_t s = (int64_t)blockIdx.x * kRBlock + threadIdx.x; s < m; s += stride) {
        const RRec q = load_rec(recs + s);
        const int cx = rcell(q.x, g.ox, g.inv_cell, g.nx), cy = rcell(q.y, g.oy, g.inv_cell, g.ny);
        const int cz = rcell(q.z, g.oz, g.inv_cell, g.nz);
        const int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
        double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        double m2 = 0.0;
        uint32_t nn = 0;
        for (int zz = max(cz - R, 0); zz <= min(cz + R, g.nz - 1); ++zz)
            for (int yy = max(cy - R, 0); yy <= min(cy + R, g.ny - 1); ++yy) {
                const uint32_t row = ((uint32_t)zz * (uint32_t)g.ny + (uint32_t)yy) * (uint32_t)g.nx;
                const uint32_t p1 = __ldg(cell_first + row + x1 + 1);
                for (uint32_t p = __ldg(cell_first + row + x0); p < p1; ++p) {
                    const RRec v = load_rec(recs + p);
                    if (!(dist2(v.x, v.y, v.z, q.x, q.y, q.z) < r2)) continue;
                    const double dx = v.x - q.x, dy = v.y - q.y, dz = v.z - q.z;      // relative: no cancellation
                    c[0] += dx; c[1] += dy; c[2] += dz;
                    c[3] += dx * dx; c[4] += dx * dy; c[5] += dx * dz; c[6] += dy * dy; c[7] += dy * dz; c[8] += dz * dz;
                    m2 = fmax(m2, (v.x * v.x + v.y * v.y) + v.z * v.z);
                    ++nn;
                }
            }
        uint8_t flag = 2;
        if (nn == 1) {
            flag = 0;                            // the point alone: covariance exactly 0, 0 / 0 = NaN, NaN < thr is False
        } else if (nn >= 2) {
            const double dn = (double)nn;
            double mean[3] = {c[0] / dn, c[1] / dn, c[2] / dn};
            double cov6[6] = {c[3] / dn - mean[0] * mean[0], c[4] / dn - mean[0] * mean[1], c[5] / dn - mean[0] * mean[2],
                              c[6] / dn - mean[1] * mean[1], c[7] / dn - mean[1] * mean[2], c[8] / dn - mean[2] * mean[2]};
            double ev[3];
            fast_eigen3x3(cov6, ev);
            // The reference's matrix is this one plus the rounding noise E of its unshifted single-pass sums: every
            // entry carries about sigma = u |p|^2 sqrt(n / 3) (n sequential additions of terms of size |p|^2, divided
            // by n).  tau = 6 sigma bounds ||E||_2 (twice its expected Frobenius norm) and |trace E|; the flag is decided
            // here only if it is the same for every eigenvalue in [ev - tau, ev + tau] and every sum in
            // [sum - tau, sum + tau] -- a wrong decision would need noise of twice that size again.
            const double tau = 6.0 * 1.1102230246251565e-16 * sqrt(dn / 3.0) * m2 + 1e-15 * (fabs(ev[0]) + fabs(ev[1]) + fabs(ev[2]));
            const double sum = ev[0] + ev[1] + ev[2];
            bool below = isfinite(sum) && sum - tau > 0.0, above = below;
            for (int i = 0; i < 3 && (below || above); ++i) {
                const double hi = (ev[i] + tau) / (ev[i] + tau >= 0.0 ? sum - tau : sum + tau);
                const double lo = (ev[i] - tau) / (ev[i] - tau >= 0.0 ? sum + tau : sum - tau);
                below = below && hi < thr;
                above = above && lo >= thr;
            }
            flag = below ? 1 : (above ? 0 : 2);
        }
        flag_c[q.cid] = flag;
        if (nb_count_c) nb_count_c[q.cid] = nn;
    }
}

// ---- ordered neighbour lists of a few points (host resolution of undecided flags) ----------
__global__ void __launch_bounds__(kRBlock)
radius_neighbours_kernel(const RRec* __restrict__ recs, const uint32_t* __restrict__ cell_first, RGrid g, double r2,
                         const uint32_t* __restrict__ pos_of_cid, const int64_t* __restrict__ queries, int64_t nq,
                         long long* __restrict__ offsets, double* __restrict__ out4, long long cap, RCounters* __restrict__ ctr) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kRBlock + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * kRBlock) >> 5;
    const int R = g.R;
    for (int64_t qi = warp; qi < nq; qi += n_warps) {
        const RRec q = load_rec(recs + pos_of_cid[queries[qi]]);
        const int cx = rcell(q.x, g.ox, g.inv_cell, g.nx), cy = rcell(q.y, g.oy, g.inv_cell, g.ny);
        const int cz = rcell(q.z, g.oz, g.inv_cell, g.nz);
        const int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
        unsigned long long base = 0;
        for (int pass = 0; pass < 2; ++pass) {
            uint32_t cnt = 0;
            for (int zz = max(cz - R, 0); zz <= min(cz + R, g.nz - 1); ++zz)
                for (int yy = max(cy - R, 0); yy <= min(cy + R, g.ny - 1); ++yy) {
                    const uint32_t row = ((uint32_t)zz * (uint32_t)g.ny + (uint32_t)yy) * (uint32_t)g.nx;
                    const uint32_t p0 = __ldg(cell_first + row + x0), p1 = __ldg(cell_first + row + x1 + 1);
                    for (uint32_t pb = p0; pb < p1; pb += 32) {
                        const uint32_t p = pb + lane;
                        bool in = false;
                        RRec v = q;
                        if (p < p1) { v = load_rec(recs + p); in = dist2(v.x, v.y, v.z, q.x, q.y, q.z) < r2; }
                        const uint32_t bal = __ballot_sync(kRFull, in);
                        if (pass == 1 && in) {
                            const unsigned long long o = base + cnt + __popc(bal & ((1u << lane) - 1u));
                            if ((long long)o < cap) {
                                out4[o * 4] = v.x; out4[o * 4 + 1] = v.y; out4[o * 4 + 2] = v.z; out4[o * 4 + 3] = (double)v.cid;
                            }
                        }
                        cnt += __popc(bal);
                    }
                }
            if (pass == 0) {
                if (lane == 0) { base = atomicAdd(&ctr->nb_total, (unsigned long long)cnt); offsets[2 * qi] = (long long)base; offsets[2 * qi + 1] = (long long)cnt; }
                base = __shfl_sync(kRFull, base, 0);
            }
        }
    }
}

// ---- growth ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRBlock)
radius_queue_seeds_kernel(const int32_t* __restrict__ state, int64_t m, int32_t* __restrict__ queue, RCounters* __restrict__ ctr) {
    const int lane = threadIdx.x & 31;
    const int64_t total = (m + 31) & ~(int64_t)31;
    const int64_t stride = (int64_t)gridDim.x * kRBlock;
    for (int64_t s = (int64_t)blockIdx.x * kRBlock + threadIdx.x; s < total; s += stride) {
        const bool push = s < m && state[s] != 0;
        const uint32_t bal = __ballot_sync(kRFull, push);
        if (bal) {
            const int leader = __ffs(bal) - 1;
            unsigned int base = 0;
            if (lane == leader) { atomicAdd(&ctr->pending, __popc(bal)); base = atomicAdd(&ctr->tail, (unsigned int)__popc(bal)); }
            base = __shfl_sync(kRFull, base, leader);
            if (push) queue[base + __popc(bal & ((1u << lane) - 1u))] = (int32_t)s;
        }
    }
}

__global__ void __launch_bounds__(kRBlock)
radius_requeue_kernel(const int64_t* __restrict__ extra_c, int64_t n_extra, const uint32_t* __restrict__ pos_of_cid,
                      const uint8_t* __restrict__ can_seed_c, const int32_t* __restrict__ state, int32_t* __restrict__ queue,
                      RCounters* __restrict__ ctr) {
    const int lane = threadIdx.x & 31;
    const int64_t total = (n_extra + 31) & ~(int64_t)31;
    const int64_t stride = (int64_t)gridDim.x * kRBlock;
    for (int64_t i = (int64_t)blockIdx.x * kRBlock + threadIdx.x; i < total; i += stride) {
        bool push = false;
        int32_t s = -1;
        if (i < n_extra) {
            s = (int32_t)pos_of_cid[extra_c[i]];
            push = state[s] == 1 && can_seed_c[extra_c[i]] == 1;     // in the region, not expanded yet (state 2 = was a seed), a seed now
        }
        const uint32_t bal = __ballot_sync(kRFull, push);
        if (bal) {
            const int leader = __ffs(bal) - 1;
            unsigned int base = 0;
            if (lane == leader) { atomicAdd(&ctr->pending, __popc(bal)); base = atomicAdd(&ctr->tail, (unsigned int)__popc(bal)); }
            base = __shfl_sync(kRFull, base, leader);
            if (push) queue[base + __popc(bal & ((1u << lane) - 1u))] = s;
        }
    }
}
__global__ void radius_rewind_kernel(RCounters* __restrict__ ctr) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { ctr->head = ctr->tail; ctr->pending = 0; }
}

// state: 0 outside, 1 in the region, 2 in the region AND expanded as a seed
__global__ void __launch_bounds__(kRGrowThreads)
radius_grow_kernel(const RRec* __restrict__ recs, const uint32_t* __restrict__ cell_first, RGrid g, double r2,
                   const double* __restrict__ normals, const uint8_t* __restrict__ can_seed_c, int64_t m, double threshold_deg,
                   int32_t* __restrict__ state, int32_t* __restrict__ queue, RCounters* __restrict__ ctr, int sleep_ns) {
    const int lane = threadIdx.x & 31;
    volatile int32_t* vqueue = queue;
    volatile int* vpending = &ctr->pending;
    const int R = g.R;

    auto expand = [&](int32_t s) {
        const RRec q = load_rec(recs + s);
        const Vec3 ns = {normals[(size_t)s * 3], normals[(size_t)s * 3 + 1], normals[(size_t)s * 3 + 2]};
        const int cx = rcell(q.x, g.ox, g.inv_cell, g.nx), cy = rcell(q.y, g.oy, g.inv_cell, g.ny);
        const int cz = rcell(q.z, g.oz, g.inv_cell, g.nz);
        // "the first returned index is dropped": the smallest compact id among the points at distance zero
        uint32_t first = q.cid;
        {
            const uint32_t c = ((uint32_t)cz * (uint32_t)g.ny + (uint32_t)cy) * (uint32_t)g.nx + (uint32_t)cx;
            const uint32_t p1 = __ldg(cell_first + c + 1);
            for (uint32_t p = __ldg(cell_first + c) + lane; p < p1; p += 32) {
                const RRec v = load_rec(recs + p);
                if (v.x == q.x && v.y == q.y && v.z == q.z) first = min(first, v.cid);
            }
            first = __reduce_min_sync(kRFull, first);
        }
        const int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
        for (int zz = max(cz - R, 0); zz <= min(cz + R, g.nz - 1); ++zz)
            for (int yy = max(cy - R, 0); yy <= min(cy + R, g.ny - 1); ++yy) {
                const uint32_t row = ((uint32_t)zz * (uint32_t)g.ny + (uint32_t)yy) * (uint32_t)g.nx;
                const uint32_t p0 = __ldg(cell_first + row + x0), p1 = __ldg(cell_first + row + x1 + 1);
                for (uint32_t pb = p0; pb < p1; pb += 32) {
                    const uint32_t p = pb + lane;
                    bool push = false;
                    if (p < p1 && __ldcg(state + p) == 0) {
                        const RRec v = load_rec(recs + p);
                        if (dist2(v.x, v.y, v.z, q.x, q.y, q.z) < r2 && v.cid != first) {
                            const Vec3 nn = {normals[(size_t)p * 3], normals[(size_t)p * 3 + 1], normals[(size_t)p * 3 + 2]};
                            if (vector_angle_deg(ns, nn) < threshold_deg && atomicExch(&state[p], 1) == 0)
                                push = can_seed_c[v.cid] == 1;
                        }
                    }
                    const uint32_t bal = __ballot_sync(kRFull, push);
                    if (bal) {
                        const int leader = __ffs(bal) - 1;
                        unsigned int base = 0;
                        if (lane == leader) { atomicAdd(&ctr->pending, __popc(bal)); base = atomicAdd(&ctr->tail, (unsigned int)__popc(bal)); }
                        base = __shfl_sync(kRFull, base, leader);
                        if (push) vqueue[base + __popc(bal & ((1u << lane) - 1u))] = (int32_t)p;
                    }
                }
            }
        if (lane == 0) state[s] = 2;
    };

    while (true) {
        unsigned int t0 = 0;
        if (lane == 0) t0 = atomicAdd(&ctr->head, 32u);
        t0 = __shfl_sync(kRFull, t0, 0);
        if ((int64_t)t0 >= m) return;                 // at most m seeds can ever be queued
        int32_t mine = -1;
        uint32_t remaining = __ballot_sync(kRFull, (int64_t)t0 + lane < m);
        unsigned int spins = 0;
        while (remaining) {
            if (((remaining >> lane) & 1u) && mine < 0) mine = vqueue[t0 + lane];
            const uint32_t ready = __ballot_sync(kRFull, ((remaining >> lane) & 1u) && mine >= 0);
            if (ready == 0) {
                int pend = 1;
                if (lane == 0) pend = *vpending;
                pend = __shfl_sync(kRFull, pend, 0);
                if (pend <= 0) return;                 // global quiescence
                __nanosleep(sleep_ns);
                if (++spins > (1u << 26)) { if (lane == 0) atomicExch(&ctr->error, 1u); return; }   // safety valve
                continue;
            }
            uint32_t todo = ready;
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                expand(__shfl_sync(kRFull, mine, src));
            }
            __syncwarp();
            if (lane == 0) { __threadfence(); atomicSub(&ctr->pending, __popc(ready)); }
            remaining &= ~ready;
            spins = 0;
        }
    }
}

__global__ void __launch_bounds__(kRBlock)
radius_finish_kernel(const RRec* __restrict__ recs, const int32_t* __restrict__ state, int64_t m, uint8_t* __restrict__ region_c,
                     RCounters* __restrict__ ctr) {
    int64_t stride = (int64_t)gridDim.x * kRBlock;
    unsigned long long cnt = 0;
    for (int64_t s = (int64_t)blockIdx.x * kRBlock + threadIdx.x; s < m; s += stride) {
        const uint8_t r = state[s] != 0;
        region_c[load_rec(recs + s).cid] = r;
        cnt += r;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kRFull, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&ctr->region_size, cnt);
}
__global__ void radius_info_kernel(const RCounters* __restrict__ ctr, long long* __restrict__ info) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { info[0] = (long long)ctr->region_size; info[1] = (long long)ctr->tail; }
}

static inline uint64_t rgrid_cap(int64_t m) {
    uint64_t cap = 16ull * (uint64_t)(m < 1 ? 1 : m) + (1ull << 20);
    return cap < (1ull << 28) ? cap : (1ull << 28);
}

struct RWs {
    RCounters* ctr; uint32_t* sel_idx; uint32_t* prim; uint32_t* keys; uint32_t* rank; RRec* recs; double* normals;
    int32_t* state; int32_t* queue; uint32_t* pos_of_cid; uint32_t* cell_first;
};

static int rgrid_from(const double* b, double radius, int64_t m, RGrid* g) {
    if (!b || !(radius > 0.0) || !isfinite(radius)) UPCP_FAIL(UPCP_E_INVALID, "radius: a finite radius > 0 is required");
    const double ex = b[3] - b[0], ey = b[4] - b[1], ez = b[5] - b[2];
    if (!(ex >= 0) || !(ey >= 0) || !(ez >= 0) || !isfinite(ex + ey + ez)) UPCP_FAIL(UPCP_E_INVALID, "radius: bounding box is not finite");
    double cell = radius / 2.0;
    const double cap = (double)rgrid_cap(m);
    while ((floor(ex / cell) + 1) * (floor(ey / cell) + 1) * (floor(ez / cell) + 1) > cap) cell *= 1.0625;
    g->ox = b[0]; g->oy = b[1]; g->oz = b[2]; g->cell = cell; g->inv_cell = 1.0 / cell;
    g->nx = (int)floor(ex / cell) + 1; g->ny = (int)floor(ey / cell) + 1; g->nz = (int)floor(ez / cell) + 1;
    // the block must reach `radius` beyond the point's own cell whatever the rounding of the cell index
    const double amax = fmax(fmax(fabs(b[0]), fabs(b[3])), fmax(fmax(fabs(b[1]), fabs(b[4])), fmax(fabs(b[2]), fabs(b[5]))));
    g->R = (int)ceil((radius + cell * 1e-6 + amax * 1e-12) / cell);
    if (g->R < 1) g->R = 1;
    return UPCP_OK;
}

static int rws(void* d_ws, size_t bytes, int64_t n, int64_t m, bool with_sel, RWs* w) {
    Workspace ws(d_ws, bytes);
    w->ctr = (RCounters*)ws.take<char>(256);
    w->sel_idx = ws.take<uint32_t>((size_t)n);
    w->prim = ws.take<uint32_t>(compact_ws_words(n) + scan_ws_words((int64_t)rgrid_cap(m) + 1) + 64);
    w->keys = ws.take<uint32_t>((size_t)m);
    w->rank = ws.take<uint32_t>((size_t)m);
    w->recs = ws.take<RRec>((size_t)m);
    w->normals = ws.take<double>((size_t)m * 3);
    w->state = ws.take<int32_t>((size_t)m);
    w->queue = ws.take<int32_t>((size_t)m);
    w->pos_of_cid = ws.take<uint32_t>((size_t)m);
    w->cell_first = ws.take<uint32_t>((size_t)rgrid_cap(m) + 1);
    (void)with_sel;
    if (!ws.ok) UPCP_FAIL(UPCP_E_WORKSPACE, "radius: workspace too small");
    return UPCP_OK;
}

}  // namespace upcp

using namespace upcp;

extern "C" {

size_t upcp_radius_ws_bytes(int64_t n, int64_t m) {
    if (n < 1) n = 1;
    if (m < 1 || m > n) m = n;
    const size_t gcap = (size_t)rgrid_cap(m);
    return align_up(256) + align_up((size_t)n * 4) + align_up((compact_ws_words(n) + scan_ws_words((int64_t)gcap + 1) + 64) * 4) +
           4 * align_up((size_t)m * 4) + align_up((size_t)m * 32) + align_up((size_t)m * 24) + align_up((size_t)m * 4) +
           align_up((gcap + 1) * 4) + 4096;
}

int upcp_radius_prepare(const double* d_x, const double* d_y, const double* d_z, int64_t n, const uint8_t* d_sel, int64_t m,
                        const double* h_bbox_sel6, double radius, const double* d_normals, const uint8_t* d_seed,
                        double threshold_curve, uint8_t* d_flag, uint32_t* d_nb_count, void* d_ws, size_t ws_bytes, void* stream) {
    if (n < 0 || m < 0 || m > n || !d_normals || !d_seed || !d_flag) UPCP_FAIL(UPCP_E_INVALID, "radius_prepare: bad argument");
    if (n >= ((int64_t)1 << 31)) UPCP_FAIL(UPCP_E_CAPACITY, "radius_prepare: n must be < 2^31");
    if (m == 0) return UPCP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    RGrid g;
    UPCP_TRY(rgrid_from(h_bbox_sel6, radius, m, &g));
    RWs w;
    UPCP_TRY(rws(d_ws, ws_bytes, n, m, d_sel != nullptr, &w));
    const uint64_t G = (uint64_t)g.nx * g.ny * g.nz;
    UPCP_CUDA_OK(cudaMemsetAsync(w.ctr, 0, 256, st));
    uint32_t* sel_idx = nullptr;
    if (d_sel) {
        RCompactIdx f{w.sel_idx};
        UPCP_TRY(compact_apply(d_sel, n, w.prim, w.prim + 64, st, f));
        sel_idx = w.sel_idx;
    } else if (m != n) UPCP_FAIL(UPCP_E_INVALID, "radius_prepare: m must equal n when no selection is given");
    UPCP_CUDA_OK(cudaMemsetAsync(w.cell_first, 0, ((size_t)G + 1) * 4, st));
    radius_keys_kernel<<<grid_for(m, kRBlock), kRBlock, 0, st>>>(d_x, d_y, d_z, sel_idx, n, m, g, w.keys, w.rank, w.cell_first);
    UPCP_LAUNCH_OK();
    UPCP_TRY(exclusive_scan_u32(w.cell_first, w.cell_first, (int64_t)G + 1, nullptr, w.prim + 64, st));
    radius_place_kernel<<<grid_for(m, kRBlock), kRBlock, 0, st>>>(d_x, d_y, d_z, sel_idx, w.keys, w.rank, w.cell_first, n, m,
                                                                  d_normals, d_seed, w.recs, w.normals, w.state, w.pos_of_cid);
    UPCP_LAUNCH_OK();
    radius_curv_kernel<<<grid_for(m, kRBlock, 16), kRBlock, 0, st>>>(w.recs, w.cell_first, g, m, radius * radius, threshold_curve,
                                                                     d_flag, d_nb_count);
    UPCP_LAUNCH_OK();
    UPCP_CUDA_OK(cudaMemsetAsync(w.queue, 0xff, (size_t)m * 4, st));
    radius_queue_seeds_kernel<<<grid_for(m, kRBlock), kRBlock, 0, st>>>(w.state, m, w.queue, w.ctr);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

int upcp_radius_grow(int64_t n, int64_t m, const double* h_bbox_sel6, double radius, const uint8_t* d_can_seed,
                     double threshold_angle_deg, const int64_t* d_extra, int64_t n_extra,
                     void* d_ws, size_t ws_bytes, void* stream) {
    if (m < 0 || !d_can_seed || n_extra < 0) UPCP_FAIL(UPCP_E_INVALID, "radius_grow: bad argument");
    if (m == 0) return UPCP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    RGrid g;
    UPCP_TRY(rgrid_from(h_bbox_sel6, radius, m, &g));
    RWs w;
    UPCP_TRY(rws(d_ws, ws_bytes, n, m, true, &w));
    if (d_extra) {
        if (n_extra == 0) return UPCP_OK;
        radius_rewind_kernel<<<1, 32, 0, st>>>(w.ctr);
        UPCP_LAUNCH_OK();
        radius_requeue_kernel<<<grid_for(n_extra, kRBlock), kRBlock, 0, st>>>(d_extra, n_extra, w.pos_of_cid, d_can_seed, w.state,
                                                                              w.queue, w.ctr);
        UPCP_LAUNCH_OK();
    }
    const upcp_tuning tune = tuning();
    int per_sm = 0;
    UPCP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, radius_grow_kernel, kRGrowThreads, 0));
    if (per_sm < 1) UPCP_FAIL(UPCP_E_CUDA, "radius_grow: kernel cannot be resident");
    if (per_sm > 2 * tune.grow_ctas_per_sm) per_sm = 2 * tune.grow_ctas_per_sm;       // a warp per seed: throughput, not a thin frontier
    int grid = num_sms() * per_sm;
    int64_t need = (m + (kRGrowThreads / 32) - 1) / (kRGrowThreads / 32);
    if (need < grid) grid = (int)(need < 1 ? 1 : need);
    prof_begin(UPCP_PROF_GROW, st);
    radius_grow_kernel<<<grid, kRGrowThreads, 0, st>>>(w.recs, w.cell_first, g, radius * radius, w.normals, d_can_seed, m,
                                                      threshold_angle_deg, w.state, w.queue, w.ctr, 64);
    prof_end(UPCP_PROF_GROW, st, m);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

int upcp_radius_region(int64_t n, int64_t m, uint8_t* d_region, int64_t* d_info, void* d_ws, size_t ws_bytes, void* stream) {
    if (m < 0 || !d_region) UPCP_FAIL(UPCP_E_INVALID, "radius_region: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (d_info) UPCP_CUDA_OK(cudaMemsetAsync(d_info, 0, 2 * sizeof(int64_t), st));
    if (m == 0) return UPCP_OK;
    RWs w;
    UPCP_TRY(rws(d_ws, ws_bytes, n, m, true, &w));
    UPCP_CUDA_OK(cudaMemsetAsync(&w.ctr->region_size, 0, sizeof(unsigned long long), st));
    radius_finish_kernel<<<grid_for(m, kRBlock), kRBlock, 0, st>>>(w.recs, w.state, m, d_region, w.ctr);
    UPCP_LAUNCH_OK();
    if (d_info) { radius_info_kernel<<<1, 32, 0, st>>>(w.ctr, (long long*)d_info); UPCP_LAUNCH_OK(); }
    unsigned int h_err = 0;
    UPCP_CUDA_OK(cudaMemcpyAsync(&h_err, &w.ctr->error, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    UPCP_CUDA_OK(cudaStreamSynchronize(st));
    if (h_err) UPCP_FAIL(UPCP_E_INTERNAL, "radius_grow: work queue did not quiesce");
    return UPCP_OK;
}

int upcp_radius_neighbours(int64_t n, int64_t m, const double* h_bbox_sel6, double radius, const int64_t* d_queries, int64_t nq,
                           int64_t* d_offsets2, double* d_out4, int64_t cap, int64_t* d_total, void* d_ws, size_t ws_bytes,
                           void* stream) {
    if (m < 0 || nq < 0 || cap < 0 || !d_offsets2 || !d_total) UPCP_FAIL(UPCP_E_INVALID, "radius_neighbours: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    UPCP_CUDA_OK(cudaMemsetAsync(d_total, 0, sizeof(int64_t), st));
    if (m == 0 || nq == 0) return UPCP_OK;
    RGrid g;
    UPCP_TRY(rgrid_from(h_bbox_sel6, radius, m, &g));
    RWs w;
    UPCP_TRY(rws(d_ws, ws_bytes, n, m, true, &w));
    UPCP_CUDA_OK(cudaMemsetAsync(&w.ctr->nb_total, 0, sizeof(unsigned long long), st));
    radius_neighbours_kernel<<<grid_for(nq * 32, kRBlock, 16), kRBlock, 0, st>>>(w.recs, w.cell_first, g, radius * radius, w.pos_of_cid,
                                                                                 d_queries, nq, (long long*)d_offsets2, d_out4,
                                                                                 (long long)cap, w.ctr);
    UPCP_LAUNCH_OK();
    UPCP_CUDA_OK(cudaMemcpyAsync(d_total, &w.ctr->nb_total, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    return UPCP_OK;
}

}  // extern "C"
